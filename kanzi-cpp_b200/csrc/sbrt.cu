// sbrt.cu -- Sort-By-Rank transforms (kanzi RANK = SBR(1/2), MTFT = SBR(0)) on sm_100a.
//
// Reference: transform/SBRT.cpp:46-97 (forward), :99-145 (inverse).
// The reference keeps a 256-entry list ordered by key q (ties: most recently
// moved first) and updates it symbol by symbol.  The list order is a pure
// function of each symbol's last two access times, so the forward direction is
// evaluated tile-parallel:
//   rank_i(c) = #{ s != c : (q_s, t_s) > (q_c, t_c) },  q = ((t1&m1) + (t2&m2)) >> sh,
//   t = last access time (never accessed: q = 0, t = -1-s; first access pairs with 0)
//   1. per 4 KiB tile: last two occurrences of every symbol (smem atomicMax)
//   2. per block: fold the tile tables into per-tile ENTRY tables (thread = symbol)
//   3. per tile: one warp replays the tile from its entry table; the 256 key
//      comparisons of a step are 8 ballots.
// The inverse needs the decoded symbol to update the list, so it is a serial
// replay per block (one thread per block, list in shared memory).
#include "common.cuh"
#include "kernels.h"

#define S_TILE 4096

__device__ __forceinline__ void sbrt_masks(int mode, u32& m1, u32& m2, int& sh)
{
    m1 = (mode == 3) ? 0u : 0xFFFFFFFFu;
    m2 = (mode == 1) ? 0u : 0xFFFFFFFFu;
    sh = (mode == 2) ? 1 : 0;
}

// Decide step shared by forward and inverse: SBRT never refuses unless the
// destination is too small (SBRT.cpp:57-60); inverse honours the skip flag.
__global__ void sbrt_decide_kernel(StageLaunch L, int inverse)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= L.nBlocks)
        return;
    const BlkState bs = L.stIn[b];
    BlkState ns = bs;
    const int bit = 1 << (7 - L.stageIdx);
    const int cap = (bs.swaps & 1) ? L.capOdd[b] : L.capEven[b];
    const bool skip = inverse ? ((bs.flags & bit) != 0) : (bs.len > cap);
    if (!skip) {
        ns.cur = next_cur(bs.cur);
        ns.swaps = bs.swaps + 1;
        if (!inverse)
            ns.flags = bs.flags & ~bit;
        else if (bs.len > cap)
            atomicExch(L.errFlag, KERR_BAD_STREAM);
    }
    L.stOut[b] = ns;
}

// 1. last two occurrences (positions + 1, 0 = none) of each symbol inside a tile
__global__ void __launch_bounds__(256)
sbrt_occ_kernel(BufTable bt, const BlkState* __restrict__ stIn, const BlkState* __restrict__ stOut, int maxTiles,
                uint2* __restrict__ occ)
{
    __shared__ u32 s_t1[256], s_t2[256];
    const int b = blockIdx.y, t = blockIdx.x;
    const BlkState bs = stIn[b];
    if (stOut[b].swaps == bs.swaps)
        return;
    const int n = bs.len;
    const int base = t * S_TILE;
    if (base >= n)
        return;
    const u8* __restrict__ src = blk_src(bt, bs, b);
    s_t1[threadIdx.x] = 0;
    s_t2[threadIdx.x] = 0;
    __syncthreads();
    const int end = min(base + S_TILE, n);
    for (int i = base + threadIdx.x; i < end; i += 256)
        atomicMax(&s_t1[src[i]], (u32)(i + 1));
    __syncthreads();
    for (int i = base + threadIdx.x; i < end; i += 256) {
        const u32 c = src[i];
        if ((u32)(i + 1) != s_t1[c])
            atomicMax(&s_t2[c], (u32)(i + 1));
    }
    __syncthreads();
    occ[((i64)b * maxTiles + t) * 256 + threadIdx.x] = make_uint2(s_t1[threadIdx.x], s_t2[threadIdx.x]);
}

// 2. exclusive fold over tiles: occ[tile] becomes the table valid at the tile's first position
__global__ void __launch_bounds__(256)
sbrt_fold_kernel(const BlkState* __restrict__ stIn, const BlkState* __restrict__ stOut, int maxTiles,
                 uint2* __restrict__ occ)
{
    const int b = blockIdx.x;
    const BlkState bs = stIn[b];
    if (stOut[b].swaps == bs.swaps)
        return;
    const int tiles = (bs.len + S_TILE - 1) / S_TILE;
    u32 p1 = 0, p2 = 0;
    uint2* o = occ + (i64)b * maxTiles * 256 + threadIdx.x;
    for (int t = 0; t < tiles; t++) {
        const uint2 a = o[(i64)t * 256];
        o[(i64)t * 256] = make_uint2(p1, p2);
        if (a.x) {
            p2 = a.y ? a.y : p1;
            p1 = a.x;
        }
    }
}

__device__ __forceinline__ u64 sbrt_key(u32 p1, u32 p2, int sym, u32 m1, u32 m2, int sh)
{
    // positions are stored +1; 0 = never.  p[c] starts at 0 in the reference, so a
    // missing second occurrence contributes time 0.
    if (p1 == 0)
        return (u64)(u32)(0x80000000u - 1u - (u32)sym); // q = 0, t = -1 - sym
    const u32 t1 = p1 - 1, t2 = p2 ? p2 - 1 : 0u;
    const u32 q = ((t1 & m1) + (t2 & m2)) >> sh;
    return ((u64)q << 32) | (u64)(t1 + 0x80000000u);
}

// 3. replay: one warp per tile
#define R_WARPS 4
__global__ void __launch_bounds__(R_WARPS * 32)
sbrt_rank_kernel(BufTable bt, const BlkState* __restrict__ stIn, const BlkState* __restrict__ stOut, int maxTiles,
                 const uint2* __restrict__ occ, int mode)
{
    __shared__ u64 s_keys[R_WARPS][256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int t = blockIdx.x * R_WARPS + w;
    const BlkState bs = stIn[b];
    if (stOut[b].swaps == bs.swaps)
        return;
    const int n = bs.len;
    const int base = t * S_TILE;
    if (base >= n)
        return;
    u32 m1, m2;
    int sh;
    sbrt_masks(mode, m1, m2, sh);
    const u8* __restrict__ src = blk_src(bt, bs, b);
    u8* __restrict__ dst = blk_dst(bt, bs, b);
    u64* keys = s_keys[w];
    const uint2* o = occ + ((i64)b * maxTiles + t) * 256;
    for (int s = lane; s < 256; s += 32) {
        const uint2 e = o[s];
        keys[s] = sbrt_key(e.x, e.y, s, m1, m2, sh);
    }
    __syncwarp();
    const int end = min(base + S_TILE, n);
    for (int g = base; g < end; g += 128) {
        // 128 positions per round: lane holds 4 input bytes, builds 4 output bytes
        u32 inw = 0;
        {
            const int p = g + 4 * lane;
            if (p + 4 <= n && ((((uintptr_t)(src + p)) & 3) == 0)) {
                inw = *reinterpret_cast<const u32*>(src + p);
            } else {
                for (int k = 0; k < 4; k++)
                    if (p + k < n)
                        inw |= (u32)src[p + k] << (8 * k);
            }
        }
        u32 outw = 0;
        const int cnt = min(128, end - g);
        for (int j = 0; j < cnt; j++) {
            const u32 wv = __shfl_sync(FULL_MASK, inw, j >> 2);
            const u32 c = (wv >> (8 * (j & 3))) & 0xFF;
            const u64 kc = keys[c];
            u32 rank = 0;
#pragma unroll
            for (int r = 0; r < 8; r++)
                rank += __popc(__ballot_sync(FULL_MASK, keys[r * 32 + lane] > kc));
            if (lane == (j >> 2))
                outw |= rank << (8 * (j & 3));
            const u32 i = (u32)(g + j);
            const int tlast = (int)((u32)kc - 0x80000000u); // < 0: never accessed
            const u32 prev = (tlast < 0) ? 0u : (u32)tlast;
            const u32 qc = ((i & m1) + (prev & m2)) >> sh;
            __syncwarp();
            if (lane == 0)
                keys[c] = ((u64)qc << 32) | (u64)(i + 0x80000000u);
            __syncwarp();
        }
        {
            const int p = g + 4 * lane;
            if (p + 4 <= end && ((((uintptr_t)(dst + p)) & 3) == 0)) {
                *reinterpret_cast<u32*>(dst + p) = outw;
            } else {
                for (int k = 0; k < 4; k++)
                    if (p + k < end)
                        dst[p + k] = (u8)(outw >> (8 * k));
            }
        }
    }
}

// Inverse: serial replay per block (SBRT.cpp:99-145).  The update needs the decoded
// symbol, so a block is ONE dependency chain; the kernel minimises the latency of a
// step: one warp per block with the 256-entry list distributed over the lanes in
// registers (lane L owns ranks 8L..8L+7; entry = key q and pb = (last access << 8) | symbol).
//   symbol at rank r      : uniform slot select + one shuffle from lane r>>3
//   new rank of the symbol: #entries with q > qc  (8 ballots; the list is sorted by q)
//   move-up               : every lane shifts its own slots, one shuffle-up for the seam
// No shared memory, no divergence; zero words (runs of rank 0) take a closed form.
// Input/output travel through registers 128 bytes at a time (coalesced, prefetched).
template <class PB>
__global__ void __launch_bounds__(32)
sbrt_inverse_kernel(BufTable bt, const BlkState* __restrict__ stIn, const BlkState* __restrict__ stOut, int mode)
{
    const int b = blockIdx.x;
    const BlkState bs = stIn[b];
    if (stOut[b].swaps == bs.swaps)
        return;
    const int n = bs.len;
    const int lane = threadIdx.x;
    const u8* __restrict__ src = blk_src(bt, bs, b);
    u8* __restrict__ dst = blk_dst(bt, bs, b);
    u32 m1, m2;
    int sh;
    sbrt_masks(mode, m1, m2, sh);
    int q[8];
    PB pb[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        q[k] = 0;
        pb[k] = (PB)(8 * lane + k);
    }
    const u32* __restrict__ src4 = reinterpret_cast<const u32*>(src); // buffers are 16-byte aligned
    const int groups = (n + 127) >> 7;
    u32 nextw = 0;
    {
        const int p = 4 * lane;
        if (p + 4 <= n)
            nextw = src4[lane];
        else
            for (int k = 0; p + k < n; k++)
                nextw |= (u32)src[p + k] << (8 * k);
    }
    for (int g = 0; g < groups; g++) {
        const u32 inw = nextw;
        {
            const int p = (g + 1) * 128 + 4 * lane; // prefetch the next 128 bytes
            nextw = 0;
            if (p + 4 <= n)
                nextw = src4[p >> 2];
            else
                for (int k = 0; p + k < n; k++)
                    nextw |= (u32)src[p + k] << (8 * k);
        }
        u32 outw = 0;
        const int base = g * 128;
        const int cnt = min(128, n - base);
        for (int j = 0; j < cnt; j += 4) {
            const u32 w4 = __shfl_sync(FULL_MASK, inw, j >> 2);
            const int lim = min(4, cnt - j);
            u32 o4 = 0;
            if (w4 == 0 && lim == 4) {
                // four accesses to the head of the list: only its key changes
                const PB e = __shfl_sync(FULL_MASK, pb[0], 0);
                const u32 c = (u32)(e & 0xFF);
                const u32 i3 = (u32)(base + j + 3);
                if (lane == 0) {
                    q[0] = (int)(((i3 & m1) + ((i3 - 1) & m2)) >> sh);
                    pb[0] = ((PB)i3 << 8) | (PB)c;
                }
                o4 = c * 0x01010101u;
            } else {
                for (int x = 0; x < lim; x++) {
                    const int r = (int)((w4 >> (8 * x)) & 0xFF);
                    const u32 i = (u32)(base + j + x);
                    const int slot = r & 7;
                    PB sel = pb[0];
#pragma unroll
                    for (int k = 1; k < 8; k++)
                        if (slot == k)
                            sel = pb[k];
                    const PB e = __shfl_sync(FULL_MASK, sel, r >> 3);
                    const u32 c = (u32)(e & 0xFF);
                    const u32 pc = (u32)(e >> 8);
                    const int qc = (int)(((i & m1) + (pc & m2)) >> sh);
                    const PB ne = ((PB)i << 8) | (PB)c;
                    o4 |= c << (8 * x);
                    if (r == 0) {
                        if (lane == 0) {
                            q[0] = qc;
                            pb[0] = ne;
                        }
                        continue;
                    }
                    int rp = 0;
#pragma unroll
                    for (int k = 0; k < 8; k++)
                        rp += __popc(__ballot_sync(FULL_MASK, q[k] > qc));
                    // ranks (rp, r] take the entry of the rank above; rank rp takes the new entry
                    const int pq = __shfl_up_sync(FULL_MASK, q[7], 1);
                    const PB ppb = __shfl_up_sync(FULL_MASK, pb[7], 1);
                    const int g0 = 8 * lane;
#pragma unroll
                    for (int k = 7; k >= 1; k--) {
                        const bool mv = (g0 + k > rp) && (g0 + k <= r);
                        q[k] = mv ? q[k - 1] : q[k];
                        pb[k] = mv ? pb[k - 1] : pb[k];
                    }
                    {
                        const bool mv = (g0 > rp) && (g0 <= r);
                        q[0] = mv ? pq : q[0];
                        pb[0] = mv ? ppb : pb[0];
                    }
#pragma unroll
                    for (int k = 0; k < 8; k++)
                        if (g0 + k == rp) {
                            q[k] = qc;
                            pb[k] = ne;
                        }
                }
            }
            if (lane == (j >> 2))
                outw = o4;
        }
        {
            const int p = base + 4 * lane;
            if (p + 4 <= n) {
                *reinterpret_cast<u32*>(dst + p) = outw;
            } else {
                for (int k = 0; k < 4; k++)
                    if (p + k < n)
                        dst[p + k] = (u8)(outw >> (8 * k));
            }
        }
    }
}

void launch_sbrt_forward(const StageLaunch& L, int mode, Workspace& ws, cudaStream_t s, u64* launches)
{
    const int maxTiles = (ws.capN + S_TILE - 1) / S_TILE;
    const int tiles = (L.maxLen + S_TILE - 1) / S_TILE;
    uint2* occ = reinterpret_cast<uint2*>(ws.occ);
    KLAUNCH(sbrt_decide_kernel, (L.nBlocks + 31) / 32, 32, s, L, 0);
    KLAUNCH(sbrt_occ_kernel, dim3(tiles, L.nBlocks), 256, s, L.bt, L.stIn, L.stOut, maxTiles, occ);
    KLAUNCH(sbrt_fold_kernel, L.nBlocks, 256, s, L.stIn, L.stOut, maxTiles, occ);
    KLAUNCH(sbrt_rank_kernel, dim3((tiles + R_WARPS - 1) / R_WARPS, L.nBlocks), R_WARPS * 32, s, L.bt, L.stIn, L.stOut,
            maxTiles, occ, mode);
    *launches += 4;
}

void launch_sbrt_inverse(const StageLaunch& L, int mode, Workspace& ws, cudaStream_t s, u64* launches)
{
    (void)ws;
    KLAUNCH(sbrt_decide_kernel, (L.nBlocks + 31) / 32, 32, s, L, 1);
    if (L.maxLen < (1 << 24))
        KLAUNCH(sbrt_inverse_kernel<u32>, L.nBlocks, 32, s, L.bt, L.stIn, L.stOut, mode);
    else
        KLAUNCH(sbrt_inverse_kernel<u64>, L.nBlocks, 32, s, L.bt, L.stIn, L.stOut, mode);
    *launches += 2;
}
