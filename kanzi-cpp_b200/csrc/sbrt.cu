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

// Inverse: serial replay per block (SBRT.cpp:99-145), one thread per block.
__global__ void __launch_bounds__(32)
sbrt_inverse_kernel(BufTable bt, const BlkState* __restrict__ stIn, const BlkState* __restrict__ stOut, int mode)
{
    __shared__ int s_q[256], s_p[256];
    __shared__ u8 s_r2s[256];
    const int b = blockIdx.x;
    const BlkState bs = stIn[b];
    if (stOut[b].swaps == bs.swaps)
        return;
    const int n = bs.len;
    const u8* __restrict__ src = blk_src(bt, bs, b);
    u8* __restrict__ dst = blk_dst(bt, bs, b);
    for (int i = threadIdx.x; i < 256; i += 32) {
        s_q[i] = 0;
        s_p[i] = 0;
        s_r2s[i] = (u8)i;
    }
    __syncwarp();
    if (threadIdx.x != 0)
        return;
    u32 m1, m2;
    int sh;
    sbrt_masks(mode, m1, m2, sh);
    int i = 0;
    // 4 bytes at a time when aligned
    while (i < n) {
        u32 wv;
        int cnt;
        if (i + 4 <= n) {
            wv = *reinterpret_cast<const u32*>(src + i); // buffers are 256-byte aligned, i % 4 == 0
            cnt = 4;
        } else {
            wv = 0;
            cnt = n - i;
            for (int k = 0; k < cnt; k++)
                wv |= (u32)src[i + k] << (8 * k);
        }
        u32 ow = 0;
        for (int k = 0; k < cnt; k++, i++) {
            int r = (int)((wv >> (8 * k)) & 0xFF);
            const int c = s_r2s[r];
            ow |= (u32)c << (8 * k);
            const int qc = (int)((((u32)i & m1) + ((u32)s_p[c] & m2)) >> sh);
            s_p[c] = i;
            s_q[c] = qc;
            while (r > 0) {
                const int above = s_r2s[r - 1];
                if (s_q[above] > qc)
                    break;
                s_r2s[r] = (u8)above;
                r--;
            }
            s_r2s[r] = (u8)c;
        }
        if (cnt == 4) {
            *reinterpret_cast<u32*>(dst + i - 4) = ow;
        } else {
            for (int k = 0; k < cnt; k++)
                dst[i - cnt + k] = (u8)(ow >> (8 * k));
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
    KLAUNCH(sbrt_inverse_kernel, L.nBlocks, 32, s, L.bt, L.stIn, L.stOut, mode);
    *launches += 2;
}
