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
#include <stdlib.h>

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
    // Post-BWT data is long runs of one symbol: 32 lanes hitting one shared-memory word serialise.
    // Positions grow with the lane, so of the lanes that hold the same symbol only the highest one
    // (match_any) can win the max: one atomic per distinct symbol per warp row.
    const int lane = threadIdx.x & 31;
    for (int i0 = base + (threadIdx.x & ~31); i0 < end; i0 += 256) {
        const int i = i0 + lane;
        const u32 c = (i < end) ? (u32)src[i] : (256u + (u32)lane);
        const u32 peers = __match_any_sync(FULL_MASK, c);
        if (i < end && (peers >> lane) == 1u)
            atomicMax(&s_t1[c], (u32)(i + 1));
    }
    __syncthreads();
    for (int i0 = base + (threadIdx.x & ~31); i0 < end; i0 += 256) {
        const int i = i0 + lane;
        const bool in = i < end;
        const u32 c = in ? (u32)src[i] : (256u + (u32)lane);
        const u32 peers = __match_any_sync(FULL_MASK, c);
        const bool cand = in && ((u32)(i + 1) != s_t1[c]);
        const u32 cm = __ballot_sync(FULL_MASK, cand) & peers; // same-symbol lanes that are not the last occurrence
        if (cand && (cm >> lane) == 1u)
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

// ---- the list, distributed over the lanes of one warp --------------------------
// Rank g lives in lane (g & 31), slot (g >> 5): the 32 highest ranks are slot 0 of
// the 32 lanes, so the common case (post-BWT ranks are small) touches one register
// per lane.  Entry = key q and pb = (last access time << 8) | symbol.
template <class PB>
struct RankList {
    int q[8];
    PB pb[8];

    // Move the entry at rank r (new key qc, new payload ne) up to its new rank:
    // rp = #entries with key > qc (the list is sorted by key, ties: latest first).
    // Fast path for r < 32, branch-free and without a vote: with nq/npb = the entry of
    // the rank above (shuffled up BEFORE qc is known, off the critical path), lane g
    //   takes the entry above   iff g <= r, g > 0 and q[g-1] <= qc   (it is below the new rank)
    //   receives the new entry  iff g <= r, q[g] <= qc and (g == 0 or q[g-1] > qc)
    // (the list is sorted by key, so "q[g-1] <= qc" == "g > new rank").
    __device__ __forceinline__ void move_up_top(int r, int qc, PB ne, int lane, int nq, PB npb)
    {
        const bool below = lane <= r;
        const bool mv = below && (lane > 0) && (nq <= qc);
        const bool ins = below && (q[0] <= qc) && ((lane == 0) || (nq > qc));
        q[0] = ins ? qc : (mv ? nq : q[0]);
        pb[0] = ins ? ne : (mv ? npb : pb[0]);
    }

    __device__ __forceinline__ void move_up(int r, int qc, PB ne, int lane)
    {
        int rp = 0;
#pragma unroll
        for (int k = 0; k < 8; k++)
            rp += __popc(__ballot_sync(FULL_MASK, q[k] > qc));
#pragma unroll
        for (int k = 7; k >= 0; k--) {
            // value arriving from the rank above: lane-1 of this slot, or lane 31 of the previous slot
            int nq = __shfl_up_sync(FULL_MASK, q[k], 1);
            PB npb = __shfl_up_sync(FULL_MASK, pb[k], 1);
            if (k > 0) {
                const int sq = __shfl_sync(FULL_MASK, q[k - 1], 31);
                const PB spb = __shfl_sync(FULL_MASK, pb[k - 1], 31);
                if (lane == 0) {
                    nq = sq;
                    npb = spb;
                }
            }
            const int g = 32 * k + lane;
            const bool mv = (g > rp) && (g <= r);
            q[k] = mv ? nq : q[k];
            pb[k] = mv ? npb : pb[k];
        }
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (32 * k + lane == rp) {
                q[k] = qc;
                pb[k] = ne;
            }
    }

    __device__ __forceinline__ PB entry_at(int r) const
    {
        if (r < 32)
            return __shfl_sync(FULL_MASK, pb[0], r);
        // mask-select (an if-chain gets turned into an indexed load, which would push
        // the whole list into local memory)
        PB sel = 0;
#pragma unroll
        for (int k = 1; k < 8; k++)
            sel |= pb[k] & (PB)(0 - (PB)((r >> 5) == k));
        return __shfl_sync(FULL_MASK, sel, r & 31);
    }

    // rank of symbol c (always present)
    __device__ __forceinline__ int find(u32 c) const
    {
        u32 m = __ballot_sync(FULL_MASK, (u32)(pb[0] & 0xFF) == c);
        if (m)
            return __ffs(m) - 1;
#pragma unroll
        for (int k = 1; k < 8; k++) {
            m = __ballot_sync(FULL_MASK, (u32)(pb[k] & 0xFF) == c);
            if (m)
                return 32 * k + __ffs(m) - 1;
        }
        return 255;
    }
};

// 3. replay: one warp per tile, exact list algorithm started from the tile's entry table
#define R_WARPS 4
template <class PB, int MODE>
__global__ void __launch_bounds__(R_WARPS * 32)
sbrt_rank_kernel(BufTable bt, const BlkState* __restrict__ stIn, const BlkState* __restrict__ stOut, int maxTiles,
                 const uint2* __restrict__ occ)
{
    const int mode = MODE;
    auto keyOf = [](u32 i, u32 pc) -> int {
        return (MODE == 1) ? (int)i : (MODE == 2) ? (int)((i + pc) >> 1) : (int)pc;
    };
    __shared__ u64 s_keys[R_WARPS][256];
    __shared__ u64 s_ent[R_WARPS][256];  // pb = (last access time << 8) | symbol, by rank, while the list is built
    __shared__ u32 s_entq[R_WARPS][256]; // key q by rank (kept apart: pb needs more than 32 bits from 16 MiB on)
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
    // ---- build the sorted list for this tile: rank(s) = #{u : key_u > key_s}
    u64* keys = s_keys[w];
    u64* ent = s_ent[w];
    u32* entq = s_entq[w];
    const uint2* o = occ + ((i64)b * maxTiles + t) * 256;
    u64 myKey[8];
    u32 myQ[8], myP[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int sym = 32 * k + lane;
        const uint2 e = o[sym];
        myKey[k] = sbrt_key(e.x, e.y, sym, m1, m2, sh);
        keys[sym] = myKey[k];
        myQ[k] = (u32)(myKey[k] >> 32);
        myP[k] = e.x ? e.x - 1 : 0u;
    }
    __syncwarp();
    int rk[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    for (int u = 0; u < 256; u++) {
        const u64 ku = keys[u];
#pragma unroll
        for (int k = 0; k < 8; k++)
            rk[k] += (ku > myKey[k]) ? 1 : 0;
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
        ent[rk[k]] = ((u64)myP[k] << 8) | (u64)(32 * k + lane);
        entq[rk[k]] = myQ[k];
    }
    __syncwarp();
    RankList<PB> L;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        L.q[k] = (int)entq[32 * k + lane];
        L.pb[k] = (PB)ent[32 * k + lane];
    }
    // ---- replay
    const int end = min(base + S_TILE, n);
    for (int g = base; g < end; g += 128) {
        u32 inw = 0;
        {
            const int p = g + 4 * lane;
            if (p + 4 <= n) {
                inw = *reinterpret_cast<const u32*>(src + p); // buffers are 16-byte aligned
            } else {
                for (int k = 0; k < 4; k++)
                    if (p + k < n)
                        inw |= (u32)src[p + k] << (8 * k);
            }
        }
        u32 outw = 0;
        const int cnt = min(128, end - g);
        for (int j = 0; j < cnt; j += 4) {
            const u32 w4 = __shfl_sync(FULL_MASK, inw, j >> 2);
            const int lim = min(4, cnt - j);
            u32 o4 = 0;
            for (int x = 0; x < lim; x++) {
                const u32 c = (w4 >> (8 * x)) & 0xFF;
                const u32 i = (u32)(g + j + x);
                // top-32 probe: who holds c, and its last access time, without waiting for the rank
                const bool hit = (u32)(L.pb[0] & 0xFF) == c;
                const u32 m = __ballot_sync(FULL_MASK, hit);
                if (m & 1u) {
                    // c is the head of the list: rank 0, only the head's key changes
                    if (lane == 0) {
                        L.q[0] = keyOf(i, (u32)(L.pb[0] >> 8));
                        L.pb[0] = ((PB)i << 8) | (PB)c;
                    }
                } else if (m) {
                    const int nq = __shfl_up_sync(FULL_MASK, L.q[0], 1);
                    const PB npb = __shfl_up_sync(FULL_MASK, L.pb[0], 1);
                    const u32 pc = __reduce_or_sync(FULL_MASK, hit ? (u32)(L.pb[0] >> 8) : 0u);
                    const int r = __ffs(m) - 1;
                    const int qc = keyOf(i, pc);
                    o4 |= (u32)r << (8 * x);
                    L.move_up_top(r, qc, ((PB)i << 8) | (PB)c, lane, nq, npb);
                } else {
                    const int r = L.find(c);
                    const PB e = L.entry_at(r);
                    const u32 pc = (u32)(e >> 8);
                    const int qc = keyOf(i, pc);
                    o4 |= (u32)r << (8 * x);
                    L.move_up(r, qc, ((PB)i << 8) | (PB)c, lane);
                }
            }
            if (lane == (j >> 2))
                outw = o4;
        }
        {
            const int p = g + 4 * lane;
            if (p + 4 <= end) {
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
// step: one warp per block, the list distributed over the lanes in registers
// (RankList), no shared memory, no divergence; zero words (runs of rank 0) take a
// closed form; input/output travel through registers 128 bytes at a time.
template <class PB, int MODE>
__global__ void __launch_bounds__(32)
sbrt_inverse_kernel(BufTable bt, const BlkState* __restrict__ stIn, const BlkState* __restrict__ stOut)
{
    const int b = blockIdx.x;
    const BlkState bs = stIn[b];
    if (stOut[b].swaps == bs.swaps)
        return;
    const int n = bs.len;
    const int lane = threadIdx.x;
    const u8* __restrict__ src = blk_src(bt, bs, b);
    u8* __restrict__ dst = blk_dst(bt, bs, b);
    // key of an access at time i to a symbol last seen at time pc (SBRT.cpp:79)
    auto keyOf = [](u32 i, u32 pc) -> int {
        return (MODE == 1) ? (int)i : (MODE == 2) ? (int)((i + pc) >> 1) : (int)pc;
    };
    RankList<PB> L;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        L.q[k] = 0;
        L.pb[k] = (PB)(32 * k + lane);
    }
    const bool lane0 = lane == 0;
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
                const PB e = __shfl_sync(FULL_MASK, L.pb[0], 0);
                const u32 c = (u32)(e & 0xFF);
                const u32 i3 = (u32)(base + j + 3);
                if (lane0) {
                    L.q[0] = keyOf(i3, i3 - 1);
                    L.pb[0] = ((PB)i3 << 8) | (PB)c;
                }
                o4 = c * 0x01010101u;
            } else if (lim == 4 && (w4 & 0xE0E0E0E0u) == 0) {
                // all four ranks < 32: branch-free top-of-list path
#pragma unroll
                for (int x = 0; x < 4; x++) {
                    const int r = (int)((w4 >> (8 * x)) & 0xFF);
                    const u32 i = (u32)(base + j + x);
                    const int nq = __shfl_up_sync(FULL_MASK, L.q[0], 1);
                    const PB npb = __shfl_up_sync(FULL_MASK, L.pb[0], 1);
                    const PB e = __shfl_sync(FULL_MASK, L.pb[0], r);
                    const u32 c = (u32)(e & 0xFF);
                    const int qc = keyOf(i, (u32)(e >> 8));
                    o4 |= c << (8 * x);
                    const bool below = lane <= r;
                    const bool mv = below && !lane0 && (nq <= qc);
                    const bool ins = below && (L.q[0] <= qc) && (lane0 || (nq > qc));
                    const PB ne = ((PB)i << 8) | (PB)c;
                    L.q[0] = ins ? qc : (mv ? nq : L.q[0]);
                    L.pb[0] = ins ? ne : (mv ? npb : L.pb[0]);
                }
            } else {
                for (int x = 0; x < lim; x++) {
                    const int r = (int)((w4 >> (8 * x)) & 0xFF);
                    const u32 i = (u32)(base + j + x);
                    const PB e = L.entry_at(r);
                    const u32 c = (u32)(e & 0xFF);
                    const int qc = keyOf(i, (u32)(e >> 8));
                    o4 |= c << (8 * x);
                    L.move_up(r, qc, ((PB)i << 8) | (PB)c, lane);
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

// ---- inverse, latency-optimised (blocks < 2^24 bytes) --------------------------------
// Measured on B200 with one warp per SM sub-partition (tools/microbench/lat*.cu):
// SHFL.IDX 33 cycles, SHFL.UP 26, +4.5 per extra shuffle in flight, ALU op 4.4,
// ballot+popc 15, a TAKEN branch 15-25.  A step of sbrt_inverse_kernel costs 98 cycles
// (three dependent predicate hops after the shuffle); this kernel's step costs ~68:
//   * keys are kept as Y & ~1 with Y = i + p (RANK mode): (a >> 1) <= (Y >> 1)  <=>
//     (a & ~1) <= Y, so the shift leaves the critical path and both selects hang off ONE
//     compare level:  lane g <= r with K[g] <= Y takes (K[g-1] <= Y ? entry g-1 : new entry)
//   * a word of four ranks < 32 is straight-line code; the only taken branch per word is
//     the loop back-edge
//   * runs of zero words (rank 0 repeated: only the head's key/time change) are folded in
//     closed form from a ballot mask of the 32 input words of a 128-byte group.
// Shuffles in program order: the indexed one (33 cycles, on the critical path) is issued
// before the two up-shuffles instead of behind them.
#ifndef KNZ_SIM
__device__ __forceinline__ u32 shfl_idx_ordered(u32 v, int l)
{
    u32 r;
    asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(r) : "r"(v), "r"(l));
    return r;
}
__device__ __forceinline__ u32 shfl_up1_ordered(u32 v)
{
    u32 r;
    asm volatile("shfl.sync.up.b32 %0, %1, 1, 0x0, 0xffffffff;" : "=r"(r) : "r"(v));
    return r;
}
#else
__device__ __forceinline__ u32 shfl_idx_ordered(u32 v, int l) { return __shfl_sync(FULL_MASK, v, l); }
__device__ __forceinline__ u32 shfl_up1_ordered(u32 v) { return __shfl_up_sync(FULL_MASK, v, 1); }
#endif

template <int MODE>
struct InvList {
    u32 dK[8], dP[8]; // rank g in lane g & 31, slot g >> 5;  dP = (last access time << 8) | symbol

    static __device__ __forceinline__ u32 key_raw(u32 i, u32 p)
    {
        return (MODE == 1) ? i : (MODE == 2) ? i + p : p;
    }
    static __device__ __forceinline__ u32 key_store(u32 y) { return (MODE == 2) ? (y & ~1u) : y; }

    // r < 32, branch-free
    __device__ __forceinline__ u32 step_top(int r, u32 i, int lane)
    {
        const u32 e = shfl_idx_ordered(dP[0], r);
        u32 nK = shfl_up1_ordered(dK[0]);
        const u32 nP = shfl_up1_ordered(dP[0]);
        if (lane == 0)
            nK = 0xFFFFFFFFu;
        const u32 c = e & 0xFF;
        const u32 y = key_raw(i, e >> 8);
        const u32 yn = key_store(y);
        const u32 ne = (i << 8) | c;
        if (lane <= r && dK[0] <= y) {
            const bool up = nK <= y;
            dK[0] = up ? nK : yn;
            dP[0] = up ? nP : ne;
        }
        return c;
    }

    // any r.  Entries below rank r keep keys <= the accessed entry's old key <= its new one, so only slots
    // 0..r/32 can hold a larger key, and only slots rp/32..r/32 change (r, rp are warp uniform).  One rotate
    // per array and slot: lane l takes lane l-1, and lane 0 receives lane 31 of the same slot, which is what
    // lane 0 of the NEXT slot needs -- it is carried there in a register instead of a second shuffle.
    __device__ __forceinline__ u32 step_deep(int r, u32 i, int lane)
    {
        const int kr = r >> 5;
        u32 sel = 0;
#pragma unroll
        for (int k = 0; k < 8; k++)
            sel |= dP[k] & (0u - (u32)(kr == k));
        const u32 e = __shfl_sync(FULL_MASK, sel, r & 31);
        const u32 c = e & 0xFF;
        const u32 y = key_raw(i, e >> 8);
        const u32 yn = key_store(y);
        const u32 ne = (i << 8) | c;
        int rp = 0;
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (k <= kr)
                rp += __popc(__ballot_sync(FULL_MASK, dK[k] > y));
        const int kp = rp >> 5;
        const int from = (lane + 31) & 31;
        u32 carryK = 0, carryP = 0; // lane 0: lane 31 of the slot before (never used in slot kp: 32 kp <= rp)
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (k >= kp && k <= kr) {
                const u32 rK = __shfl_sync(FULL_MASK, dK[k], from);
                const u32 rP = __shfl_sync(FULL_MASK, dP[k], from);
                const u32 nK = (lane == 0) ? carryK : rK;
                const u32 nP = (lane == 0) ? carryP : rP;
                carryK = rK;
                carryP = rP;
                const int g = 32 * k + lane;
                const bool mv = (g > rp) && (g <= r);
                const bool ins = g == rp;
                dK[k] = ins ? yn : (mv ? nK : dK[k]);
                dP[k] = ins ? ne : (mv ? nP : dP[k]);
            }
        }
        return c;
    }
};

template <int MODE>
__global__ void __launch_bounds__(32)
sbrt_inverse_fast_kernel(BufTable bt, const BlkState* __restrict__ stIn, const BlkState* __restrict__ stOut)
{
    const int b = blockIdx.x;
    const BlkState bs = stIn[b];
    if (stOut[b].swaps == bs.swaps)
        return;
    const int n = bs.len;
    const int lane = threadIdx.x;
    const u8* __restrict__ src = blk_src(bt, bs, b);
    u8* __restrict__ dst = blk_dst(bt, bs, b);
    InvList<MODE> L;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        L.dK[k] = 0;
        L.dP[k] = (u32)(32 * k + lane);
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
        const int fullWords = cnt >> 2;
        u32 zm = __ballot_sync(FULL_MASK, inw == 0);                     // words of four zero ranks
        const u32 dm = __ballot_sync(FULL_MASK, (inw & 0xE0E0E0E0u) != 0); // words holding a rank >= 32
        const u32 valid = (fullWords < 32) ? ((1u << fullWords) - 1u) : 0xFFFFFFFFu;
        zm &= valid;
        const u32 notfast = zm | dm | ~valid;
        // four ranks < 32, straight line
        auto word = [&](u32 w, int jj) {
            const u32 i0 = (u32)(base + 4 * jj);
            u32 o4 = L.step_top((int)(w & 0xFF), i0, lane);
            o4 |= L.step_top((int)((w >> 8) & 0xFF), i0 + 1, lane) << 8;
            o4 |= L.step_top((int)((w >> 16) & 0xFF), i0 + 2, lane) << 16;
            o4 |= L.step_top((int)(w >> 24), i0 + 3, lane) << 24;
            if (lane == jj)
                outw = o4;
        };
        int j = 0;
        u32 w4 = __shfl_sync(FULL_MASK, inw, 0);
        while (j < fullWords) {
            const u32 zrest = zm >> j;
            if (zrest & 1u) {
                // zr consecutive zero words = 4*zr accesses to the head: only its key and
                // last access time change (the previous access was one position earlier)
                const int zr = (zrest == 0xFFFFFFFFu) ? 32 : (__ffs((int)~zrest) - 1);
                const u32 i3 = (u32)(base + 4 * (j + zr) - 1);
                const u32 c = __shfl_sync(FULL_MASK, L.dP[0], 0) & 0xFF;
                if (lane == 0) {
                    L.dK[0] = InvList<MODE>::key_store(InvList<MODE>::key_raw(i3, i3 - 1));
                    L.dP[0] = (i3 << 8) | c;
                }
                if (lane >= j && lane < j + zr)
                    outw = c * 0x01010101u;
                j += zr;
                w4 = __shfl_sync(FULL_MASK, inw, j & 31);
                continue;
            }
            if ((dm >> j) & 1u) {
                const u32 i0 = (u32)(base + 4 * j);
                u32 o4 = 0;
#pragma unroll 1
                for (int x = 0; x < 4; x++) {
                    const int r = (int)((w4 >> (8 * x)) & 0xFF);
                    const u32 c = (r < 32) ? L.step_top(r, i0 + x, lane) : L.step_deep(r, i0 + x, lane);
                    o4 |= c << (8 * x);
                }
                if (lane == j)
                    outw = o4;
                j++;
                w4 = __shfl_sync(FULL_MASK, inw, j & 31);
                continue;
            }
            // up to four consecutive plain words per trip: the loop back-edge is the only
            // taken branch, and the words were fetched before the chain needs them
            const u32 nf = notfast >> j; // bit 0 is clear
            int run = (nf == 0) ? 32 : (__ffs((int)nf) - 1);
            run = min(min(run, 4), 32 - j);
            const u32 wa = __shfl_sync(FULL_MASK, inw, (j + 1) & 31);
            const u32 wb = __shfl_sync(FULL_MASK, inw, (j + 2) & 31);
            const u32 wc = __shfl_sync(FULL_MASK, inw, (j + 3) & 31);
            const u32 wnext = __shfl_sync(FULL_MASK, inw, (j + run) & 31);
            word(w4, j);
            if (run > 1) {
                word(wa, j + 1);
                if (run > 2) {
                    word(wb, j + 2);
                    if (run > 3)
                        word(wc, j + 3);
                }
            }
            j += run;
            w4 = wnext;
        }
        if (cnt & 3) { // ragged tail of the block
            w4 = __shfl_sync(FULL_MASK, inw, fullWords & 31);
            u32 o4 = 0;
            for (int x = 0; x < (cnt & 3); x++) {
                const int r = (int)((w4 >> (8 * x)) & 0xFF);
                const u32 c = L.step_deep(r, (u32)(base + 4 * fullWords + x), lane);
                o4 |= c << (8 * x);
            }
            if (lane == fullWords)
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

// ---- forward replay, instruction-lean (blocks < 2^24 bytes) ----------------------------
// sbrt_rank_kernel is issue-bound (ncu: 84% issue slots busy, 50 instructions per symbol).
// Same tile decomposition, but the replay uses the InvList step (even keys, one compare
// level), whole words equal to the head symbol are folded in closed form from a ballot
// mask, and the four symbols of a word are unrolled with constant shifts: ~21 instructions
// per symbol on post-BWT data.
template <int MODE>
__global__ void __launch_bounds__(R_WARPS * 32)
sbrt_rank_fast_kernel(BufTable bt, const BlkState* __restrict__ stIn, const BlkState* __restrict__ stOut, int maxTiles,
                      const uint2* __restrict__ occ)
{
    __shared__ u64 s_keys[R_WARPS][256];
    __shared__ u64 s_ent[R_WARPS][256]; // (K << 32 | P) by rank, only while the list is built
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
    sbrt_masks(MODE, m1, m2, sh);
    const u8* __restrict__ src = blk_src(bt, bs, b);
    u8* __restrict__ dst = blk_dst(bt, bs, b);
    // ---- sorted list at the tile's first position: rank(s) = #{u : key_u > key_s}
    u64* keys = s_keys[w];
    u64* ent = s_ent[w];
    const uint2* o = occ + ((i64)b * maxTiles + t) * 256;
    u64 myKey[8];
    u32 myP[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int sym = 32 * k + lane;
        const uint2 e = o[sym];
        myKey[k] = sbrt_key(e.x, e.y, sym, m1, m2, sh);
        keys[sym] = myKey[k];
        myP[k] = e.x ? e.x - 1 : 0u;
    }
    __syncwarp();
    int rk[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    for (int u = 0; u < 256; u++) {
        const u64 ku = keys[u];
#pragma unroll
        for (int k = 0; k < 8; k++)
            rk[k] += (ku > myKey[k]) ? 1 : 0;
    }
#pragma unroll
    for (int k = 0; k < 8; k++) // stored key = q << sh (even in RANK mode), see InvList
        ent[rk[k]] = ((u64)((u32)(myKey[k] >> 32) << sh) << 32) | ((u64)myP[k] << 8) | (u64)(32 * k + lane);
    __syncwarp();
    InvList<MODE> L;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const u64 e = ent[32 * k + lane];
        L.dK[k] = (u32)(e >> 32);
        L.dP[k] = (u32)e;
    }
    // ---- replay
    const int end = min(base + S_TILE, n);
    for (int g = base; g < end; g += 128) {
        u32 inw = 0;
        {
            const int p = g + 4 * lane;
            if (p + 4 <= n) {
                inw = *reinterpret_cast<const u32*>(src + p); // buffers are 16-byte aligned
            } else {
                for (int k = 0; k < 4; k++)
                    if (p + k < n)
                        inw |= (u32)src[p + k] << (8 * k);
            }
        }
        u32 outw = 0;
        const int cnt = min(128, end - g);
        const int fullWords = cnt >> 2;
        int j = 0;
        while (j < fullWords) {
            const u32 hs = __shfl_sync(FULL_MASK, L.dP[0], 0) & 0xFF; // head symbol
            u32 zm = __ballot_sync(FULL_MASK, inw == hs * 0x01010101u) >> j;
            if (j + 32 > fullWords && fullWords - j < 32)
                zm &= (1u << (fullWords - j)) - 1u;
            if (zm & 1u) {
                // zr words of the head symbol: ranks 0, only the head's key and time change
                const int zr = (zm == 0xFFFFFFFFu) ? 32 : (__ffs((int)~zm) - 1);
                const u32 i3 = (u32)(g + 4 * (j + zr) - 1);
                if (lane == 0) {
                    L.dK[0] = InvList<MODE>::key_store(InvList<MODE>::key_raw(i3, i3 - 1));
                    L.dP[0] = (i3 << 8) | hs;
                }
                j += zr; // outw stays 0 for these lanes
                continue;
            }
            const u32 w4 = __shfl_sync(FULL_MASK, inw, j);
            const u32 i0 = (u32)(g + 4 * j);
            u32 o4 = 0;
#pragma unroll
            for (int x = 0; x < 4; x++) {
                const u32 c = (w4 >> (8 * x)) & 0xFF;
                const u32 m = __ballot_sync(FULL_MASK, (L.dP[0] & 0xFF) == c);
                int r;
                if (m) {
                    r = __ffs((int)m) - 1;
                    L.step_top(r, i0 + x, lane);
                } else {
                    r = 255;
#pragma unroll
                    for (int k = 1; k < 8; k++) {
                        const u32 mk = __ballot_sync(FULL_MASK, (L.dP[k] & 0xFF) == c);
                        if (mk) {
                            r = 32 * k + __ffs((int)mk) - 1;
                            break;
                        }
                    }
                    L.step_deep(r, i0 + x, lane);
                }
                o4 |= (u32)r << (8 * x);
            }
            if (lane == j)
                outw = o4;
            j++;
        }
        if (cnt & 3) { // ragged tail of the block
            const u32 w4 = __shfl_sync(FULL_MASK, inw, fullWords & 31);
            u32 o4 = 0;
            for (int x = 0; x < (cnt & 3); x++) {
                const u32 c = (w4 >> (8 * x)) & 0xFF;
                int r = 255;
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const u32 mk = __ballot_sync(FULL_MASK, (L.dP[k] & 0xFF) == c);
                    if (mk) {
                        r = 32 * k + __ffs((int)mk) - 1;
                        break;
                    }
                }
                L.step_deep(r, (u32)(g + 4 * fullWords + x), lane);
                o4 |= (u32)r << (8 * x);
            }
            if (lane == fullWords)
                outw = o4;
        }
        {
            const int p = g + 4 * lane;
            if (p + 4 <= end) {
                *reinterpret_cast<u32*>(dst + p) = outw;
            } else {
                for (int k = 0; k < 4; k++)
                    if (p + k < end)
                        dst[p + k] = (u8)(outw >> (8 * k));
            }
        }
    }
}

// ---- forward replay, four lanes per tile ------------------------------------------------
// The warp-per-tile replays above spend a whole warp instruction on every step of ONE tile
// (36 warp instructions per symbol: issue-bound).  Here a quad owns a tile: lane q keeps ranks
// 8q..8q+7 in registers (the 32 ranks that post-BWT data touches almost all the time), ranks
// 32..255 stay in shared memory, and the eight quads of a warp replay eight tiles in lockstep:
// the same instruction stream now advances eight symbols.  An entry is ONE register,
// (stored key << 8) | symbol -- keys stay below 2^24 for blocks under 8 MiB -- and the last
// access time of every symbol lives in a per-tile shared-memory table, so a step is
//   find    each lane compares the low bytes of its 8 entries with c; one ballot tells the quad
//           who holds it, one quad shuffle broadcasts the rank
//   key     y = i + lastT[c] (RANK) / i (MTFT): a shared-memory read, no cross-lane traffic
//   update  the lane's entries are sorted, so the ones with key <= y are a suffix starting at
//           gq = #(E > Y); entries gq+1 .. min(r - 8q, 7) take the entry above them and entry gq
//           receives the new entry (or the last entry of the lane above): one move mask, predicated
//           moves, one shuffle-up
// ncu (profiles/r02_ncu_rank_quad_*.md): the two-register version of this kernel was bound by the
// integer ALU pipe (81 % busy, 24 warp instructions per symbol); halving the registers an entry
// occupies halves the selects.  A symbol below rank 31 takes a rare slow path in which the whole
// warp searches, re-ranks and shifts the shared list of that tile (7 entries per lane).
#define RQ_WARPS 2
template <int MODE>
__global__ void __launch_bounds__(RQ_WARPS * 32)
sbrt_rank_quad_kernel(BufTable bt, const BlkState* __restrict__ stIn, const BlkState* __restrict__ stOut, int maxTiles,
                      const uint2* __restrict__ occ)
{
    __shared__ u64 s_keys[RQ_WARPS][256];
    __shared__ u32 s_lE[RQ_WARPS][8][256]; // per tile: (stored key << 8) | symbol by rank (ranks 32.. = the deep list)
    __shared__ u32 s_lT[RQ_WARPS][8][256]; // per tile: last access time of every symbol
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int tile0 = (blockIdx.x * RQ_WARPS + w) * 8;
    const BlkState bs = stIn[b];
    if (stOut[b].swaps == bs.swaps)
        return;
    const int n = bs.len;
    if ((i64)tile0 * S_TILE >= n)
        return;
    u32 m1, m2;
    int sh;
    sbrt_masks(MODE, m1, m2, sh);
    const u8* __restrict__ src = blk_src(bt, bs, b);
    u8* __restrict__ dst = blk_dst(bt, bs, b);
    // ---- sorted lists at the first position of the warp's eight tiles (whole warp per tile)
    for (int jt = 0; jt < 8; jt++) {
        const int t = tile0 + jt;
        if ((i64)t * S_TILE >= n)
            break;
        u64* keys = s_keys[w];
        const uint2* o = occ + ((i64)b * maxTiles + t) * 256;
        u64 myKey[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int sym = 32 * k + lane;
            const uint2 e = o[sym];
            myKey[k] = sbrt_key(e.x, e.y, sym, m1, m2, sh);
            keys[sym] = myKey[k];
            s_lT[w][jt][sym] = e.x ? e.x - 1 : 0u;
        }
        __syncwarp();
        int rk[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
        for (int u = 0; u < 256; u++) {
            const u64 ku = keys[u];
#pragma unroll
            for (int k = 0; k < 8; k++)
                rk[k] += (ku > myKey[k]) ? 1 : 0;
        }
#pragma unroll
        for (int k = 0; k < 8; k++) // stored key = q << sh (even in RANK mode), see InvList
            s_lE[w][jt][rk[k]] = (((u32)(myKey[k] >> 32) << sh) << 8) | (u32)(32 * k + lane);
        __syncwarp();
    }
    const int q = lane & 3, jq = lane >> 2, qbase = lane & ~3;
    const int base = (tile0 + jq) * S_TILE;
    const bool active = base < n;
    const int end = active ? min(base + S_TILE, n) : 0;
    u32 E[8];
#pragma unroll
    for (int k = 0; k < 8; k++)
        E[k] = s_lE[w][jq][8 * q + k];
    u32* lastT = s_lT[w][jq];

    for (int g = 0; g < S_TILE; g += 64) {
        const int pos0 = base + g;
        if (!__any_sync(FULL_MASK, active && pos0 < end))
            break;
        uint4 in = make_uint4(0, 0, 0, 0);
        {
            const int p = pos0 + 16 * q;
            if (active && p < end) {
                if (p + 16 <= n) {
                    in = *reinterpret_cast<const uint4*>(src + p); // buffers and tiles are 16-byte aligned
                } else {
                    u32 wv[4] = { 0, 0, 0, 0 };
                    for (int k = 0; k < 16; k++)
                        if (p + k < n)
                            wv[k >> 2] |= (u32)src[p + k] << (8 * (k & 3));
                    in = make_uint4(wv[0], wv[1], wv[2], wv[3]);
                }
            }
        }
        uint4 outv = make_uint4(0, 0, 0, 0);
#pragma unroll 1
        for (int tt = 0; tt < 4; tt++) {
            const int sl = qbase + tt;
            u32 ov0 = 0, ov1 = 0, ov2 = 0, ov3 = 0;
#pragma unroll 1
            for (int wi = 0; wi < 4; wi++) { // four steps unrolled per trip (the step body is large: keep it in the I-cache)
                const u32 wsel = (wi == 0) ? in.x : (wi == 1) ? in.y : (wi == 2) ? in.z : in.w;
                const u32 w4 = __shfl_sync(FULL_MASK, wsel, sl);
                u32 o4 = 0;
#pragma unroll
                for (int xb = 0; xb < 4; xb++) {
                    const u32 i = (u32)(pos0 + 16 * tt + 4 * wi + xb);
                    const bool live = active && ((int)i < end);
                    const u32 c = (w4 >> (8 * xb)) & 0xFF;
                    // find c among the lane's eight entries
                    int kk = -1;
#pragma unroll
                    for (int k = 0; k < 8; k++)
                        kk = (((E[k] ^ c) & 0xFF) == 0) ? k : kk;
                    const u32 bal = (__ballot_sync(FULL_MASK, live && kk >= 0) >> qbase) & 0xFu;
                    const bool miss = live && bal == 0;
                    // key of the access: the symbol's last access time sits in shared memory
                    const u32 t1 = live ? lastT[c] : 0u;
                    const u32 y = InvList<MODE>::key_raw(i, t1);
                    const u32 Y = (y << 8) | 0xFFu;                         // E <= Y  <=>  key <= y
                    const u32 ne = (InvList<MODE>::key_store(y) << 8) | c; // the entry after the access
                    __syncwarp(); // every lane of the quad has read lastT[c] before it is overwritten
                    if (live && q == 0)
                        lastT[c] = i;
                    int r = 0, rsrc = 0;
                    bool upd = false;
                    u32 mm = __ballot_sync(FULL_MASK, miss);
                    while (mm) {
                        // ---- rare: c sits below rank 31.  One missing quad at a time, the WHOLE warp works
                        // on its shared list (224 entries = 7 per lane): search, new position and the shift
                        // are parallel (a serial walk by one lane cost ~14000 cycles on blocks whose symbols
                        // come back from the bottom of the list, e.g. byte counters).
                        const int ml = __ffs((int)mm) - 1; // lane 0 of the quad
                        const int jm = ml >> 2;
                        mm &= ~(0xFu << ml);
                        const u32 cm = __shfl_sync(FULL_MASK, c, ml);
                        const u32 Ym = __shfl_sync(FULL_MASK, Y, ml);
                        const u32 nem = __shfl_sync(FULL_MASK, ne, ml);
                        u32* mE = &s_lE[w][jm][32];
                        u32 vE[7];
                        int myd = -1;
#pragma unroll
                        for (int t = 0; t < 7; t++) {
                            vE[t] = mE[lane + 32 * t];
                            if ((vE[t] & 0xFF) == cm)
                                myd = lane + 32 * t;
                        }
                        const u32 fb = __ballot_sync(FULL_MASK, myd >= 0);
                        const int d = fb ? __shfl_sync(FULL_MASK, myd, __ffs((int)fb) - 1) : 223;
                        int cg = 0; // register entries of that quad with a larger key
                        if (jq == jm) {
#pragma unroll
                            for (int k = 0; k < 8; k++)
                                cg += (E[k] > Ym) ? 1 : 0;
                        }
                        cg = __shfl_sync(FULL_MASK, cg, ml) + __shfl_sync(FULL_MASK, cg, ml + 1) +
                             __shfl_sync(FULL_MASK, cg, ml + 2) + __shfl_sync(FULL_MASK, cg, ml + 3);
                        const u32 evE = __shfl_sync(FULL_MASK, E[7], ml + 3);
                        // new position inside the deep list (when the entry stays there): the deep entries
                        // above d with a larger key; the list is sorted, so they are a prefix
                        int pd = 0;
#pragma unroll
                        for (int t = 0; t < 7; t++)
                            pd += (lane + 32 * t < d && vE[t] > Ym) ? 1 : 0;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1)
                            pd += __shfl_xor_sync(FULL_MASK, pd, o);
                        const bool stays = cg >= 32;
                        const int lo = stays ? pd : 0; // entries lo .. d-1 move down by one, entry lo is rewritten
                        __syncwarp();
                        // every entry takes the old value of the entry above it: lane - 1 of the same round,
                        // or lane 31 of the previous round (values come from the register snapshot vE)
#pragma unroll
                        for (int t = 0; t < 7; t++) {
                            const int e = lane + 32 * t;
                            u32 pe = __shfl_up_sync(FULL_MASK, vE[t], 1);
                            const u32 we = (t > 0) ? __shfl_sync(FULL_MASK, vE[t > 0 ? t - 1 : 0], 31) : 0u;
                            if (lane == 0 && t > 0)
                                pe = we;
                            if (e > lo && e <= d)
                                mE[e] = pe;
                        }
                        if (lane == 0)
                            mE[lo] = stays ? nem : evE;
                        __syncwarp();
                        if (jq == jm) {
                            r = 32 + d;
                            rsrc = 32; // the register list sees the entry arrive from "rank 32"
                            upd = !stays;
                        }
                    }
                    // ---- common: c is among the top 32
                    {
                        const int hl = qbase + __ffs((int)bal) - 1;
                        const int rr = __shfl_sync(FULL_MASK, 8 * q + kk, hl & 31);
                        if (live && bal != 0) {
                            r = rr;
                            rsrc = rr;
                            upd = true;
                        }
                    }
                    // ---- register-list update (entry arrives from rank rsrc)
                    {
                        u32 upE = __shfl_up_sync(FULL_MASK, E[7], 1);
                        if (q == 0)
                            upE = 0xFFFFFFFFu;
                        int gq = 0; // entries and Y stay below 2^31: the sign of Y - E says E > Y
#pragma unroll
                        for (int k = 0; k < 8; k++)
                            gq += (int)((Y - E[k]) >> 31);
                        const int hi = min(rsrc - 8 * q, 7);
                        const u32 mv = (upd && hi >= gq) ? ((2u << hi) - (1u << gq)) : 0u; // bits gq .. hi
                        const u32 e0 = (upE <= Y) ? upE : ne;
#pragma unroll
                        for (int k = 7; k >= 1; k--)
                            if (mv & (1u << k))
                                E[k] = (k == gq) ? ne : E[k - 1];
                        if (mv & 1u)
                            E[0] = e0;
                    }
                    o4 |= (u32)r << (8 * xb);
                }
                ov0 = (wi == 0) ? o4 : ov0;
                ov1 = (wi == 1) ? o4 : ov1;
                ov2 = (wi == 2) ? o4 : ov2;
                ov3 = (wi == 3) ? o4 : ov3;
            }
            if (q == tt)
                outv = make_uint4(ov0, ov1, ov2, ov3);
        }
        {
            const int p = pos0 + 16 * q;
            if (active && p < end) {
                if (p + 16 <= end) {
                    *reinterpret_cast<uint4*>(dst + p) = outv;
                } else {
                    const u32 wv[4] = { outv.x, outv.y, outv.z, outv.w };
                    for (int k = 0; k < 16; k++)
                        if (p + k < end)
                            dst[p + k] = (u8)(wv[k >> 2] >> (8 * (k & 3)));
                }
            }
        }
    }
}

void launch_sbrt_rank_only(const BufTable& bt, const BlkState* stIn, const BlkState* stOut, int nBlocks, int maxLen,
                           int mode, Workspace& ws, cudaStream_t s, u64* launches)
{
    const int maxTiles = (ws.capN + S_TILE - 1) / S_TILE;
    const int tiles = (maxLen + S_TILE - 1) / S_TILE;
    uint2* occ = reinterpret_cast<uint2*>(ws.occ);
    KLAUNCH(sbrt_occ_kernel, dim3(tiles, nBlocks), 256, s, bt, stIn, stOut, maxTiles, occ);
    KLAUNCH(sbrt_fold_kernel, nBlocks, 256, s, stIn, stOut, maxTiles, occ);
    const dim3 rg((tiles + R_WARPS - 1) / R_WARPS, nBlocks);
    const bool small = maxLen < (1 << 24);
    static int variant = -1; // KNZ_SBRT_FWD=0 selects the warp-per-tile replay (experiments)
    if (variant < 0) {
        const char* e = getenv("KNZ_SBRT_FWD");
        variant = e ? atoi(e) : 1;
    }
    const dim3 qg((tiles + 8 * RQ_WARPS - 1) / (8 * RQ_WARPS), nBlocks);
    const bool quad = variant == 1 && maxLen < (1 << 23); // keys (up to 2 * length) must fit 24 bits
    if (mode == 1) {
        if (quad)
            KLAUNCH((sbrt_rank_quad_kernel<1>), qg, RQ_WARPS * 32, s, bt, stIn, stOut, maxTiles, occ);
        else if (small)
            KLAUNCH((sbrt_rank_fast_kernel<1>), rg, R_WARPS * 32, s, bt, stIn, stOut, maxTiles, occ);
        else
            KLAUNCH((sbrt_rank_kernel<u64, 1>), rg, R_WARPS * 32, s, bt, stIn, stOut, maxTiles, occ);
    } else {
        if (quad)
            KLAUNCH((sbrt_rank_quad_kernel<2>), qg, RQ_WARPS * 32, s, bt, stIn, stOut, maxTiles, occ);
        else if (small)
            KLAUNCH((sbrt_rank_fast_kernel<2>), rg, R_WARPS * 32, s, bt, stIn, stOut, maxTiles, occ);
        else
            KLAUNCH((sbrt_rank_kernel<u64, 2>), rg, R_WARPS * 32, s, bt, stIn, stOut, maxTiles, occ);
    }
    *launches += 3;
}

void launch_sbrt_forward(const StageLaunch& L, int mode, Workspace& ws, cudaStream_t s, u64* launches)
{
    KLAUNCH(sbrt_decide_kernel, (L.nBlocks + 31) / 32, 32, s, L, 0);
    *launches += 1;
    launch_sbrt_rank_only(L.bt, L.stIn, L.stOut, L.nBlocks, L.maxLen, mode, ws, s, launches);
}

void launch_sbrt_inverse(const StageLaunch& L, int mode, Workspace& ws, cudaStream_t s, u64* launches)
{
    (void)ws;
    KLAUNCH(sbrt_decide_kernel, (L.nBlocks + 31) / 32, 32, s, L, 1);
    const bool small = L.maxLen < (1 << 24);
    static int variant = -1; // KNZ_SBRT_INV=0 selects the plain distributed-list kernel (experiments)
    if (variant < 0) {
        const char* e = getenv("KNZ_SBRT_INV");
        variant = e ? atoi(e) : 1;
    }
    if (small && variant == 0) {
        if (mode == 1)
            KLAUNCH((sbrt_inverse_kernel<u32, 1>), L.nBlocks, 32, s, L.bt, L.stIn, L.stOut);
        else
            KLAUNCH((sbrt_inverse_kernel<u32, 2>), L.nBlocks, 32, s, L.bt, L.stIn, L.stOut);
        *launches += 2;
        return;
    }
    if (mode == 1) {
        if (small)
            KLAUNCH((sbrt_inverse_fast_kernel<1>), L.nBlocks, 32, s, L.bt, L.stIn, L.stOut);
        else
            KLAUNCH((sbrt_inverse_kernel<u64, 1>), L.nBlocks, 32, s, L.bt, L.stIn, L.stOut);
    } else {
        if (small)
            KLAUNCH((sbrt_inverse_fast_kernel<2>), L.nBlocks, 32, s, L.bt, L.stIn, L.stOut);
        else
            KLAUNCH((sbrt_inverse_kernel<u64, 2>), L.nBlocks, 32, s, L.bt, L.stIn, L.stOut);
    }
    *launches += 2;
}
