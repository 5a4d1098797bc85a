// zrlt.cu -- Zero Run Length Transform (kanzi ZRLT) on sm_100a, both directions.
//
// Reference: transform/ZRLT.cpp:27-117 (forward), :119-215 (inverse).
// Forward: a run of L zeros becomes the binary digits of L+1 below its MSB, one
// digit per byte (0/1); v -> v+1; 0xFE/0xFF -> 0xFF, v-0xFE.
// The sequential scanner is replaced by a three-step scan: each 16-byte thread
// segment is summarised by an associative element (leading run, interior output
// bytes, trailing run), CTA tiles are reduced, one thread per block folds the
// tile summaries into tile entry states, and the emit pass re-walks each segment
// from its exact entry state.  Algorithmic traffic: n read twice + z written.
#include "common.cuh"
#include "kernels.h"

#define Z_TILE 4096
#define Z_THREADS 256

// Associative summary of a byte segment.  For the forward direction a "run
// symbol" is a zero byte and run lengths are plain counts (bits unused); for the
// inverse direction a run symbol is a digit byte and (cnt,bits) spell the binary
// digits seen so far.  allz: the whole segment consists of run symbols.
struct ZSum {
    u32 leadCnt, leadBits;
    u32 trailCnt, trailBits;
    u32 inter; // output bytes of everything strictly inside (complete tokens)
    u32 allz;
};

__device__ __forceinline__ u32 sadd(u32 a, u32 b) // saturating: malformed inputs must not wrap
{
    const u32 r = a + b;
    return (r < a) ? 0xFFFFFFFFu : r;
}

template <bool INV>
__device__ __forceinline__ u32 run_cost(u32 cnt, u32 bits)
{
    if (cnt == 0)
        return 0;
    if (INV) // digits -> number of zeros: ((1<<cnt)|bits) - 1   (ZRLT.cpp:141-153)
        return (cnt >= 31) ? 0xFFFFFFFFu : (((1u << cnt) | bits) - 1u);
    return (u32)ilog2_u32(cnt + 1); // zeros -> number of digits   (ZRLT.cpp:63-65)
}

template <bool INV>
__device__ __forceinline__ void run_append(u32& cnt, u32& bits, u32 cnt2, u32 bits2)
{
    if (INV) {
        bits = (cnt2 >= 32) ? bits2 : ((bits << cnt2) | bits2);
        cnt = (cnt + cnt2 > 64) ? 64 : cnt + cnt2;
    } else {
        cnt += cnt2;
    }
}

template <bool INV>
__device__ __forceinline__ ZSum zcombine(const ZSum& X, const ZSum& Y)
{
    ZSum R;
    if (Y.allz) {
        R = X;
        run_append<INV>(R.trailCnt, R.trailBits, Y.leadCnt, Y.leadBits);
        if (X.allz) {
            R.leadCnt = R.trailCnt;
            R.leadBits = R.trailBits;
        }
        return R;
    }
    if (X.allz) {
        R = Y;
        R.leadCnt = X.leadCnt;
        R.leadBits = X.leadBits;
        run_append<INV>(R.leadCnt, R.leadBits, Y.leadCnt, Y.leadBits);
        return R;
    }
    u32 mc = X.trailCnt, mb = X.trailBits;
    run_append<INV>(mc, mb, Y.leadCnt, Y.leadBits);
    R.leadCnt = X.leadCnt;
    R.leadBits = X.leadBits;
    R.trailCnt = Y.trailCnt;
    R.trailBits = Y.trailBits;
    R.inter = sadd(sadd(X.inter, run_cost<INV>(mc, mb)), Y.inter);
    R.allz = 0;
    return R;
}

__device__ __forceinline__ ZSum zshfl_up(const ZSum& v, int o)
{
    ZSum r;
    r.leadCnt = __shfl_up_sync(FULL_MASK, v.leadCnt, o);
    r.leadBits = __shfl_up_sync(FULL_MASK, v.leadBits, o);
    r.trailCnt = __shfl_up_sync(FULL_MASK, v.trailCnt, o);
    r.trailBits = __shfl_up_sync(FULL_MASK, v.trailBits, o);
    r.inter = __shfl_up_sync(FULL_MASK, v.inter, o);
    r.allz = __shfl_up_sync(FULL_MASK, v.allz, o);
    return r;
}

__device__ __forceinline__ ZSum zidentity()
{
    ZSum r;
    r.leadCnt = r.leadBits = r.trailCnt = r.trailBits = r.inter = 0;
    r.allz = 1;
    return r;
}

// CTA-wide exclusive scan (Z_THREADS threads).  Returns the exclusive prefix of
// `mine`; *total = combination of all threads (valid in every thread).
template <bool INV>
__device__ ZSum zblock_scan(const ZSum& mine, ZSum* s_warp /*[8]*/, ZSum* total)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    ZSum inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const ZSum t = zshfl_up(inc, o);
        if (lane >= o)
            inc = zcombine<INV>(t, inc);
    }
    ZSum exc = zshfl_up(inc, 1);
    if (lane == 0)
        exc = zidentity();
    if (lane == 31)
        s_warp[w] = inc;
    __syncthreads();
    ZSum base = zidentity(), tot = zidentity();
    for (int i = 0; i < Z_THREADS / 32; i++) {
        if (i == w)
            base = tot;
        tot = zcombine<INV>(tot, s_warp[i]);
    }
    __syncthreads();
    *total = tot;
    return zcombine<INV>(base, exc);
}

// Tile entry state produced by the fold: output bytes completed so far and the
// run still open when the tile starts.
struct ZEntry {
    u32 out, cnt, bits, pad;
};

template <bool INV>
__device__ __forceinline__ ZEntry zapply(const ZEntry& e, const ZSum& S)
{
    ZEntry r = e;
    if (S.allz) {
        run_append<INV>(r.cnt, r.bits, S.leadCnt, S.leadBits);
        return r;
    }
    u32 c = e.cnt, b = e.bits;
    run_append<INV>(c, b, S.leadCnt, S.leadBits);
    r.out = sadd(sadd(e.out, run_cost<INV>(c, b)), S.inter);
    r.cnt = S.trailCnt;
    r.bits = S.trailBits;
    return r;
}

// A thread's 16-byte segment in four registers: one 128-bit load when the segment is whole and
// aligned (stage buffers always are; a caller's input may not be), else byte by byte.  Returns the
// number of valid bytes.
__device__ __forceinline__ int zload16(const u8* __restrict__ src, int pos, int n, u32 w[4])
{
    const int cnt = min(16, n - pos);
    const u8* p = src + pos;
    if (cnt == 16 && (reinterpret_cast<size_t>(p) & 15) == 0) {
        const uint4 v = *reinterpret_cast<const uint4*>(p);
        w[0] = v.x, w[1] = v.y, w[2] = v.z, w[3] = v.w;
    } else {
        w[0] = w[1] = w[2] = w[3] = 0;
        for (int k = 0; k < cnt; k++)
            w[k >> 2] |= (u32)p[k] << (8 * (k & 3));
    }
    return cnt;
}
#define ZBYTE(w, k) (((w)[(k) >> 2] >> (8 * ((k)&3))) & 0xFFu)

// ---- forward: per-thread walk over <= 16 bytes
__device__ __forceinline__ ZSum zfwd_summary(const u8* __restrict__ src, int pos, int n)
{
    ZSum S = zidentity();
    S.allz = 1;
    u32 run = 0;
    bool seen = false;
    u32 w[4];
    const int cnt = zload16(src, pos, n, w);
#pragma unroll
    for (int k = 0; k < 16; k++) {
        if (k >= cnt)
            break;
        const u32 v = ZBYTE(w, k);
        if (v == 0) {
            run++;
            continue;
        }
        if (!seen) {
            S.leadCnt = run;
            seen = true;
        } else if (run) {
            S.inter += (u32)ilog2_u32(run + 1);
        }
        run = 0;
        S.inter += (v >= 0xFE) ? 2u : 1u;
    }
    if (!seen) {
        S.leadCnt = S.trailCnt = run;
    } else {
        S.allz = 0;
        S.trailCnt = run;
    }
    return S;
}

__global__ void __launch_bounds__(Z_THREADS)
zrlt_fwd_sum_kernel(BufTable bt, const BlkState* __restrict__ st, int maxTiles, ZSum* __restrict__ tileSum)
{
    __shared__ ZSum s_warp[Z_THREADS / 32];
    const int b = blockIdx.y, t = blockIdx.x;
    const BlkState bs = st[b];
    const int n = bs.len;
    if ((i64)t * Z_TILE >= n)
        return;
    const u8* __restrict__ src = blk_src(bt, bs, b);
    const int pos = t * Z_TILE + threadIdx.x * 16;
    const ZSum mine = (pos < n) ? zfwd_summary(src, pos, n) : zidentity();
    ZSum total;
    zblock_scan<false>(mine, s_warp, &total);
    if (threadIdx.x == 0)
        tileSum[(i64)b * maxTiles + t] = total;
}

// One thread per block: fold tile summaries, decide accept/refuse, publish next state.
template <bool INV>
__global__ void zrlt_fold_kernel(StageLaunch L, int maxTiles, const ZSum* __restrict__ tileSum,
                                 ZEntry* __restrict__ tileEntry, u32* __restrict__ zlen)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= L.nBlocks)
        return;
    const BlkState bs = L.stIn[b];
    BlkState ns = bs;
    const int bit = 1 << (7 - L.stageIdx);
    if (INV && (bs.flags & bit)) { // stage was skipped by the encoder
        L.stOut[b] = ns;
        zlen[b] = 0xFFFFFFFFu;
        return;
    }
    const int n = bs.len;
    const int tiles = (n + Z_TILE - 1) / Z_TILE;
    ZEntry e;
    e.out = e.cnt = e.bits = e.pad = 0;
    bool overflow = false;
    for (int t = 0; t < tiles; t++) {
        tileEntry[(i64)b * maxTiles + t] = e;
        const ZSum S = tileSum[(i64)b * maxTiles + t];
        e = zapply<INV>(e, S);
    }
    const u32 tailCost = run_cost<INV>(e.cnt, e.bits);
    const u32 z = sadd(e.out, tailCost);
    if (z >= 0x7FFFFFFFu)
        overflow = true;
    const int cap = (bs.swaps & 1) ? L.capOdd[b] : L.capEven[b];
    bool ok;
    if (!INV) {
        // ZRLT.cpp:41-42 (needs cap >= n) and :67/:96 (every token must fit)
        ok = !overflow && cap >= n && z <= (u32)cap;
    } else {
        ok = !overflow && z <= (u32)cap;
        if (!ok)
            atomicExch(L.errFlag, KERR_BAD_STREAM);
    }
    if (ok) {
        ns.len = (int)z;
        ns.cur = next_cur(bs.cur);
        ns.swaps = bs.swaps + 1;
        if (!INV)
            ns.flags = bs.flags & ~bit;
    }
    zlen[b] = ok ? z : 0xFFFFFFFFu;
    L.stOut[b] = ns;
}

__device__ __forceinline__ u32 zemit_run(u8* __restrict__ dst, u32 out, u32 run)
{
    const u32 r = run + 1;
    for (int k = ilog2_u32(r) - 1; k >= 0; k--)
        dst[out++] = (u8)((r >> k) & 1);
    return out;
}

__global__ void __launch_bounds__(Z_THREADS)
zrlt_fwd_emit_kernel(BufTable bt, const BlkState* __restrict__ st, int maxTiles,
                     const ZEntry* __restrict__ tileEntry, const u32* __restrict__ zlen)
{
    __shared__ ZSum s_warp[Z_THREADS / 32];
    const int b = blockIdx.y, t = blockIdx.x;
    if (zlen[b] == 0xFFFFFFFFu)
        return;
    const BlkState bs = st[b];
    const int n = bs.len;
    if ((i64)t * Z_TILE >= n)
        return;
    const u8* __restrict__ src = blk_src(bt, bs, b);
    u8* __restrict__ dst = blk_dst(bt, bs, b);
    const int pos = t * Z_TILE + threadIdx.x * 16;
    const ZSum mine = (pos < n) ? zfwd_summary(src, pos, n) : zidentity();
    ZSum total;
    const ZSum pre = zblock_scan<false>(mine, s_warp, &total);
    if (pos >= n)
        return;
    const ZEntry e = zapply<false>(tileEntry[(i64)b * maxTiles + t], pre);
    u32 out = e.out, run = e.cnt;
    const int end = min(pos + 16, n);
    u32 w[4];
    const int cnt = zload16(src, pos, n, w);
#pragma unroll
    for (int k = 0; k < 16; k++) {
        if (k >= cnt)
            break;
        const u32 v = ZBYTE(w, k);
        if (v == 0) {
            run++;
            continue;
        }
        if (run) {
            out = zemit_run(dst, out, run);
            run = 0;
        }
        if (v >= 0xFE) {
            dst[out++] = 0xFF;
            dst[out++] = (u8)(v - 0xFE);
        } else {
            dst[out++] = (u8)(v + 1);
        }
    }
    if (end == n && run)
        zemit_run(dst, out, run);
}

// ---- inverse
// Token classes: digit (0/1, not an escape payload) | escape lead 0xFF | payload
// (byte after a lead, any value) | literal.
__device__ __forceinline__ bool zinv_first_is_payload(const u8* __restrict__ src, int pos)
{
    int k = 0;
    while (pos - 1 - k >= 0 && src[pos - 1 - k] == 0xFF)
        k++;
    return (k & 1) != 0;
}

__device__ __forceinline__ ZSum zinv_summary(const u8* __restrict__ src, int pos, int n)
{
    ZSum S = zidentity();
    u32 cnt = 0, bits = 0;
    bool seen = false;
    bool payload = zinv_first_is_payload(src, pos);
    u32 w[4];
    const int nb = zload16(src, pos, n, w);
#pragma unroll
    for (int k = 0; k < 16; k++) {
        if (k >= nb)
            break;
        const u32 v = ZBYTE(w, k);
        if (!payload && v <= 1) {
            bits = (bits << 1) | v;
            cnt = (cnt < 64) ? cnt + 1 : cnt;
            continue;
        }
        if (!seen) {
            S.leadCnt = cnt;
            S.leadBits = bits;
            seen = true;
        } else {
            S.inter = sadd(S.inter, run_cost<true>(cnt, bits));
        }
        cnt = bits = 0;
        if (payload) {
            S.inter = sadd(S.inter, 1);
            payload = false;
        } else if (v == 0xFF) {
            payload = true;
        } else {
            S.inter = sadd(S.inter, 1);
        }
    }
    if (!seen) {
        S.leadCnt = S.trailCnt = cnt;
        S.leadBits = S.trailBits = bits;
    } else {
        S.allz = 0;
        S.trailCnt = cnt;
        S.trailBits = bits;
    }
    return S;
}

__global__ void __launch_bounds__(Z_THREADS)
zrlt_inv_sum_kernel(BufTable bt, const BlkState* __restrict__ st, int stageBit, int maxTiles,
                    ZSum* __restrict__ tileSum, int* __restrict__ errFlag)
{
    __shared__ ZSum s_warp[Z_THREADS / 32];
    const int b = blockIdx.y, t = blockIdx.x;
    const BlkState bs = st[b];
    if (bs.flags & stageBit)
        return;
    const int n = bs.len;
    if ((i64)t * Z_TILE >= n)
        return;
    const u8* __restrict__ src = blk_src(bt, bs, b);
    const int pos = t * Z_TILE + threadIdx.x * 16;
    const ZSum mine = (pos < n) ? zinv_summary(src, pos, n) : zidentity();
    ZSum total;
    zblock_scan<true>(mine, s_warp, &total);
    if (threadIdx.x == 0) {
        tileSum[(i64)b * maxTiles + t] = total;
        // a stream may not end on an escape lead (ZRLT.cpp:182-185)
        if ((i64)(t + 1) * Z_TILE >= n && src[n - 1] == 0xFF && !zinv_first_is_payload(src, n - 1))
            atomicExch(errFlag, KERR_BAD_STREAM);
    }
}

__global__ void __launch_bounds__(256)
zrlt_inv_zero_kernel(BufTable bt, const BlkState* __restrict__ st, const u32* __restrict__ zlen)
{
    const int b = blockIdx.y;
    const u32 z = zlen[b];
    if (z == 0xFFFFFFFFu)
        return;
    const BlkState bs = st[b];
    u8* dst = blk_dst(bt, bs, b); // 16-byte aligned (buffer strides are multiples of 256)
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    const u32 n16 = (z + 15) >> 4;
    const uint4 zero = make_uint4(0, 0, 0, 0);
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x)
        d4[i] = zero;
}

__global__ void __launch_bounds__(Z_THREADS)
zrlt_inv_emit_kernel(BufTable bt, const BlkState* __restrict__ st, int maxTiles,
                     const ZEntry* __restrict__ tileEntry, const u32* __restrict__ zlen)
{
    __shared__ ZSum s_warp[Z_THREADS / 32];
    const int b = blockIdx.y, t = blockIdx.x;
    if (zlen[b] == 0xFFFFFFFFu)
        return;
    const BlkState bs = st[b];
    const int n = bs.len;
    if ((i64)t * Z_TILE >= n)
        return;
    const u8* __restrict__ src = blk_src(bt, bs, b);
    u8* __restrict__ dst = blk_dst(bt, bs, b);
    const int pos = t * Z_TILE + threadIdx.x * 16;
    const ZSum mine = (pos < n) ? zinv_summary(src, pos, n) : zidentity();
    ZSum total;
    const ZSum pre = zblock_scan<true>(mine, s_warp, &total);
    if (pos >= n)
        return;
    const ZEntry e = zapply<true>(tileEntry[(i64)b * maxTiles + t], pre);
    u32 out = e.out, cnt = e.cnt, bits = e.bits;
    bool payload = zinv_first_is_payload(src, pos);
    u32 w[4];
    const int nb = zload16(src, pos, n, w);
#pragma unroll
    for (int k = 0; k < 16; k++) {
        if (k >= nb)
            break;
        const u32 v = ZBYTE(w, k);
        if (!payload && v <= 1) {
            bits = (bits << 1) | v;
            cnt++;
            continue;
        }
        out += run_cost<true>(cnt, bits); // zeros are already in place (pre-zeroed output)
        cnt = bits = 0;
        if (payload) {
            dst[out++] = (u8)(0xFE + v);
            payload = false;
        } else if (v == 0xFF) {
            payload = true;
        } else {
            dst[out++] = (u8)(v - 1);
        }
    }
}

void launch_zrlt_forward(const StageLaunch& L, Workspace& ws, cudaStream_t s, u64* launches)
{
    const int maxTiles = (ws.capN + Z_TILE - 1) / Z_TILE;
    const int tiles = (L.maxLen + Z_TILE - 1) / Z_TILE;
    ZSum* tileSum = reinterpret_cast<ZSum*>(ws.tileA);
    ZEntry* tileEntry = reinterpret_cast<ZEntry*>(ws.tileB);
    u32* zlen = ws.zlen;
    KLAUNCH(zrlt_fwd_sum_kernel, dim3(tiles, L.nBlocks), Z_THREADS, s, L.bt, L.stIn, maxTiles, tileSum);
    KLAUNCH(zrlt_fold_kernel<false>, (L.nBlocks + 31) / 32, 32, s, L, maxTiles, tileSum, tileEntry, zlen);
    KLAUNCH(zrlt_fwd_emit_kernel, dim3(tiles, L.nBlocks), Z_THREADS, s, L.bt, L.stIn, maxTiles, tileEntry, zlen);
    *launches += 3;
}

void launch_zrlt_inverse(const StageLaunch& L, Workspace& ws, cudaStream_t s, u64* launches)
{
    const int maxTiles = (ws.capN + Z_TILE - 1) / Z_TILE;
    const int tiles = (L.maxLen + Z_TILE - 1) / Z_TILE;
    ZSum* tileSum = reinterpret_cast<ZSum*>(ws.tileA) + (i64)L.wsBlock0 * maxTiles;
    ZEntry* tileEntry = reinterpret_cast<ZEntry*>(ws.tileB) + (i64)L.wsBlock0 * maxTiles;
    u32* zlen = ws.zlen + L.wsBlock0;
    const int bit = 1 << (7 - L.stageIdx);
    KLAUNCH(zrlt_inv_sum_kernel, dim3(tiles, L.nBlocks), Z_THREADS, s, L.bt, L.stIn, bit, maxTiles, tileSum, L.errFlag);
    KLAUNCH(zrlt_fold_kernel<true>, (L.nBlocks + 31) / 32, 32, s, L, maxTiles, tileSum, tileEntry, zlen);
    KLAUNCH(zrlt_inv_zero_kernel, dim3(32, L.nBlocks), 256, s, L.bt, L.stIn, zlen);
    KLAUNCH(zrlt_inv_emit_kernel, dim3(tiles, L.nBlocks), Z_THREADS, s, L.bt, L.stIn, maxTiles, tileEntry, zlen);
    *launches += 4;
}
