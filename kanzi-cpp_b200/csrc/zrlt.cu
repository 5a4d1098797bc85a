// zrlt.cu -- Zero Run Length Transform (kanzi ZRLT) on sm_100a, both directions.
//
// Reference: transform/ZRLT.cpp:27-117 (forward), :119-215 (inverse).
// Forward: a run of L zeros becomes the binary digits of L+1 below its MSB, one
// digit per byte (0/1); v -> v+1; 0xFE/0xFF -> 0xFF, v-0xFE.
// The sequential scanner is replaced by a three-step scan: each 16-byte thread
// segment is summarised by an associative element (leading run, interior output
// bytes, trailing run), CTA tiles are reduced, one warp per block folds the
// tile summaries into tile entry states, and the emit pass re-walks each segment
// from its exact entry state; a tile's output is one contiguous range, staged in
// shared memory and written as aligned 128-bit stores.  Traffic: n read twice +
// z written (forward); the inverse also zero-fills its output first.
// Whole 16-byte segments are summarised and walked through 16-bit byte-class MASKS
// (zero / non-zero, >= 0xFE, digit, escape lead, payload): run lengths come from
// ffs / clz / popc and the walks visit tokens, not bytes, so lanes do not each follow
// their own per-byte branch pattern.  KNZ_ZRLT_BYTEWALK=1 selects the byte walks
// (kept for the ragged last segment of a block).
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

template <bool INV>
__device__ __forceinline__ ZSum zshfl_up(const ZSum& v, int o)
{
    ZSum r;
    r.leadCnt = __shfl_up_sync(FULL_MASK, v.leadCnt, o);
    r.trailCnt = __shfl_up_sync(FULL_MASK, v.trailCnt, o);
    r.inter = __shfl_up_sync(FULL_MASK, v.inter, o);
    r.allz = __shfl_up_sync(FULL_MASK, v.allz, o);
    if (INV) { // the forward direction carries plain counts: the digit fields stay zero
        r.leadBits = __shfl_up_sync(FULL_MASK, v.leadBits, o);
        r.trailBits = __shfl_up_sync(FULL_MASK, v.trailBits, o);
    } else {
        r.leadBits = r.trailBits = 0;
    }
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
// `mine`; *total = combination of all threads (valid in every thread).  The warp totals are
// scanned by warp 0 (s_warp holds Z_THREADS / 32 + 1 elements: the exclusive warp prefixes, then
// the CTA total).
template <bool INV>
__device__ ZSum zblock_scan(const ZSum& mine, ZSum* s_warp /*[Z_THREADS / 32 + 1]*/, ZSum* total)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    ZSum inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const ZSum t = zshfl_up<INV>(inc, o);
        if (lane >= o)
            inc = zcombine<INV>(t, inc);
    }
    ZSum exc = zshfl_up<INV>(inc, 1);
    if (lane == 0)
        exc = zidentity();
    if (lane == 31)
        s_warp[w] = inc;
    __syncthreads();
    if (w == 0) {
        ZSum wi = (lane < Z_THREADS / 32) ? s_warp[lane] : zidentity();
#pragma unroll
        for (int o = 1; o < Z_THREADS / 32; o <<= 1) {
            const ZSum t = zshfl_up<INV>(wi, o);
            if (lane >= o)
                wi = zcombine<INV>(t, wi);
        }
        ZSum we = zshfl_up<INV>(wi, 1);
        if (lane == 0)
            we = zidentity();
        if (lane < Z_THREADS / 32)
            s_warp[lane] = we;
        if (lane == Z_THREADS / 32 - 1)
            s_warp[Z_THREADS / 32] = wi;
    }
    __syncthreads();
    const ZSum base = s_warp[w];
    *total = s_warp[Z_THREADS / 32];
    __syncthreads();
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

// ---- byte-class masks of a whole 16-byte segment (bit k <-> byte k)
__device__ __forceinline__ u32 znz4(u32 w) // bit k set iff byte k of w is non-zero
{
    const u32 y = ((((w & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | w) >> 7) & 0x01010101u;
    return (y * 0x01020408u) >> 24;
}

__device__ __forceinline__ u32 znz16(u32 a, u32 b, u32 c, u32 d)
{
    return znz4(a) | (znz4(b) << 4) | (znz4(c) << 8) | (znz4(d) << 12);
}

__device__ __forceinline__ u32 zbyte_at(const u32 w[4], int k) // byte k, k not a compile-time constant
{
    const u32 a = (k & 8) ? w[2] : w[0];
    const u32 b = (k & 8) ? w[3] : w[1];
    return __byte_perm(a, b, (u32)(k & 7)) & 0xFFu;
}

// the `len` mask bits from bit `s` on, read as a binary number whose first bit is the most significant
__device__ __forceinline__ u32 zdigits(u32 vm, int s, int len) // 1 <= len <= 16
{
    return __brev((vm >> s) & ((1u << len) - 1u)) >> (32 - len);
}

// ---- forward: per-thread walk over <= 16 bytes
__device__ __forceinline__ ZSum zfwd_summary_bytes(const u32 w[4], int cnt)
{
    ZSum S = zidentity();
    S.allz = 1;
    u32 run = 0;
    bool seen = false;
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

// Whole segment from masks.  An interior zero run (non-zero bytes on both sides) is at most 14 long,
// so its digit count ilog2(len + 1) is [len >= 1] + [len >= 3] + [len >= 7]: three popcounts over the
// run starts.  *nzOut = mask of the non-zero bytes.
__device__ __forceinline__ ZSum zfwd_summary16(const u32 w[4], u32* nzOut)
{
    ZSum S = zidentity();
    const u32 nz = znz16(w[0], w[1], w[2], w[3]);
    *nzOut = nz;
    if (nz == 0) {
        S.leadCnt = S.trailCnt = 16;
        return S;
    }
    // v >= 0xFE  <=>  (~v & 0xFE) == 0
    const u32 fe = ~znz16(~w[0] & 0xFEFEFEFEu, ~w[1] & 0xFEFEFEFEu, ~w[2] & 0xFEFEFEFEu, ~w[3] & 0xFEFEFEFEu) & 0xFFFFu;
    const int lead = __ffs((int)nz) - 1;
    const int trail = __clz((int)nz) - 16;
    const u32 zi = ~nz & ~((1u << lead) - 1u) & (0xFFFFu >> trail); // zero bytes of the interior runs
    const u32 starts = zi & ~(zi << 1);
    const u32 a3 = zi & (zi >> 1) & (zi >> 2);
    const u32 a7 = a3 & (a3 >> 3) & (zi >> 6);
    S.inter = (u32)(__popc(nz) + __popc(fe) + __popc(starts) + __popc(starts & a3) + __popc(starts & a7));
    S.leadCnt = (u32)lead;
    S.trailCnt = (u32)trail;
    S.allz = 0;
    return S;
}

template <bool LEAN>
__device__ __forceinline__ ZSum zfwd_summary(const u32 w[4], int cnt, u32* nz)
{
    if (LEAN && cnt == 16)
        return zfwd_summary16(w, nz);
    *nz = 0;
    return zfwd_summary_bytes(w, cnt);
}

template <bool LEAN>
__global__ void __launch_bounds__(Z_THREADS)
zrlt_fwd_sum_kernel(BufTable bt, const BlkState* __restrict__ st, int maxTiles, ZSum* __restrict__ tileSum)
{
    __shared__ ZSum s_warp[Z_THREADS / 32 + 1];
    const int b = blockIdx.y, t = blockIdx.x;
    const BlkState bs = st[b];
    const int n = bs.len;
    if ((i64)t * Z_TILE >= n)
        return;
    const u8* __restrict__ src = blk_src(bt, bs, b);
    const int pos = t * Z_TILE + threadIdx.x * 16;
    ZSum mine = zidentity();
    if (pos < n) {
        u32 w[4], nz;
        const int cnt = zload16(src, pos, n, w);
        mine = zfwd_summary<LEAN>(w, cnt, &nz);
    }
    ZSum total;
    zblock_scan<false>(mine, s_warp, &total);
    if (threadIdx.x == 0)
        tileSum[(i64)b * maxTiles + t] = total;
}

// One warp per block: fold the tile summaries into tile entry states (32 tiles per trip: warp scan of the
// summaries, applied to the state the trip is entered in), decide accept/refuse, publish the next state.
template <bool INV>
__global__ void __launch_bounds__(32)
zrlt_fold_kernel(StageLaunch L, int maxTiles, const ZSum* __restrict__ tileSum,
                 ZEntry* __restrict__ tileEntry, u32* __restrict__ zlen)
{
    const int b = blockIdx.x;
    const int lane = threadIdx.x;
    if (b >= L.nBlocks)
        return;
    const BlkState bs = L.stIn[b];
    BlkState ns = bs;
    const int bit = 1 << (7 - L.stageIdx);
    if (INV && (bs.flags & bit)) { // stage was skipped by the encoder
        if (lane == 0) {
            L.stOut[b] = ns;
            zlen[b] = 0xFFFFFFFFu;
        }
        return;
    }
    const int n = bs.len;
    const int tiles = (n + Z_TILE - 1) / Z_TILE;
    const ZSum* __restrict__ ts = tileSum + (i64)b * maxTiles;
    ZEntry* __restrict__ te = tileEntry + (i64)b * maxTiles;
    ZEntry e;
    e.out = e.cnt = e.bits = e.pad = 0;
    bool overflow = false;
    ZSum nxt = (lane < tiles) ? ts[lane] : zidentity();
    for (int t0 = 0; t0 < tiles; t0 += 32) {
        const int t = t0 + lane;
        ZSum inc = nxt;
        nxt = (t + 32 < tiles) ? ts[t + 32] : zidentity(); // the next trip's summaries are in flight during the scan
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const ZSum up = zshfl_up<INV>(inc, o);
            if (lane >= o)
                inc = zcombine<INV>(up, inc);
        }
        ZSum exc = zshfl_up<INV>(inc, 1);
        if (lane == 0)
            exc = zidentity();
        if (t < tiles)
            te[t] = zapply<INV>(e, exc);
        ZSum tot;
        tot.leadCnt = __shfl_sync(FULL_MASK, inc.leadCnt, 31);
        tot.leadBits = __shfl_sync(FULL_MASK, inc.leadBits, 31);
        tot.trailCnt = __shfl_sync(FULL_MASK, inc.trailCnt, 31);
        tot.trailBits = __shfl_sync(FULL_MASK, inc.trailBits, 31);
        tot.inter = __shfl_sync(FULL_MASK, inc.inter, 31);
        tot.allz = __shfl_sync(FULL_MASK, inc.allz, 31);
        e = zapply<INV>(e, tot);
    }
    if (lane != 0)
        return;
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

// Output bytes of one tile staged in shared memory: 2 bytes per input byte at most, plus the digits of the run
// that was open when the tile started, plus the 16-byte alignment slack of the destination.
#define Z_STAGE_BYTES (2 * Z_TILE + 64)

// Copies the staged bytes [a0, a0 + len) of s_out to g0 + a0 .. (g0 is 16-byte aligned): whole 16-byte
// vectors inside, single bytes on the two ragged ends (their neighbours belong to other tiles).
__device__ __forceinline__ void zstage_flush(const u8* s_out, u8* __restrict__ g0, u32 a0, u32 len)
{
    const u32 tot = a0 + len;
    const u32 nvec = (tot + 15) >> 4;
    for (u32 v = threadIdx.x; v < nvec; v += Z_THREADS) {
        const u32 lo = v << 4, hi = lo + 16;
        if (lo >= a0 && hi <= tot) {
            *reinterpret_cast<uint4*>(g0 + lo) = *reinterpret_cast<const uint4*>(s_out + lo);
        } else {
            const u32 k1 = (hi < tot) ? hi : tot;
            for (u32 k = (lo > a0) ? lo : a0; k < k1; k++)
                g0[k] = s_out[k];
        }
    }
}

// LEAN: mask walks, and the tile's output goes through shared memory (it is one contiguous range,
// [entry(t).out, entry(t + 1).out) or up to the block's output length for the last tile) and leaves as
// aligned 128-bit stores; otherwise byte walks that store straight to global memory.
template <bool LEAN>
__global__ void __launch_bounds__(Z_THREADS)
zrlt_fwd_emit_kernel(BufTable bt, const BlkState* __restrict__ st, int maxTiles,
                     const ZEntry* __restrict__ tileEntry, const u32* __restrict__ zlen)
{
    __shared__ ZSum s_warp[Z_THREADS / 32 + 1];
    __shared__ __align__(16) u8 s_out[LEAN ? Z_STAGE_BYTES : 16];
    const int b = blockIdx.y, t = blockIdx.x;
    const u32 z = zlen[b];
    if (z == 0xFFFFFFFFu)
        return;
    const BlkState bs = st[b];
    const int n = bs.len;
    if ((i64)t * Z_TILE >= n)
        return;
    const u8* __restrict__ src = blk_src(bt, bs, b);
    u8* __restrict__ dst = blk_dst(bt, bs, b);
    const int pos = t * Z_TILE + threadIdx.x * 16;
    u32 w[4] = { 0, 0, 0, 0 };
    u32 nz = 0;
    int cnt = 0;
    ZSum mine = zidentity();
    if (pos < n) {
        cnt = zload16(src, pos, n, w);
        mine = zfwd_summary<LEAN>(w, cnt, &nz);
    }
    ZSum total;
    const ZSum pre = zblock_scan<false>(mine, s_warp, &total);
    const ZEntry te = tileEntry[(i64)b * maxTiles + t];
    // where this tile's bytes are written: wdst[k - wbase] is output byte k of the block
    const u32 wbase = LEAN ? te.out : 0u;
    const u32 a0 = LEAN ? (u32)(reinterpret_cast<size_t>(dst + te.out) & 15) : 0u;
    u8* __restrict__ wdst = LEAN ? (s_out + a0) : dst;
    if (pos < n) {
        const ZEntry e = zapply<false>(te, pre);
        u32 out = e.out - wbase, run = e.cnt;
        const int end = min(pos + 16, n);
        if (LEAN && cnt == 16) {
            // one trip per non-zero byte; the zeros in front of it are a difference of bit positions
            u32 m = nz;
            int prev = 0;
            while (m) {
                const int k = __ffs((int)m) - 1;
                m &= m - 1;
                const u32 v = zbyte_at(w, k);
                run += (u32)(k - prev);
                prev = k + 1;
                const u32 r = run + 1;
                if (r >= 16) { // a run that came in from the segments before: any number of digits
                    out = zemit_run(wdst, out, run);
                } else if (r > 1) { // at most three digits, most significant first
                    const int nd = ilog2_u32(r);
                    if (nd > 2)
                        wdst[out + nd - 3] = (u8)((r >> 2) & 1);
                    if (nd > 1)
                        wdst[out + nd - 2] = (u8)((r >> 1) & 1);
                    wdst[out + nd - 1] = (u8)(r & 1);
                    out += (u32)nd;
                }
                run = 0;
                const bool big = v >= 0xFE;
                wdst[out] = big ? (u8)0xFF : (u8)(v + 1);
                if (big)
                    wdst[out + 1] = (u8)(v - 0xFE);
                out += big ? 2u : 1u;
            }
            run += (u32)(16 - prev);
        } else {
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
                    out = zemit_run(wdst, out, run);
                    run = 0;
                }
                if (v >= 0xFE) {
                    wdst[out++] = 0xFF;
                    wdst[out++] = (u8)(v - 0xFE);
                } else {
                    wdst[out++] = (u8)(v + 1);
                }
            }
        }
        if (end == n && run)
            zemit_run(wdst, out, run);
    }
    if (LEAN) {
        const bool lastTile = (i64)(t + 1) * Z_TILE >= n;
        const u32 endOut = lastTile ? z : tileEntry[(i64)b * maxTiles + t + 1].out;
        __syncthreads();
        zstage_flush(s_out, dst + te.out - a0, a0, endOut - te.out);
    }
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

__device__ __forceinline__ ZSum zinv_summary_bytes(const u32 w[4], int nb, bool payload)
{
    ZSum S = zidentity();
    u32 cnt = 0, bits = 0;
    bool seen = false;
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

// Byte classes of a whole segment: dig = digit bytes (0 / 1, not a payload), lead = escape leads,
// pay = payloads, vm = bit 0 of every byte (the digit values).
struct ZInvMasks {
    u32 dig, lead, pay, vm;
};

__device__ __forceinline__ ZInvMasks zinv_masks16(const u32 w[4], bool payload0)
{
    ZInvMasks M;
    const u32 le1 = ~znz16(w[0] & 0xFEFEFEFEu, w[1] & 0xFEFEFEFEu, w[2] & 0xFEFEFEFEu, w[3] & 0xFEFEFEFEu) & 0xFFFFu;
    const u32 ff = ~znz16(~w[0], ~w[1], ~w[2], ~w[3]) & 0xFFFFu;
    M.vm = znz16(w[0] & 0x01010101u, w[1] & 0x01010101u, w[2] & 0x01010101u, w[3] & 0x01010101u);
    u32 pm = 0;
    if (ff | (u32)payload0) {
        // a 0xFF is a lead unless it is itself the payload of the lead before it: inside a run of
        // 0xFF bytes the classes alternate, starting from the state the segment is entered in
        u32 p = payload0 ? 1u : 0u;
#pragma unroll
        for (int k = 0; k < 16; k++) {
            pm |= p << k;
            p = ((ff >> k) & 1u) & (p ^ 1u);
        }
    }
    M.pay = pm;
    M.dig = le1 & ~pm;
    M.lead = ff & ~pm;
    return M;
}

__device__ __forceinline__ ZSum zinv_summary16(const ZInvMasks& M)
{
    ZSum S = zidentity();
    const u32 nd = ~M.dig & 0xFFFFu; // tokens that end a digit run
    if (nd == 0) {
        S.leadCnt = S.trailCnt = 16;
        S.leadBits = S.trailBits = __brev(M.vm) >> 16;
        return S;
    }
    const int lead = __ffs((int)nd) - 1;
    const int trail = __clz((int)nd) - 16;
    S.allz = 0;
    S.leadCnt = (u32)lead;
    S.leadBits = lead ? zdigits(M.vm, 0, lead) : 0u;
    S.trailCnt = (u32)trail;
    S.trailBits = trail ? zdigits(M.vm, 16 - trail, trail) : 0u;
    u32 inter = (u32)__popc(nd & ~M.lead); // literals and payloads: one output byte each
    u32 di = M.dig & ~((1u << lead) - 1u) & (0xFFFFu >> trail); // digits of the interior runs (<= 14 long)
    while (di) {
        const int s = __ffs((int)di) - 1;
        const int len = __ffs((int)~(di >> s)) - 1;
        inter += ((1u << len) | zdigits(M.vm, s, len)) - 1u;
        di &= ~(((1u << len) - 1u) << s);
    }
    S.inter = inter;
    return S;
}

template <bool LEAN>
__global__ void __launch_bounds__(Z_THREADS)
zrlt_inv_sum_kernel(BufTable bt, const BlkState* __restrict__ st, int stageBit, int maxTiles,
                    ZSum* __restrict__ tileSum, int* __restrict__ errFlag)
{
    __shared__ ZSum s_warp[Z_THREADS / 32 + 1];
    const int b = blockIdx.y, t = blockIdx.x;
    const BlkState bs = st[b];
    if (bs.flags & stageBit)
        return;
    const int n = bs.len;
    if ((i64)t * Z_TILE >= n)
        return;
    const u8* __restrict__ src = blk_src(bt, bs, b);
    const int pos = t * Z_TILE + threadIdx.x * 16;
    ZSum mine = zidentity();
    if (pos < n) {
        u32 w[4];
        const int nb = zload16(src, pos, n, w);
        const bool payload0 = zinv_first_is_payload(src, pos);
        mine = (LEAN && nb == 16) ? zinv_summary16(zinv_masks16(w, payload0)) : zinv_summary_bytes(w, nb, payload0);
    }
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

// A tile's output range, zeros included, staged in shared memory when it is at most this long (a tile
// whose runs expand further stores its literals straight into the pre-zeroed output).
#define Z_INV_STAGE_BYTES 16384

template <bool LEAN>
__global__ void __launch_bounds__(Z_THREADS)
zrlt_inv_emit_kernel(BufTable bt, const BlkState* __restrict__ st, int maxTiles,
                     const ZEntry* __restrict__ tileEntry, const u32* __restrict__ zlen)
{
    __shared__ ZSum s_warp[Z_THREADS / 32 + 1];
    __shared__ __align__(16) u8 s_out[LEAN ? Z_INV_STAGE_BYTES : 16];
    const int b = blockIdx.y, t = blockIdx.x;
    const u32 z = zlen[b];
    if (z == 0xFFFFFFFFu)
        return;
    const BlkState bs = st[b];
    const int n = bs.len;
    if ((i64)t * Z_TILE >= n)
        return;
    const u8* __restrict__ src = blk_src(bt, bs, b);
    u8* __restrict__ dst = blk_dst(bt, bs, b);
    const ZEntry te = tileEntry[(i64)b * maxTiles + t];
    // the tile's output range [te.out, endOut): literals of its tokens and the zeros of the runs they end
    const bool lastTile = (i64)(t + 1) * Z_TILE >= n;
    const u32 endOut = lastTile ? z : tileEntry[(i64)b * maxTiles + t + 1].out;
    const u32 span = endOut - te.out;
    const u32 a0 = (u32)(reinterpret_cast<size_t>(dst + te.out) & 15);
    const bool staged = LEAN && (a0 + span <= (u32)Z_INV_STAGE_BYTES); // uniform over the CTA
    if (staged) {
        const u32 nvec = (a0 + span + 15) >> 4;
        const uint4 zero = make_uint4(0, 0, 0, 0);
        for (u32 v = threadIdx.x; v < nvec; v += Z_THREADS)
            reinterpret_cast<uint4*>(s_out)[v] = zero;
        __syncthreads();
    }
    const u32 wbase = staged ? te.out : 0u;
    u8* __restrict__ wdst = staged ? (s_out + a0) : dst;
    const int pos = t * Z_TILE + threadIdx.x * 16;
    u32 w[4] = { 0, 0, 0, 0 };
    int nb = 0;
    bool payload = false;
    ZInvMasks M;
    M.dig = M.lead = M.pay = M.vm = 0;
    ZSum mine = zidentity();
    if (pos < n) {
        nb = zload16(src, pos, n, w);
        payload = zinv_first_is_payload(src, pos);
        if (LEAN && nb == 16) {
            M = zinv_masks16(w, payload);
            mine = zinv_summary16(M);
        } else {
            mine = zinv_summary_bytes(w, nb, payload);
        }
    }
    ZSum total;
    const ZSum pre = zblock_scan<true>(mine, s_warp, &total);
    if (pos < n) {
        const ZEntry e = zapply<true>(te, pre);
        u32 out = e.out - wbase, cnt = e.cnt, bits = e.bits;
        if (LEAN && nb == 16) {
            // one trip per token that is not a digit; the digits in front of it come out of the masks
            u32 m = ~M.dig & 0xFFFFu;
            int prev = 0;
            while (m) {
                const int k = __ffs((int)m) - 1;
                m &= m - 1;
                const int gap = k - prev;
                if (gap) {
                    bits = (bits << gap) | zdigits(M.vm, prev, gap);
                    cnt += (u32)gap;
                }
                prev = k + 1;
                out += run_cost<true>(cnt, bits); // the zeros are already in place
                cnt = bits = 0;
                const u32 v = zbyte_at(w, k);
                if ((M.pay >> k) & 1u)
                    wdst[out++] = (u8)(0xFE + v);
                else if (!((M.lead >> k) & 1u))
                    wdst[out++] = (u8)(v - 1);
            }
        } else {
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
                out += run_cost<true>(cnt, bits); // the zeros are already in place
                cnt = bits = 0;
                if (payload) {
                    wdst[out++] = (u8)(0xFE + v);
                    payload = false;
                } else if (v == 0xFF) {
                    payload = true;
                } else {
                    wdst[out++] = (u8)(v - 1);
                }
            }
        }
    }
    if (staged) {
        __syncthreads();
        zstage_flush(s_out, dst + te.out - a0, a0, span);
    }
}

// KNZ_ZRLT_BYTEWALK=1: the byte-by-byte walks for every segment (the form the masks replaced)
static bool zrlt_lean()
{
    static const bool lean = [] {
        const char* e = getenv("KNZ_ZRLT_BYTEWALK");
        return !(e && e[0] == '1');
    }();
    return lean;
}

void launch_zrlt_forward(const StageLaunch& L, Workspace& ws, cudaStream_t s, u64* launches)
{
    const int maxTiles = (ws.capN + Z_TILE - 1) / Z_TILE;
    const int tiles = (L.maxLen + Z_TILE - 1) / Z_TILE;
    ZSum* tileSum = reinterpret_cast<ZSum*>(ws.tileA);
    ZEntry* tileEntry = reinterpret_cast<ZEntry*>(ws.tileB);
    u32* zlen = ws.zlen;
    const bool lean = zrlt_lean();
    if (lean)
        KLAUNCH(zrlt_fwd_sum_kernel<true>, dim3(tiles, L.nBlocks), Z_THREADS, s, L.bt, L.stIn, maxTiles, tileSum);
    else
        KLAUNCH(zrlt_fwd_sum_kernel<false>, dim3(tiles, L.nBlocks), Z_THREADS, s, L.bt, L.stIn, maxTiles, tileSum);
    KLAUNCH(zrlt_fold_kernel<false>, L.nBlocks, 32, s, L, maxTiles, tileSum, tileEntry, zlen);
    if (lean)
        KLAUNCH(zrlt_fwd_emit_kernel<true>, dim3(tiles, L.nBlocks), Z_THREADS, s, L.bt, L.stIn, maxTiles, tileEntry, zlen);
    else
        KLAUNCH(zrlt_fwd_emit_kernel<false>, dim3(tiles, L.nBlocks), Z_THREADS, s, L.bt, L.stIn, maxTiles, tileEntry, zlen);
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
    const bool lean = zrlt_lean();
    if (lean)
        KLAUNCH(zrlt_inv_sum_kernel<true>, dim3(tiles, L.nBlocks), Z_THREADS, s, L.bt, L.stIn, bit, maxTiles, tileSum, L.errFlag);
    else
        KLAUNCH(zrlt_inv_sum_kernel<false>, dim3(tiles, L.nBlocks), Z_THREADS, s, L.bt, L.stIn, bit, maxTiles, tileSum, L.errFlag);
    KLAUNCH(zrlt_fold_kernel<true>, L.nBlocks, 32, s, L, maxTiles, tileSum, tileEntry, zlen);
    KLAUNCH(zrlt_inv_zero_kernel, dim3(32, L.nBlocks), 256, s, L.bt, L.stIn, zlen);
    if (lean)
        KLAUNCH(zrlt_inv_emit_kernel<true>, dim3(tiles, L.nBlocks), Z_THREADS, s, L.bt, L.stIn, maxTiles, tileEntry, zlen);
    else
        KLAUNCH(zrlt_inv_emit_kernel<false>, dim3(tiles, L.nBlocks), Z_THREADS, s, L.bt, L.stIn, maxTiles, tileEntry, zlen);
    *launches += 4;
}
