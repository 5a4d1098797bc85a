// ans1.cu -- order-1 rANS block coder (kanzi ANS1) on sm_100a.
//
// Replaces ANSRangeEncoder / ANSRangeDecoder with order = 1
// (entropy/ANSRangeEncoder.cpp:59-67,83-155,194-287, entropy/ANSRangeDecoder.cpp:80-292,
// order-1 histogram Global.cpp:226-307).  An order-1 chunk is 16384 << 8 = 4 MiB, coded with
// logRange 11 by four rANS states, one per quarter of the chunk; the context of a symbol is the
// byte before it, 0 for the first symbol of every quarter.  What is parallel and what is not:
//   statistics    256 x 256 pair histogram per chunk: global atomics, every position in parallel
//   tables        one warp per (chunk, context): normalisation, header bits, encoder entries
//   pre-mapping   every position's 8-byte encoder entry is gathered by a streaming kernel, so the
//                 serial coding loop has no table and no input bytes on its dependency chain
//   coding        four serial state chains per chunk: one quad per chunk (lane k = state k), the
//                 emission order inside a step recovered with one ballot, as in the order-0 coder
//   assembly      per chunk: logRange | 256 context headers | size varint | 4 states | payload,
//                 every piece at an arbitrary bit offset (shift-merge with bitcopy)
// Decoding walks the 256 context headers of a chunk serially (each starts where the previous one
// ends), builds one slot table per context (symbol | freq | cumFreq in one 32-bit word) and runs
// one quad per chunk.  With 4 MiB blocks a block is ONE chunk = four chains of 1 Mi steps: the
// coder is latency-bound there; small blocks give one chunk per block and fill the machine.
#include <stdlib.h>

#include "common.cuh"
#include "ans_tables.cuh"
#include "kernels.h"

using namespace knz;

#define A1_CH ANS1_CHUNK
#define A1_HSLOT 512 // bytes reserved per context header (longest: 6 + 256 + 32 * (4 + 8 * 11) bits = 401 bytes)

struct A1Geom {
    int m, nChunks;
};

__device__ __forceinline__ int a1_chunks(int m) { return (m <= 32) ? 0 : (m + A1_CH - 1) / A1_CH; }

// Position o of a chunk of sz bytes: is it coded (not a raw tail byte), and is it the first symbol
// of its quarter (context 0)?  (ANSRangeEncoder.cpp:216-245 / Global.cpp:271-307)
__device__ __forceinline__ bool a1_coded(int o, int sz, bool* first)
{
    const int q = sz >> 2;
    if (q == 0) { // fewer than 4 bytes: everything is tail, but the statistics still see the bytes
        *first = (o == 0);
        return true;
    }
    *first = (o == 0) || (o == q) || (o == 2 * q) || (o == 3 * q);
    return o < 4 * q;
}

// ------------------------------------------------------------------ statistics
// freq[chunk][ctx][sym]: one global atomic per position.
__global__ void __launch_bounds__(256)
ans1_hist_kernel(BufTable bt, const BlkState* __restrict__ st, int cpb, u32* __restrict__ freq)
{
    const int b = blockIdx.y;
    const BlkState bs = st[b];
    const int m = bs.len;
    if (m <= 32)
        return;
    const u8* __restrict__ src = blk_src(bt, bs, b);
    const int base = blockIdx.x * 16384;
    for (int it = 0; it < 4; it++) {
        const int j0 = base + (it * 256 + threadIdx.x) * 16;
        if (j0 >= m)
            break;
        u8 v[17];
        v[0] = (j0 > 0) ? src[j0 - 1] : 0;
        if (j0 + 16 <= m && ((((size_t)src) + j0) & 15) == 0) {
            const uint4 q = *reinterpret_cast<const uint4*>(src + j0);
            const u32 w[4] = { q.x, q.y, q.z, q.w };
#pragma unroll
            for (int i = 0; i < 16; i++)
                v[1 + i] = (u8)(w[i >> 2] >> (8 * (i & 3)));
        } else {
            for (int i = 0; i < 16; i++)
                v[1 + i] = (j0 + i < m) ? src[j0 + i] : 0;
        }
        const int c = j0 / A1_CH; // 16-byte groups never straddle a chunk (4 MiB is a multiple of 16)
        const int sz = min(A1_CH, m - c * A1_CH);
        u32* __restrict__ f = freq + ((i64)b * cpb + c) * 65536;
        const int o0 = j0 - c * A1_CH;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (j0 + i >= m)
                break;
            bool first;
            const bool coded = a1_coded(o0 + i, sz, &first);
            // the histogram covers the coded positions only (rebuildStatistics hands 4 * quarter bytes
            // to computeHistogram); with fewer than 4 bytes it covers them all
            if (!coded)
                continue;
            const u32 ctx = first ? 0u : (u32)v[i];
            atomicAdd(&f[ctx * 256 + v[1 + i]], 1u);
        }
    }
}

// ------------------------------------------------------------------ tables
__device__ __forceinline__ u32 a1_warp_sum(u32 v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

__device__ __forceinline__ u32 a1_warp_max(u32 v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v = max(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}

__device__ __forceinline__ void a1_sput(u32* hw, u32 pos, u32 v, int n)
{
    const u32 w = pos >> 5;
    const int off = (int)(pos & 31);
    const u64 x = ((u64)v << (64 - n)) >> off;
    const u32 hi = (u32)(x >> 32), lo = (u32)x;
    if (hi)
        atomicOr(&hw[w], hi);
    if (lo)
        atomicOr(&hw[w + 1], lo);
}

// Packed encoder entry: lo = invFreq; hi = bias[0:13] | (invShift-32)[13:18] | cmplFreq[20:32]
// (same fields as ANSEncSymbol::reset, entropy/ANSRangeEncoder.hpp:92-116, any logRange <= 12).
__device__ __forceinline__ u64 a1_make_entry(u32 cum, u32 freq, int lr)
{
    if (freq >= (1u << lr))
        freq = (1u << lr) - 1;
    u32 inv, sh, bias;
    if (freq < 2) {
        inv = 0xFFFFFFFFu;
        sh = 0;
        bias = cum + (1u << lr) - 1;
    } else {
        const int shift = 32 - __clz((int)(freq - 1)); // smallest shift with freq <= 2^shift
        const u32 n1 = 1u << (shift + 15);
        const u32 q1 = n1 / freq, r1 = n1 - q1 * freq;
        const u32 n2 = r1 << 16;
        const u32 q2 = n2 / freq, r2 = n2 - q2 * freq;
        inv = (q1 << 16) + q2 + (r2 ? 1u : 0u); // ceil(2^(shift+31) / freq) mod 2^32
        sh = (u32)(shift - 1);
        bias = cum;
    }
    const u32 hi = bias | (sh << 13) | (((1u << lr) - freq) << 20);
    return ((u64)hi << 32) | inv;
}

// Serial slow path of the normalisation (entropy/EntropyUtils.cpp:205-244), lane 0 only.
__device__ __noinline__ void a1_normalize_spread(u32* f, int delta, int idxMax)
{
    const int errThr = (int)f[idxMax] >> 4;
    if (delta < 0) {
        delta += errThr;
        f[idxMax] += (u32)errThr;
    } else {
        delta -= errThr;
        f[idxMax] -= (u32)errThr;
    }
    const int inc = (delta < 0) ? 1 : -1;
    delta = (delta < 0) ? -delta : delta;
    int round = 0;
    while ((++round < 6) && (delta > 0)) {
        int adjustments = 0;
        for (int i = 0; i < 256; i++) {
            if (f[i] <= 2)
                continue;
            f[i] += (u32)inc;
            adjustments++;
            delta--;
            if (delta == 0)
                break;
        }
        if (adjustments == 0)
            break;
    }
    const u32 v = f[idxMax] - (u32)delta;
    f[idxMax] = (v > 1u) ? v : 1u;
}

#define A1_BUILD_WARPS 4
// One warp per (chunk, context): lane t owns symbols 8t..8t+7.
__global__ void __launch_bounds__(A1_BUILD_WARPS * 32)
ans1_build_kernel(const BlkState* __restrict__ st, int nBlocks, int cpb, int lr, const u32* __restrict__ freq,
                  u64* __restrict__ tenc, u8* __restrict__ hdr, u32* __restrict__ hbits)
{
    __shared__ u32 s_region[A1_BUILD_WARPS][256];
    __shared__ u16 s_v16[A1_BUILD_WARPS][256];
    __shared__ u32 s_hw[A1_BUILD_WARPS][128];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int g = blockIdx.y; // chunk index over the batch
    const int b = g / cpb, c = g - b * cpb;
    if (b >= nBlocks || c >= a1_chunks(st[b].len))
        return;
    const int ctx = blockIdx.x * A1_BUILD_WARPS + wib;
    u32* region = s_region[wib];
    u16* v16 = s_v16[wib];
    u32* hw = s_hw[wib];
    const u32* __restrict__ f = freq + (i64)g * 65536 + ctx * 256;
    u64* __restrict__ ent = tenc + (i64)g * 65536 + ctx * 256;
    u32 c8[8];
    {
        const uint4 a = *reinterpret_cast<const uint4*>(f + lane * 8);
        const uint4 d = *reinterpret_cast<const uint4*>(f + lane * 8 + 4);
        c8[0] = a.x, c8[1] = a.y, c8[2] = a.z, c8[3] = a.w;
        c8[4] = d.x, c8[5] = d.y, c8[6] = d.z, c8[7] = d.w;
    }
#pragma unroll
    for (int t = 0; t < 4; t++)
        hw[lane * 4 + t] = 0;
    u32 lsum = 0, pm = 0;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        lsum += c8[t];
        pm |= (c8[t] != 0) ? (1u << t) : 0u;
    }
    const u32 total = a1_warp_sum(lsum);
    u32* hout = reinterpret_cast<u32*>(hdr + ((i64)g * 256 + ctx) * A1_HSLOT);
    if (total == 0) { // empty context: "01" = full-alphabet flag + "no symbol" (EntropyUtils.cpp:67-70)
        if (lane == 0) {
            hout[0] = bswap32(0x40000000u);
            hbits[(i64)g * 256 + ctx] = 2;
        }
        return;
    }
    const u32 cntp = (u32)__popc(pm);
    const u32 incl = warp_incl_sum(cntp, lane);
    const int asz = (int)__shfl_sync(FULL_MASK, incl, 31);
    const u32 lanesP = __ballot_sync(FULL_MASK, pm != 0);
    const u32 scale = 1u << lr;
    if (total != scale) {
        u32 ssum = 0, best = 0; // best = (scaled << 8) | (255 - symbol): max value, lowest symbol
#pragma unroll
        for (int t = 0; t < 8; t++)
            if (c8[t]) {
                const u64 sf = (u64)c8[t] << lr; // up to 2^33 for a 4 MiB chunk
                const u32 sc = (sf <= (u64)total) ? 1u : (u32)((sf + (u64)(total >> 1)) / (u64)total);
                c8[t] = sc;
                ssum += sc;
                best = max(best, (sc << 8) | (u32)(255 - (lane * 8 + t)));
            }
        ssum = a1_warp_sum(ssum);
        best = a1_warp_max(best);
        const int idxMax = 255 - (int)(best & 0xFF);
        if (asz == 1) {
#pragma unroll
            for (int t = 0; t < 8; t++)
                if (c8[t])
                    c8[t] = scale;
        } else if (ssum != scale) {
            const int delta = (int)ssum - (int)scale;
            const int errThr = (int)(best >> 8) >> 4;
            if (((delta < 0) ? -delta : delta) <= errThr) {
#pragma unroll
                for (int t = 0; t < 8; t++)
                    if (lane * 8 + t == idxMax)
                        c8[t] -= (u32)delta;
            } else {
#pragma unroll
                for (int t = 0; t < 8; t++)
                    region[lane * 8 + t] = c8[t];
                __syncwarp();
                if (lane == 0)
                    a1_normalize_spread(region, delta, idxMax);
                __syncwarp();
#pragma unroll
                for (int t = 0; t < 8; t++)
                    c8[t] = region[lane * 8 + t];
            }
        }
    }
    __syncwarp();
    // ---- header of this context: alphabet, then frequency groups (no logRange here: it is
    // written once per chunk in front of the 256 headers, ANSRangeEncoder.cpp:87)
    u32 pos0;
    if (asz == 256) {
        pos0 = 2; // "00"
    } else {
        const int last = 31 - __clz((int)lanesP);
        if (lane == 0)
            a1_sput(hw, 0, 0x20u | (u32)last, 6);
        if (lane <= last)
            a1_sput(hw, 6 + 8 * lane, pm, 8);
        pos0 = 6 + 8 * (u32)(last + 1);
    }
    u32 bits = pos0;
    if (asz > 1) {
        u32 r = incl - cntp;
#pragma unroll
        for (int t = 0; t < 8; t++)
            if (c8[t]) {
                if (r >= 1)
                    v16[r - 1] = (u16)(c8[t] - 1);
                r++;
            }
        __syncwarp();
        const int chk = (asz >= 64) ? 8 : 6;
        const int llr = ilog2_u32((u32)lr) + 1;
        const int nv = asz - 1;
        u32 glen[2], glog[2];
        int gcnt[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int gi = lane + 32 * u;
            const int cnt = min(chk, nv - gi * chk);
            u32 mx = 0;
            for (int k = 0; k < cnt; k++)
                mx = max(mx, (u32)v16[gi * chk + k]);
            const u32 lm = mx ? (u32)ilog2_u32(mx) + 1u : 0u;
            gcnt[u] = cnt;
            glog[u] = lm;
            glen[u] = (cnt > 0) ? (u32)llr + (u32)cnt * lm : 0u;
        }
        const u32 inc0 = warp_incl_sum(glen[0], lane);
        const u32 tot0 = __shfl_sync(FULL_MASK, inc0, 31);
        const u32 inc1 = warp_incl_sum(glen[1], lane);
        const u32 tot1 = __shfl_sync(FULL_MASK, inc1, 31);
        const u32 off[2] = { pos0 + inc0 - glen[0], pos0 + tot0 + inc1 - glen[1] };
#pragma unroll
        for (int u = 0; u < 2; u++) {
            if (gcnt[u] <= 0)
                continue;
            const int gi = lane + 32 * u;
            a1_sput(hw, off[u], glog[u], llr);
            u32 o = off[u] + (u32)llr;
            if (glog[u])
                for (int k = 0; k < gcnt[u]; k++) {
                    a1_sput(hw, o, (u32)v16[gi * chk + k], (int)glog[u]);
                    o += glog[u];
                }
        }
        bits = pos0 + tot0 + tot1;
    }
    __syncwarp();
    const int nw = (int)((bits + 31) >> 5);
    for (int i = lane; i < nw; i += 32)
        hout[i] = bswap32(hw[i]);
    if (lane == 0)
        hbits[(i64)g * 256 + ctx] = bits;
    // ---- encoder entries
    u32 ls = 0;
#pragma unroll
    for (int t = 0; t < 8; t++)
        ls += c8[t];
    u32 cum = warp_incl_sum(ls, lane) - ls;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        ent[lane * 8 + t] = (c8[t] == 0) ? 0ull : a1_make_entry(cum, c8[t], lr);
        cum += c8[t];
    }
}

// ------------------------------------------------------------------ pre-mapping
// rec[b][j] = encoder entry of position j (its symbol in its context).
__global__ void __launch_bounds__(256)
ans1_map_kernel(BufTable bt, const BlkState* __restrict__ st, int cpb, const u64* __restrict__ tenc,
                u64* __restrict__ rec, i64 recStride)
{
    const int b = blockIdx.y;
    const BlkState bs = st[b];
    const int m = bs.len;
    if (m <= 32)
        return;
    const u8* __restrict__ src = blk_src(bt, bs, b);
    u64* __restrict__ r = rec + (i64)b * recStride;
    const int base = blockIdx.x * 4096;
    for (int it = 0; it < 16; it++) {
        const int j = base + it * 256 + threadIdx.x;
        if (j >= m)
            break;
        const int c = j / A1_CH;
        const int sz = min(A1_CH, m - c * A1_CH);
        if ((sz >> 2) == 0)
            continue; // no coded symbol in a chunk shorter than 4 bytes
        bool first;
        if (!a1_coded(j - c * A1_CH, sz, &first))
            continue;
        const u32 ctx = first ? 0u : (u32)src[j - 1];
        r[j] = __ldg(tenc + ((i64)b * cpb + c) * 65536 + ctx * 256 + src[j]);
    }
}

// ------------------------------------------------------------------ coding
struct A1Trailer {
    u32 payBytes; // 2 * words + tail bytes
    u32 st[4];
    u32 pad[3];
};

#define A1_STEP(E, LIVE)                                                                       \
    {                                                                                          \
        const u32 hi_ = (u32)((E) >> 32), inv_ = (u32)(E);                                     \
        const u32 cmpl_ = hi_ >> 20;                                                           \
        const u32 xmax_ = 0x80000000u - (cmpl_ << xsh); /* freq << (31 - lr) */                \
        const bool did_ = (LIVE) && (state >= xmax_);                                          \
        const u32 bal_ = __ballot_sync(FULL_MASK, did_);                                       \
        if (did_) { /* an overflowing payload keeps rewriting the last slot; cnt reports it afterwards */ \
            const u32 idx_ = min(cnt + (u32)__popc(bal_ & mBelow), maxWords - 1);             \
            wlast[-(i64)idx_] = (u16)__byte_perm(state, 0, 0x4401);                            \
            state >>= 16;                                                                      \
        }                                                                                      \
        cnt += (u32)__popc(bal_ & mQuad);                                                      \
        const u32 q_ = __funnelshift_r(__umulhi(state, inv_), 0, hi_ >> 13);                   \
        if (LIVE)                                                                              \
            state = state + (hi_ & 0x1FFFu) + q_ * cmpl_;                                      \
    }

#define A1_CODE_WARPS 1
// One quad per chunk, lane k owns state k = quarter k of the chunk, walked from its last symbol
// to its first (ANSRangeEncoder.cpp:216-245).  Words are stored backwards from the top of the
// chunk's payload region so that they land in decode order; the raw tail bytes sit above them.
__global__ void __launch_bounds__(A1_CODE_WARPS * 32)
ans1_code_kernel(BufTable bt, const BlkState* __restrict__ st, int nBlocks, int cpb, int lr,
                 const u64* __restrict__ rec, i64 recStride, u8* __restrict__ pay, i64 payStride, i64 payRegion,
                 A1Trailer* __restrict__ trailer, int* __restrict__ errFlag, int qpw)
{
    // qpw = quads (chunks) per warp: 8 when there are many chunks (throughput), 1 when there are few, so
    // that every serial chain gets an SM -- and its L1 / L2 slice -- to itself
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int j = lane >> 2, k = lane & 3;
    const i64 g = ((i64)blockIdx.x * A1_CODE_WARPS + wib) * qpw + j;
    const int b = (int)(g / cpb), c = (int)(g - (i64)b * cpb);
    int m = 0;
    bool valid = false;
    if (j < qpw && b < nBlocks) {
        m = st[b].len;
        valid = c < a1_chunks(m);
    }
    const int sz = valid ? min(A1_CH, m - c * A1_CH) : 0;
    const int quarter = sz >> 2;
    const int steps = quarter;
    int maxSteps = steps;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        maxSteps = max(maxSteps, __shfl_xor_sync(FULL_MASK, maxSteps, o));
    const int xsh = 31 - lr;
    u8* region = pay + (i64)(valid ? b : 0) * payStride + (i64)(valid ? c : 0) * payRegion;
    u8* wtop = region + payRegion - 16; // words below, tail bytes at wtop[0..2]
    u16* const wlast = reinterpret_cast<u16*>(wtop) - 1;
    const u32 maxWords = (u32)((payRegion - 16) >> 1);
    const u64* __restrict__ rp = rec + (i64)(valid ? b : 0) * recStride + (i64)c * A1_CH + (i64)(k + 1) * quarter - 1;
    // rp[-s] = entry of step s
    u32 state = 1u << 15;
    u32 cnt = 0;
    const int qsh = lane & ~3;
    const u32 mQuad = 0xFu << qsh;
    const u32 mBelow = ((1u << k) - 1u) << qsh;
    // the records of the next 24 steps are in flight while a step runs (they stream from HBM once)
    u64 ring[24];
#pragma unroll
    for (int x = 0; x < 24; x++)
        ring[x] = (x < steps) ? __ldg(rp - x) : 0ull;
    int s = 0;
    for (; s < maxSteps; s += 24) { // the last trip runs past the end with its steps masked off
#pragma unroll
        for (int x = 0; x < 24; x++) {
            const u64 e = ring[x];
            const int sn = s + 24 + x;
            ring[x] = (sn < steps) ? __ldg(rp - sn) : 0ull;
            A1_STEP(e, (s + x) < steps)
        }
    }
    const u32 s1 = __shfl_sync(FULL_MASK, state, (lane & ~3) + 1);
    const u32 s2 = __shfl_sync(FULL_MASK, state, (lane & ~3) + 2);
    const u32 s3 = __shfl_sync(FULL_MASK, state, (lane & ~3) + 3);
    if (valid && k == 0) {
        const int end4 = quarter << 2, tail = sz - end4;
        if (cnt > maxWords) {
            atomicExch(errFlag, KERR_OUT_OVERFLOW);
            cnt = maxWords;
        }
        const u8* __restrict__ src = blk_src(bt, st[b], b) + (i64)c * A1_CH;
        for (int t = 0; t < tail; t++)
            wtop[t] = src[end4 + t];
        A1Trailer T;
        T.payBytes = 2 * cnt + (u32)tail;
        T.st[0] = state, T.st[1] = s1, T.st[2] = s2, T.st[3] = s3;
        T.pad[0] = T.pad[1] = T.pad[2] = 0;
        trailer[g] = T;
    }
}

// ------------------------------------------------------------------ assembly
__device__ __forceinline__ int a1_varint_len(u32 v)
{
    int n = 1;
    while (v >= 128) {
        v >>= 7;
        n++;
    }
    return n;
}

// One CTA per block: bit offsets of the 258 pieces of every chunk, total bits of the block.
__global__ void __launch_bounds__(256)
ans1_scan_kernel(const BlkState* __restrict__ st, int cpb, int nTransforms, const u32* __restrict__ hbits,
                 const A1Trailer* __restrict__ trailer, u64* __restrict__ pieceOff, u64* __restrict__ blockBits,
                 i64 outStride, int* __restrict__ errFlag)
{
    __shared__ u32 s_w[8];
    const int b = blockIdx.x;
    const int m = st[b].len;
    const int hdrBytes = knz_hdr_bytes(m, nTransforms);
    u64 base = (u64)hdrBytes * 8;
    const int nch = a1_chunks(m);
    if (nch == 0) { // stored raw (ANSRangeEncoder.cpp:160-163)
        if (threadIdx.x == 0) {
            pieceOff[(i64)b * cpb * 258] = base;
            blockBits[b] = base + 8ull * (u64)m;
        }
        return;
    }
    for (int c = 0; c < nch; c++) {
        const i64 g = (i64)b * cpb + c;
        u32 tot;
        const u32 v = hbits[g * 256 + threadIdx.x];
        const u32 ex = block_excl_sum_256(v, s_w, &tot);
        pieceOff[g * 258 + threadIdx.x] = base + 3 + ex;
        const u32 pb = trailer[g].payBytes;
        const u64 tr = base + 3 + tot;
        const u64 pl = tr + 8ull * (u64)a1_varint_len(pb) + 128;
        if (threadIdx.x == 0) {
            pieceOff[g * 258 + 256] = tr;
            pieceOff[g * 258 + 257] = pl;
        }
        base = pl + 8ull * pb;
    }
    if (threadIdx.x == 0) {
        blockBits[b] = base;
        if ((i64)((base + 7) >> 3) + 8 > outStride)
            atomicExch(errFlag, KERR_OUT_OVERFLOW);
    }
}

// grid (64, cpb, nBlocks): CTA x merges 4 context headers; CTA 0 adds logRange and the trailer;
// the payload copy is spread over all CTAs of the chunk.
__global__ void __launch_bounds__(128)
ans1_concat_kernel(BufTable bt, const BlkState* __restrict__ st, int cpb, int lr, const u8* __restrict__ hdr,
                   const u32* __restrict__ hbits, const A1Trailer* __restrict__ trailer,
                   const u64* __restrict__ pieceOff, const u8* __restrict__ pay, i64 payStride, i64 payRegion,
                   u8* __restrict__ out, i64 outStride)
{
    const int b = blockIdx.z, c = blockIdx.y;
    const BlkState bs = st[b];
    const int m = bs.len;
    u32* dst = reinterpret_cast<u32*>(out + (i64)b * outStride);
    const int nch = a1_chunks(m);
    if (nch == 0) {
        if (c == 0 && blockIdx.x == 0)
            bitcopy(dst, pieceOff[(i64)b * cpb * 258], blk_src(bt, bs, b), 8ll * m, threadIdx.x, blockDim.x);
        return;
    }
    if (c >= nch)
        return;
    const i64 g = (i64)b * cpb + c;
    const u64* po = pieceOff + g * 258;
    for (int q = 0; q < 4; q++) {
        const int ctx = blockIdx.x * 4 + q;
        bitcopy(dst, po[ctx], hdr + (g * 256 + ctx) * A1_HSLOT, (i64)hbits[g * 256 + ctx], threadIdx.x, blockDim.x);
    }
    const A1Trailer T = trailer[g];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        put_bits_atomic(dst, po[0] - 3, (u32)(lr - 8), 3);
        u64 p = po[256];
        u32 v = T.payBytes;
        while (v >= 128) { // EntropyUtils.cpp:247-259
            put_bits_atomic(dst, p, 0x80u | (v & 0x7Fu), 8);
            p += 8;
            v >>= 7;
        }
        put_bits_atomic(dst, p, v, 8);
        p += 8;
        for (int i = 0; i < 4; i++) {
            put_bits_atomic(dst, p, T.st[i], 32);
            p += 32;
        }
    }
    // payload: [words in decode order][tail bytes] ends at wtop + tail
    const int sz = min(A1_CH, m - c * A1_CH);
    const int tail = sz - ((sz >> 2) << 2);
    const u8* region = pay + (i64)b * payStride + (i64)c * payRegion;
    const u8* src = region + payRegion - 16 + tail - (i64)T.payBytes;
    const i64 nbits = 8ll * T.payBytes;
    if (nbits <= 0)
        return;
    const u64 off = po[257];
    const u64 w0 = off >> 5, w1 = (off + (u64)nbits - 1) >> 5;
    for (u64 w = w0 + (u64)blockIdx.x * blockDim.x + threadIdx.x; w <= w1; w += (u64)gridDim.x * blockDim.x) {
        const i64 sft = (i64)(w << 5) - (i64)off;
        const u32 v = src_bits32(src, nbits, sft);
        if (sft >= 0 && sft + 32 <= nbits)
            dst[w] = bswap32(v);
        else if (v)
            atomicOr(&dst[w], bswap32(v));
    }
}

void launch_ans1_encode(const EncodeLaunch& L, cudaStream_t s, u64* launches)
{
    Ans1Work& W = *L.a1;
    const int nB = L.nBlocks, cpb = W.cpb;
    const i64 nch = (i64)nB * cpb;
    cudaMemsetAsync(W.freq, 0, (size_t)nch * 65536 * sizeof(u32), s);
    const int tiles16 = (int)((W.stageCap + 16383) / 16384);
    KLAUNCH(ans1_hist_kernel, dim3(tiles16, nB), 256, s, L.bt, L.st, cpb, W.freq);
    KLAUNCH(ans1_build_kernel, dim3(256 / A1_BUILD_WARPS, (unsigned)nch), A1_BUILD_WARPS * 32, s, L.st, nB, cpb, ANS1_LR,
            W.freq, W.tenc, W.hdr, W.hbits);
    const int tiles4 = (int)((W.stageCap + 4095) / 4096);
    KLAUNCH(ans1_map_kernel, dim3(tiles4, nB), 256, s, L.bt, L.st, cpb, W.tenc, W.rec, W.recStride);
    if (L.evK0)
        cudaEventRecord(L.evK0, s);
    const int qpwE = (nch <= 2048) ? 1 : 8;
    const int quadsPerCta = A1_CODE_WARPS * qpwE;
    KLAUNCH(ans1_code_kernel, (unsigned)((nch + quadsPerCta - 1) / quadsPerCta), A1_CODE_WARPS * 32, s, L.bt, L.st, nB,
            cpb, ANS1_LR, W.rec, W.recStride, W.pay, W.payStride, W.payRegion, (A1Trailer*)W.trailer, L.errFlag,
            qpwE);
    if (L.evK1)
        cudaEventRecord(L.evK1, s);
    KLAUNCH(ans1_scan_kernel, nB, 256, s, L.st, cpb, L.nTransforms, W.hbits, (const A1Trailer*)W.trailer, W.pieceOff,
            L.blockBits, L.outStride, L.errFlag);
    launch_out_prepare_and_header(L, s, launches);
    KLAUNCH(ans1_concat_kernel, dim3(64, cpb, nB), 128, s, L.bt, L.st, cpb, ANS1_LR, W.hdr, W.hbits,
            (const A1Trailer*)W.trailer, W.pieceOff, W.pay, W.payStride, W.payRegion, L.out, L.outStride);
    *launches += 6;
}

// ------------------------------------------------------------------ decoder
__device__ __forceinline__ u32 a1_rd_bits(const u8* __restrict__ p, u64 pos, int n)
{
    const u64 b0 = pos >> 3;
    u64 w = 0;
#pragma unroll
    for (int k = 0; k < 5; k++)
        w = (w << 8) | p[b0 + k];
    const int sh = (int)(pos & 7);
    return (u32)((w >> (40 - sh - n)) & ((n == 32) ? 0xFFFFFFFFull : ((1ull << n) - 1)));
}

__device__ __forceinline__ u32 a1_rd_win(const u32* __restrict__ w, u32 rel, int n)
{
    const u32 i = rel >> 5;
    const u64 v = ((u64)w[i] << 32) | (u64)w[i + 1];
    return (u32)((v >> (64 - (int)(rel & 31) - n)) & ((n == 32) ? 0xFFFFFFFFull : ((1ull << n) - 1)));
}

struct A1DecMeta {
    u64 payPos; // bit position of the payload (words then tail bytes)
    u32 psz;    // payload bytes
    u32 st[4];
    u32 lr;
};

#define A1_WIN_WORDS 256 // 8192 bits: refilled when fewer than 4096 remain (a context header is <= 3206 bits)
// Pass 1, one warp per block: the 256 context headers of a chunk start where the previous one
// ends (ANSRangeDecoder.cpp:80-175), and the next chunk starts after this chunk's payload: one
// serial walk.  The warp stages a window of the bit string in shared memory, lane 0 parses it and
// leaves, per context, the list (symbol | freq << 8) of its present symbols.
__global__ void __launch_bounds__(32)
ans1_dec_scan_kernel(DecodeLaunch L, int cpb, u32* __restrict__ dlist, u32* __restrict__ dasz,
                     A1DecMeta* __restrict__ meta)
{
    __shared__ u32 s_win[A1_WIN_WORDS + 2];
    __shared__ u8 s_alpha[256];
    const int b = blockIdx.x, lane = threadIdx.x;
    const int m = L.preLen[b];
    const u64 endBits = L.inBits[b];
    u64 pos = L.payStart[b];
    const u8* __restrict__ p = L.in + (i64)b * L.inStride;
    const int nch = a1_chunks(m);
    if (nch == 0) {
        if (lane == 0) {
            meta[(i64)b * cpb].payPos = pos;
            if (pos + 8ull * (u64)m > endBits)
                atomicExch(L.errFlag, KERR_BAD_STREAM);
        }
        return;
    }
    const u32* __restrict__ pw = reinterpret_cast<const u32*>(p);
    const bool aligned = (((size_t)p) & 3) == 0;
    const u64 lastWord = (endBits + 31) >> 5;
    u64 winBase = ~0ull; // bit position of s_win[0]; ~0 = nothing staged
    int err = 0;
    for (int c = 0; c < nch && !err; c++) {
        const i64 g = (i64)b * cpb + c;
        int lr = 0;
        for (int k = -1; k < 257 && !err; k++) {
            // k = -1: logRange; 0..255: context headers; 256: size varint + states
            if (winBase == ~0ull || pos < winBase || pos + 4096 > winBase + 32ull * A1_WIN_WORDS) {
                const u64 w0 = pos >> 5;
                __syncwarp();
                for (int i = lane; i < A1_WIN_WORDS + 2; i += 32) {
                    const u64 wi = w0 + (u64)i;
                    u32 v = 0;
                    if (wi < lastWord) {
                        if (aligned) {
                            v = bswap32(__ldg(&pw[wi]));
                        } else {
                            const u8* q = p + wi * 4;
                            v = ((u32)q[0] << 24) | ((u32)q[1] << 16) | ((u32)q[2] << 8) | (u32)q[3];
                        }
                    }
                    s_win[i] = v;
                }
                winBase = w0 << 5;
                __syncwarp();
            }
            u64 next = pos;
            int e = 0;
            if (lane == 0) {
                u32 rel = (u32)(pos - winBase);
                if (pos + 2 > endBits) {
                    e = KERR_BAD_STREAM;
                } else if (k < 0) {
                    lr = 8 + (int)a1_rd_win(s_win, rel, 3);
                    rel += 3;
                    if (lr > ANS1_LR)
                        e = (lr > 15) ? KERR_BAD_STREAM : KERR_UNSUPPORTED; // tables are sized for logRange <= 11
                } else if (k < 256) {
                    int asz = 0;
                    if (a1_rd_win(s_win, rel, 1) == 0) {
                        asz = (a1_rd_win(s_win, rel + 1, 1) == 0) ? 256 : 0;
                        rel += 2;
                        for (int i = 0; i < asz; i++)
                            s_alpha[i] = (u8)i;
                    } else {
                        const int last = (int)a1_rd_win(s_win, rel + 1, 5);
                        rel += 6;
                        for (int i = 0; i <= last; i++) {
                            u32 mk = a1_rd_win(s_win, rel, 8);
                            rel += 8;
                            while (mk) {
                                const int bit = __ffs((int)mk) - 1;
                                s_alpha[asz++] = (u8)(8 * i + bit);
                                mk &= mk - 1;
                            }
                        }
                    }
                    u32* __restrict__ dl = dlist + g * 65536 + k * 256;
                    const u32 scale = 1u << lr;
                    u32 sum = 0;
                    if (asz > 1) {
                        const int chk = (asz >= 64) ? 8 : 6;
                        const int llr = ilog2_u32((u32)lr) + 1;
                        for (int i = 1; i < asz && !e; i += chk) {
                            const int logMax = (int)a1_rd_win(s_win, rel, llr);
                            rel += llr;
                            if (logMax > lr) {
                                e = KERR_BAD_STREAM;
                                break;
                            }
                            const int endj = min(i + chk, asz);
                            for (int jx = i; jx < endj; jx++) {
                                u32 fq = 1;
                                if (logMax) {
                                    fq = a1_rd_win(s_win, rel, logMax) + 1;
                                    rel += logMax;
                                }
                                if (fq >= scale)
                                    e = KERR_BAD_STREAM;
                                dl[jx] = (u32)s_alpha[jx] | (fq << 8);
                                sum += fq;
                            }
                        }
                    }
                    if (asz > 0) {
                        if (scale <= sum)
                            e = KERR_BAD_STREAM;
                        else
                            dl[0] = (u32)s_alpha[0] | ((scale - sum) << 8);
                    }
                    dasz[g * 256 + k] = (u32)asz;
                } else {
                    u32 v = a1_rd_win(s_win, rel, 8);
                    rel += 8;
                    u32 szp = v & 0x7F;
                    for (int shift = 7; v >= 128 && shift <= 28; shift += 7) {
                        v = a1_rd_win(s_win, rel, 8);
                        rel += 8;
                        szp |= (v & 0x7F) << shift;
                    }
                    A1DecMeta M;
                    for (int i = 0; i < 4; i++) {
                        M.st[i] = a1_rd_win(s_win, rel, 32);
                        rel += 32;
                    }
                    M.psz = szp;
                    M.lr = (u32)lr;
                    M.payPos = winBase + rel;
                    meta[g] = M;
                    // ANSRangeDecoder.cpp:223: sz < MAX_CHUNK_SIZE and sz <= bufferSize - 2 (= 2 * chunk - 2)
                    if (szp >= (1u << 27) || (u64)szp > 2ull * A1_CH - 2)
                        e = KERR_BAD_STREAM;
                    rel += 0;
                    next = winBase + rel + 8ull * szp;
                    if (next > endBits)
                        e = KERR_BAD_STREAM;
                }
                if (k < 256)
                    next = winBase + rel;
                if (next > endBits)
                    e = KERR_BAD_STREAM;
            }
            err = __shfl_sync(FULL_MASK, e, 0);
            pos = __shfl_sync(FULL_MASK, next, 0);
            lr = __shfl_sync(FULL_MASK, lr, 0);
        }
    }
    if (err && lane == 0)
        atomicExch(L.errFlag, err);
}

#define A1_TAB_WARPS 4
// Pass 2, one warp per (chunk, context): slot table  symbol | freq << 8 | cumFreq << 20.
__global__ void __launch_bounds__(A1_TAB_WARPS * 32)
ans1_dec_tables_kernel(DecodeLaunch L, int cpb, const u32* __restrict__ dlist, const u32* __restrict__ dasz,
                       const A1DecMeta* __restrict__ meta, u32* __restrict__ tdec)
{
    __shared__ u32 s_ent[A1_TAB_WARPS][256];
    __shared__ u16 s_cum[A1_TAB_WARPS][258];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const i64 g = blockIdx.y;
    const int b = (int)(g / cpb), c = (int)(g - (i64)b * cpb);
    if (*L.errFlag != 0 || b >= L.nBlocks || c >= a1_chunks(L.preLen[b]))
        return;
    const int ctx = blockIdx.x * A1_TAB_WARPS + wib;
    const int asz = (int)dasz[g * 256 + ctx];
    if (asz <= 0 || asz > 256)
        return;
    const int lr = (int)meta[g].lr;
    const u32 scale = 1u << lr;
    const u32* __restrict__ dl = dlist + g * 65536 + ctx * 256;
    u32 e8[8], ls = 0;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        const int i = lane * 8 + t;
        e8[t] = (i < asz) ? dl[i] : 0u;
        ls += e8[t] >> 8;
    }
    u32 cum = warp_incl_sum(ls, lane) - ls;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        const int i = lane * 8 + t;
        if (i < asz) {
            const u32 fq = e8[t] >> 8;
            const u32 fs = (fq >= scale) ? scale - 1 : fq; // ANSDecSymbol::reset mirrors the encoder's clamp
            s_ent[wib][i] = (e8[t] & 0xFFu) | (fs << 8) | (cum << 20);
            s_cum[wib][i] = (u16)cum;
            cum += fq;
        }
    }
    __syncwarp();
    u32* __restrict__ T = tdec + (g * 256 + ctx) * 2048;
    for (u32 s = (u32)lane; s < scale; s += 32) {
        int lo = 0, hi = asz - 1; // last i with cum[i] <= s
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if ((u32)s_cum[wib][mid] <= s)
                lo = mid;
            else
                hi = mid - 1;
        }
        T[s] = s_ent[wib][lo];
    }
}

#define A1_DEC_WARPS 1
// Pass 3, one quad per chunk, lane k = state k = quarter k (ANSRangeDecoder.cpp:259-285).
__global__ void __launch_bounds__(A1_DEC_WARPS * 32)
ans1_decode_kernel(DecodeLaunch L, int cpb, const A1DecMeta* __restrict__ meta, const u32* __restrict__ tdec, int qpw)
{
    // qpw = quads (chunks) per warp: with few chunks every chain gets an SM of its own, whose L1 then
    // holds the slot tables of the contexts the chunk keeps returning to
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int j = lane >> 2, k = lane & 3;
    const i64 g = ((i64)blockIdx.x * A1_DEC_WARPS + wib) * qpw + j;
    const int b = (int)(g / cpb), c = (int)(g - (i64)b * cpb);
    if (*L.errFlag != 0)
        return;
    int m = 0;
    bool valid = false;
    if (j < qpw && b < L.nBlocks) {
        m = L.preLen[b];
        valid = c < a1_chunks(m);
        if (m <= 32 && c == 0) { // raw block: this quad copies it
            const u8* __restrict__ p = L.in + (i64)b * L.inStride;
            u8* __restrict__ out = L.dst + (i64)b * L.dstStride;
            const u64 pos = meta[(i64)b * cpb].payPos;
            for (int i = k; i < m; i += 4)
                out[i] = (u8)a1_rd_bits(p, pos + 8ull * i, 8);
        }
    }
    const int sz = valid ? min(A1_CH, m - c * A1_CH) : 0;
    const int quarter = sz >> 2;
    const int steps = quarter;
    int maxSteps = steps;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        maxSteps = max(maxSteps, __shfl_xor_sync(FULL_MASK, maxSteps, o));
    if (maxSteps == 0 && !valid)
        return;
    A1DecMeta M;
    M.payPos = 0, M.psz = 0, M.lr = ANS1_LR;
    M.st[0] = M.st[1] = M.st[2] = M.st[3] = 0;
    if (valid)
        M = meta[g];
    const int lr = (int)M.lr;
    const u32 mask = (1u << lr) - 1;
    u32 state = M.st[k];
    const u8* __restrict__ p = L.in + (i64)(valid ? b : 0) * L.inStride;
    u8* __restrict__ o = L.dst + (i64)(valid ? b : 0) * L.dstStride + (i64)c * A1_CH + (i64)k * quarter;
    const u32* __restrict__ T = tdec + (valid ? g : 0) * 256 * 2048;
    const u32* __restrict__ pw32 = reinterpret_cast<const u32*>(p);
    const bool aligned4 = (((size_t)p) & 3) == 0;
    const u32 maxWords = M.psz >> 1;
    const int qsh = lane & ~3;
    const u32 mQuad = 0xFu << qsh;
    const u32 mAbove = ((0xFu << (k + 1)) & 0xFu) << qsh; // states consumed before this lane's in a step: 3, 2, 1, 0
    u32 cnt = 0, prv = 0;
    bool bad = false;
    for (int s = 0; s < maxSteps; s++) {
        const bool live = s < steps;
        u32 sym = 0;
        bool need = false;
        if (live) {
            const u32 slot = state & mask;
            const u32 e = __ldg(T + prv * 2048 + slot);
            sym = e & 0xFFu;
            state = ((e >> 8) & 0xFFFu) * (state >> lr) + slot - (e >> 20);
            need = state < (1u << 15);
        }
        const u32 bal = __ballot_sync(FULL_MASK, need);
        if (need) {
            const u32 idx = cnt + (u32)__popc(bal & mAbove);
            u32 w = 0;
            if (idx < maxWords) {
                const u64 bp = M.payPos + 16ull * idx;
                if (aligned4) { // two aligned 32-bit loads + funnel shift instead of five byte loads
                    const u64 wi = bp >> 5;
                    const u32 hi = bswap32(__ldg(pw32 + wi)), lo = bswap32(__ldg(pw32 + wi + 1));
                    w = __funnelshift_l(lo, hi, (u32)(bp & 31)) >> 16;
                } else {
                    w = a1_rd_bits(p, bp, 16);
                }
            } else {
                bad = true;
            }
            state = (state << 16) | w;
        }
        cnt += (u32)__popc(bal & mQuad);
        if (live) {
            o[s] = (u8)sym;
            prv = sym;
        }
    }
    if (valid && k == 0) {
        const int count4 = quarter << 2, tail = sz - count4;
        u8* __restrict__ oc = L.dst + (i64)b * L.dstStride + (i64)c * A1_CH;
        if (2 * cnt + (u32)tail != M.psz) {
            bad = true;
        } else {
            for (int t = 0; t < tail; t++)
                oc[count4 + t] = (u8)a1_rd_bits(p, M.payPos + 16ull * cnt + 8ull * t, 8);
        }
    }
    // A payload that does not match its tables is reported through the second flag word: the other
    // chunks of the batch still decode (the header walk's flag, word 0, stops everything because chunk
    // positions derived from a bad header are not trustworthy).  The reference produces such streams
    // itself: when EntropyUtils::normalizeFrequencies leaves a residual (EntropyUtils.cpp:243) the
    // decoder infers a different first frequency than the encoder used (ANSRangeDecoder.cpp:150-158)
    // and ANSRangeDecoder::decodeChunk returns false -- reproduced here, block for block.
    if (bad)
        atomicExch(L.errFlag + 1, KERR_BAD_STREAM);
}

// Pass 3 for FEW LONG chunks (blocks of 1 MiB and more: at most a few hundred chains, each up to 2^20
// steps).  ans1_decode_kernel's step is two dependent trips to L2 / HBM (the 2 MiB slot table of the
// chunk, then the renormalisation word): ~1100 cycles.  Here ONE warp owns a chunk and keeps the
// chunk's model in its SM's shared memory instead of a slot table:
//   cum[ctx][i]   u16  cumulative frequency of the i-th present symbol of the context   129 KiB
//   sym[ctx][i]   u8   that symbol                                                        64 KiB
//   idx[ctx][k]   u8   last i with cum <= k * 2^(lr-6): entry point of the search          16 KiB
//   ring          the next 128 renormalisation words, realigned, refilled by all lanes behind the chain
// State k is replicated in the eight lanes 8k..8k+7: they look up eight consecutive cumulative
// frequencies at once and a ballot counts how many are <= the slot (one round for symbols within
// eight places of the bucket's entry point, which the 64 buckets make the common case), then all
// eight advance the same state -- no broadcast on the chain.  A step is four shared-memory reads, two
// ballots and the multiply-add: ~200 cycles.
#define A1S_CUM_STRIDE 258
#define A1S_RING 128
#define A1S_SMEM (256 * A1S_CUM_STRIDE * 2 + 256 * 256 + 256 * 64 + 256 * 2 + 256 * 2 + A1S_RING * 4)
__global__ void __launch_bounds__(32)
ans1_decode_smem_kernel(DecodeLaunch L, int cpb, const A1DecMeta* __restrict__ meta, const u32* __restrict__ dlist,
                        const u32* __restrict__ dasz)
{
    KNZ_DYN_SMEM(a1s_smem);
    u16* s_cum = reinterpret_cast<u16*>(a1s_smem);                             // [256][258]
    u8* s_sym = a1s_smem + 256 * A1S_CUM_STRIDE * 2;                           // [256][256]
    u8* s_idx = s_sym + 256 * 256;                                             // [256][64]
    u16* s_tot = reinterpret_cast<u16*>(s_idx + 256 * 64);                     // [256] sum of frequencies
    u16* s_asz = s_tot + 256;                                                  // [256]
    u32* s_ring = reinterpret_cast<u32*>(s_asz + 256);                         // [128]
    const int lane = threadIdx.x;
    const i64 g = blockIdx.x;
    const int b = (int)(g / cpb), c = (int)(g - (i64)b * cpb);
    if (*L.errFlag != 0 || b >= L.nBlocks)
        return;
    const int m = L.preLen[b];
    const u8* __restrict__ p = L.in + (i64)b * L.inStride;
    if (m <= 32) { // raw block
        if (c == 0) {
            u8* __restrict__ out = L.dst + (i64)b * L.dstStride;
            const u64 pos = meta[(i64)b * cpb].payPos;
            for (int i = lane; i < m; i += 32)
                out[i] = (u8)a1_rd_bits(p, pos + 8ull * i, 8);
        }
        return;
    }
    if (c >= a1_chunks(m))
        return;
    const A1DecMeta M = meta[g];
    const int lr = (int)M.lr;
    const u32 scale = 1u << lr, mask = scale - 1;
    const int bsh = lr - 6; // 64 buckets per context
    // ---- model of the chunk
    for (int ctx = 0; ctx < 256; ctx++) {
        const int asz = min((int)dasz[g * 256 + ctx], 256);
        const u32* __restrict__ dl = dlist + g * 65536 + ctx * 256;
        u32 run = 0;
        for (int i0 = 0; i0 < asz; i0 += 32) {
            const int i = i0 + lane;
            const u32 e = (i < asz) ? dl[i] : 0u;
            const u32 fq = e >> 8;
            const u32 inc = warp_incl_sum(fq, lane);
            if (i < asz) {
                s_cum[ctx * A1S_CUM_STRIDE + i] = (u16)min(run + inc - fq, 0xFFFFu);
                s_sym[ctx * 256 + i] = (u8)e;
            }
            run += __shfl_sync(FULL_MASK, inc, 31);
        }
        if (lane == 0) {
            s_tot[ctx] = (u16)min(run, 0xFFFFu);
            s_asz[ctx] = (u16)max(asz, 0);
        }
    }
    __syncwarp();
    for (int q = lane; q < 256 * 64; q += 32) {
        const int ctx = q >> 6;
        const u32 lim = (u32)(q & 63) << bsh;
        const int asz = s_asz[ctx];
        int lo = 0, hi = asz - 1; // last i with cum[i] <= lim (cum[0] = 0)
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if ((u32)s_cum[ctx * A1S_CUM_STRIDE + mid] <= lim)
                lo = mid;
            else
                hi = mid - 1;
        }
        s_idx[q] = (u8)max(lo, 0);
    }
    // ---- renormalisation words: ring entry j = bits [payPos + 32 j, + 32) of the block's bit string
    const u32* __restrict__ pw32 = reinterpret_cast<const u32*>(p);
    const bool aligned4 = (((size_t)p) & 3) == 0;
    const u64 wiEnd = (L.inBits[b] + 31) >> 5; // 32-bit words of the block's bit string
    auto load_word = [&](u32 j) -> u32 {
        const u64 bp = M.payPos + 32ull * j;
        const u64 wi = bp >> 5;
        if (aligned4) {
            const u32 hi = (wi < wiEnd) ? bswap32(__ldg(pw32 + wi)) : 0u;
            const u32 lo = (wi + 1 < wiEnd) ? bswap32(__ldg(pw32 + wi + 1)) : 0u;
            return __funnelshift_l(lo, hi, (u32)(bp & 31));
        }
        if ((bp >> 3) + 5 > ((L.inBits[b] + 7) >> 3) + 4)
            return 0u;
        return a1_rd_bits(p, bp, 32);
    };
    u32 filled = 0; // ring entries loaded so far
    for (; filled < A1S_RING; filled += 32)
        s_ring[(filled + lane) & (A1S_RING - 1)] = load_word(filled + (u32)lane);
    __syncwarp();
    const int sz = min(A1_CH, m - c * A1_CH);
    const int quarter = sz >> 2;
    const int k = lane >> 3, t = lane & 7;
    const u32 grp = 0xFFu << (8 * k);
    const u32 heads = 0x01010101u;
    const u32 headsAbove = heads & ~((2u << (8 * k)) - 1u); // states consumed before state k in a step: 3, 2, 1, 0
    u32 state = M.st[k];
    u8* __restrict__ o = L.dst + (i64)b * L.dstStride + (i64)c * A1_CH + (i64)k * quarter;
    const u32 maxWords = M.psz >> 1;
    u32 cnt = 0, prv = 0, pend = 0;
    int pendAge = -1;
    bool bad = false;
    for (int s = 0; s < quarter; s++) {
        // refill behind the chain: the load is issued now and stored three steps later, when it has landed
        if (pendAge >= 0) {
            if (++pendAge == 3) {
                s_ring[(filled + lane) & (A1S_RING - 1)] = pend;
                filled += 32;
                pendAge = -1;
                __syncwarp();
            }
        } else if (2 * filled - cnt < 2 * A1S_RING - 64 - 8) { // room for 32 more entries
            pend = load_word(filled + (u32)lane);
            pendAge = 0;
        }
        const u32 slot = state & mask;
        const int asz = s_asz[prv];
        const u16* __restrict__ cr = s_cum + prv * A1S_CUM_STRIDE;
        int i0 = s_idx[prv * 64 + (slot >> bsh)];
        int i;
        for (;;) {
            const int it = i0 + t;
            const bool le = (it < asz) && ((u32)cr[it] <= slot);
            const u32 v = __ballot_sync(FULL_MASK, le) & grp;
            const int n = __popc(v);
            const bool more = (n == 8);
            i = i0 + n - 1;
            if (!__any_sync(FULL_MASK, more))
                break;
            if (more)
                i0 += 8; // the symbol lies further down the context's list (entry i0 + 7 is still <= slot)
            else
                i0 = i; // settled: cum[i] <= slot < cum[i + 1]; the next round finds it again in place 0
        }
        i = max(i, 0);
        const u32 sym = s_sym[prv * 256 + i];
        const u32 cumi = cr[i];
        const u32 cnext = (i + 1 < asz) ? (u32)cr[i + 1] : (u32)s_tot[prv];
        const u32 fq = min(cnext - cumi, scale - 1); // ANSDecSymbol::reset clamps like the encoder
        state = fq * (state >> lr) + slot - cumi;
        const bool need = state < (1u << 15);
        const u32 bal = __ballot_sync(FULL_MASK, need) & heads;
        if (need) {
            const u32 wix = cnt + (u32)__popc(bal & headsAbove);
            u32 w = 0;
            if (wix < maxWords)
                w = (s_ring[(wix >> 1) & (A1S_RING - 1)] >> (16 * (1 - (wix & 1)))) & 0xFFFFu;
            else
                bad = true;
            state = (state << 16) | w;
        }
        cnt += (u32)__popc(bal);
        if (t == 0)
            o[s] = (u8)sym;
        prv = sym;
    }
    if (lane == 0) {
        const int count4 = quarter << 2, tail = sz - count4;
        u8* __restrict__ oc = L.dst + (i64)b * L.dstStride + (i64)c * A1_CH;
        if (2 * cnt + (u32)tail != M.psz) {
            bad = true;
        } else {
            for (int x = 0; x < tail; x++)
                oc[count4 + x] = (u8)a1_rd_bits(p, M.payPos + 16ull * cnt + 8ull * x, 8);
        }
    }
    if (bad)
        atomicExch(L.errFlag + 1, KERR_BAD_STREAM); // soft: the other chunks still decode (see ans1_decode_kernel)
}

void launch_ans1_decode(const DecodeLaunch& L, cudaStream_t s, u64* launches)
{
    Ans1Work& W = *L.a1;
    const int cpb = W.cpb;
    const i64 nch = (i64)L.nBlocks * cpb;
    // dlist / dasz alias the encoder's freq / hbits arrays (a context is never encoding and decoding at once)
    u32* dlist = W.freq;
    u32* dasz = W.hbits;
    cudaMemsetAsync(dasz, 0, (size_t)nch * 256 * sizeof(u32), s);
    KLAUNCH(ans1_dec_scan_kernel, L.nBlocks, 32, s, L, cpb, dlist, dasz, (A1DecMeta*)W.dmeta);
    if (L.evK0)
        cudaEventRecord(L.evK0, s);
    // few long chains: model in shared memory, one warp (= one SM's worth of shared memory) per chunk
    static int smemMax = -1;
    if (smemMax < 0) {
        const char* e = getenv("KNZ_ANS1_SMEM_CHUNKS");
        smemMax = e ? atoi(e) : 300; // two waves of one-CTA-per-SM chunks still beat the slot-table kernel
#ifndef KNZ_SIM
        cudaFuncSetAttribute(ans1_decode_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A1S_SMEM);
#endif
    }
    if (nch <= smemMax) {
        KLAUNCH_DYN(ans1_decode_smem_kernel, (unsigned)nch, 32, A1S_SMEM, s, L, cpb, (const A1DecMeta*)W.dmeta, dlist, dasz);
    } else {
        KLAUNCH(ans1_dec_tables_kernel, dim3(256 / A1_TAB_WARPS, (unsigned)nch), A1_TAB_WARPS * 32, s, L, cpb, dlist, dasz,
                (const A1DecMeta*)W.dmeta, W.tdec);
        const int qpwD = (nch <= 2048) ? 1 : 8;
        const int quadsPerCta = A1_DEC_WARPS * qpwD;
        KLAUNCH(ans1_decode_kernel, (unsigned)((nch + quadsPerCta - 1) / quadsPerCta), A1_DEC_WARPS * 32, s, L, cpb,
                (const A1DecMeta*)W.dmeta, W.tdec, qpwD);
    }
    if (L.evK1)
        cudaEventRecord(L.evK1, s);
    *launches += 3;
}

// ------------------------------------------------------------------ workspace
bool ans1_work_alloc(Ans1Work& W, int maxBlocks, int stageCap)
{
    memset(&W, 0, sizeof(W));
    W.maxBlocks = maxBlocks;
    W.stageCap = stageCap;
    W.cpb = (stageCap + A1_CH - 1) / A1_CH;
    const i64 nch = (i64)maxBlocks * W.cpb;
    const i64 chunkCap = (stageCap < A1_CH) ? stageCap : A1_CH;
    W.payRegion = ((chunkCap + (chunkCap >> 3) + 64 + 16) + 255) / 256 * 256;
    W.payStride = W.payRegion * W.cpb;
    W.recStride = ((i64)stageCap + 255) / 256 * 256;
    bool ok = true;
#define A1ALLOC(p, bytes) ok = ok && (cudaMalloc((void**)&(p), (size_t)(bytes)) == cudaSuccess)
    A1ALLOC(W.freq, nch * 65536 * sizeof(u32));
    A1ALLOC(W.tenc, nch * 65536 * sizeof(u64));
    A1ALLOC(W.rec, (i64)maxBlocks * W.recStride * sizeof(u64));
    A1ALLOC(W.hdr, nch * 256 * A1_HSLOT);
    A1ALLOC(W.hbits, nch * 256 * sizeof(u32));
    A1ALLOC(W.pay, (i64)maxBlocks * W.payStride);
    A1ALLOC(W.trailer, nch * sizeof(A1Trailer));
    A1ALLOC(W.pieceOff, nch * 258 * sizeof(u64));
    A1ALLOC(W.tdec, nch * 256 * 2048 * sizeof(u32));
    A1ALLOC(W.dmeta, nch * sizeof(A1DecMeta));
#undef A1ALLOC
    if (!ok)
        ans1_work_free(W);
    return ok;
}

void ans1_work_free(Ans1Work& W)
{
    void* d[] = { W.freq, W.tenc, W.rec, W.hdr, W.hbits, W.pay, W.trailer, W.pieceOff, W.tdec, W.dmeta };
    for (size_t i = 0; i < sizeof(d) / sizeof(d[0]); i++)
        if (d[i])
            cudaFree(d[i]);
    memset(&W, 0, sizeof(W));
}
