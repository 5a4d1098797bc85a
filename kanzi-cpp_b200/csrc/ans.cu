// ans.cu -- order-0 rANS block coder (kanzi ANS0) on sm_100a.
//
// Replaces ANSRangeEncoder::encode / ANSRangeDecoder::decode
// (entropy/ANSRangeEncoder.cpp:158-261, entropy/ANSRangeDecoder.cpp:177-292).
// A block is cut into 16 KiB chunks with independent statistics; inside a chunk
// the four interleaved rANS states evolve independently and only the byte
// positions of their 16-bit renormalisation words interleave.  Mapping:
//   one warp  = 8 consecutive chunks of one block
//   one quad  = one chunk, lane k of the quad owns state k
//   emission order inside a step is recovered with one warp ballot.
// Each chunk is encoded into a private staging slot (header bit string at the
// front, renormalisation words written backwards so they land in decode order)
// and then bit-concatenated into the block buffer by bitcat.cu.
#include <stdlib.h>

#include "common.cuh"
#include "ans_tables.cuh"
#include "kernels.h"

using namespace knz;

// ------------------------------------------------------------------ encoder
// smem per warp: 8 table regions x 256 x 8 B = 16 KiB.  While the histograms are built the
// upper KiB of every region is free: four of them hold the 4 privatised copies of the chunk
// being counted; during table construction the upper KiB of the chunk's own region is the
// scratch of the header builder (compacted frequencies + header bit string).
#define ENC_WARPS 2

__device__ __forceinline__ u32 warp_sum_u32(u32 v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

__device__ __forceinline__ u32 warp_max_u32(u32 v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v = max(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}

// OR n (1..32) bits of v (v < 2^n) into an MSB-first bit string kept as 32-bit words whose
// bit 31 is the first bit of the word (converted to stream byte order when copied out).
__device__ __forceinline__ void sput(u32* hw, u32 pos, u32 v, int n)
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

// Packed encoder entry (8 bytes), laid out so that every field costs one instruction in the
// coding loop: lo = invFreq; hi = bias[0:13] | (invShift-32)[13:18] | cmplFreq[20:32].
// Same arithmetic as ANSEncSymbol::reset (entropy/ANSRangeEncoder.hpp:92-116); the 64-bit
// division of the reciprocal is done as two 32-bit long-division steps.
__device__ __forceinline__ u64 make_enc_entry12(u32 cum, u32 freq)
{
    if (freq >= (1u << ANS0_LR))
        freq = (1u << ANS0_LR) - 1;
    u32 inv, sh, bias;
    if (freq < 2) {
        inv = 0xFFFFFFFFu;
        sh = 0;
        bias = cum + (1u << ANS0_LR) - 1;
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
    const u32 hi = bias | (sh << 13) | (((1u << ANS0_LR) - freq) << 20);
    return ((u64)hi << 32) | inv;
}

// One coding step of a quad lane (state k of a chunk).  LIVE = false masks lanes whose chunk
// has no step here; a zero entry (inactive quads) never emits and leaves the state alone.
#define ANS0_STEP(E, LIVE)                                                                   \
    {                                                                                        \
        const u32 hi_ = (u32)((E) >> 32), inv_ = (u32)(E);                                   \
        const u32 cmpl_ = hi_ >> 20;                                                         \
        const u32 xmax_ = 0x80000000u - (cmpl_ << 19); /* freq << (31 - lr) */               \
        const bool did_ = (LIVE) && (state >= xmax_);                                        \
        const u32 bal_ = __ballot_sync(FULL_MASK, did_);                                     \
        if (did_) { /* memory order [hi][lo] (ANSRangeEncoder.hpp:122-126) */                \
            wlast[-(int)(cnt + (u32)__popc(bal_ & mBelow))] = (u16)__byte_perm(state, 0, 0x4401); \
            state >>= 16;                                                                    \
        }                                                                                    \
        cnt += (u32)__popc(bal_ & mQuad);                                                    \
        const u32 q_ = __funnelshift_r(__umulhi(state, inv_), 0, hi_ >> 13);                 \
        if (LIVE)                                                                            \
            state = state + (hi_ & 0x1FFFu) + q_ * cmpl_;                                    \
    }

// Lane 0 only: slow path of the normalisation (error spread over the frequencies),
// entropy/EntropyUtils.cpp:205-244, on the scaled counts f[].
__device__ __noinline__ void normalize_spread(u32* f, int delta, int idxMax)
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
            if (f[i] <= 2) // absent symbols (0) are skipped by the same test
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

// Whole warp, one chunk: normalise the histogram in `region` (256 u32), write the chunk header
// (logRange, alphabet, frequency groups: ANSRangeEncoder.cpp:83-155, EntropyUtils.cpp:57-89)
// to `slot`, then overwrite the region with the 256 packed encoder entries.  Lane t owns
// symbols 8t..8t+7.  Returns the alphabet size; *hdrBits = header length, *partial = the last,
// partly filled header byte (the chunk's quad continues the bit string after the coding loop).
__device__ __forceinline__ int ans0_build_chunk(u32* region, int len, u8* slot, int lane, u32* hdrBits, u32* partial)
{
    u16* v16 = reinterpret_cast<u16*>(region + 256);
    u32* hw = region + 384;
    u32 c[8];
    {
        const uint4 a = *reinterpret_cast<const uint4*>(region + lane * 8);
        const uint4 b = *reinterpret_cast<const uint4*>(region + lane * 8 + 4);
        c[0] = a.x, c[1] = a.y, c[2] = a.z, c[3] = a.w;
        c[4] = b.x, c[5] = b.y, c[6] = b.z, c[7] = b.w;
    }
#pragma unroll
    for (int t = 0; t < 4; t++)
        hw[lane * 4 + t] = 0;
    u32 pm = 0;
#pragma unroll
    for (int t = 0; t < 8; t++)
        pm |= (c[t] != 0) ? (1u << t) : 0u;
    const u32 cntp = (u32)__popc(pm);
    const u32 incl = warp_incl_sum(cntp, lane);
    const int asz = (int)__shfl_sync(FULL_MASK, incl, 31);
    const u32 lanesP = __ballot_sync(FULL_MASK, pm != 0);
    const u32 total = (u32)len, scale = 1u << ANS0_LR;
    if (total != scale) {
        u32 ssum = 0, best = 0; // best = (scaled << 8) | (255 - symbol): max value, lowest symbol
#pragma unroll
        for (int t = 0; t < 8; t++)
            if (c[t]) {
                const u32 sf = c[t] << ANS0_LR; // <= 2^26
                const u32 sc = (sf <= total) ? 1u : (sf + (total >> 1)) / total;
                c[t] = sc;
                ssum += sc;
                best = max(best, (sc << 8) | (u32)(255 - (lane * 8 + t)));
            }
        ssum = warp_sum_u32(ssum);
        best = warp_max_u32(best);
        const int idxMax = 255 - (int)(best & 0xFF);
        if (asz == 1) {
#pragma unroll
            for (int t = 0; t < 8; t++)
                if (c[t])
                    c[t] = scale;
        } else if (ssum != scale) {
            const int delta = (int)ssum - (int)scale;
            const int errThr = (int)(best >> 8) >> 4;
            if (((delta < 0) ? -delta : delta) <= errThr) {
#pragma unroll
                for (int t = 0; t < 8; t++)
                    if (lane * 8 + t == idxMax)
                        c[t] -= (u32)delta;
            } else {
                __syncwarp();
#pragma unroll
                for (int t = 0; t < 8; t++)
                    region[lane * 8 + t] = c[t];
                __syncwarp();
                if (lane == 0)
                    normalize_spread(region, delta, idxMax);
                __syncwarp();
#pragma unroll
                for (int t = 0; t < 8; t++)
                    c[t] = region[lane * 8 + t];
            }
        }
    }
    __syncwarp(); // hw[] zeroed, region reads done
    // ---- header: logRange - 8 (3 bits), alphabet
    u32 pos0;
    if (asz == 256) {
        if (lane == 0)
            sput(hw, 0, (u32)(ANS0_LR - 8) << 2, 5); // "00" = full alphabet
        pos0 = 5;
    } else {
        const int last = 31 - __clz((int)lanesP); // last mask byte = symbols 8*last..
        if (lane == 0)
            sput(hw, 0, ((u32)(ANS0_LR - 8) << 6) | 0x20u | (u32)last, 9);
        if (lane <= last)
            sput(hw, 9 + 8 * lane, pm, 8); // pm == 0 writes nothing
        pos0 = 9 + 8 * (u32)(last + 1);
    }
    u32 bits = pos0;
    if (asz > 1) {
        // ---- frequency groups: present symbols in increasing order, the first one implicit
        u32 r = incl - cntp;
#pragma unroll
        for (int t = 0; t < 8; t++)
            if (c[t]) {
                if (r >= 1)
                    v16[r - 1] = (u16)(c[t] - 1);
                r++;
            }
        __syncwarp();
        const int chk = (asz >= 64) ? 8 : 6;
        const int nv = asz - 1;
        u32 glen[2], glog[2];
        int gcnt[2];
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int g = lane + 32 * u;
            const int cnt = min(chk, nv - g * chk); // <= 0: no such group
            u32 mx = 0;
            for (int k = 0; k < cnt; k++)
                mx = max(mx, (u32)v16[g * chk + k]);
            const u32 lm = mx ? (u32)ilog2_u32(mx) + 1u : 0u;
            gcnt[u] = cnt;
            glog[u] = lm;
            glen[u] = (cnt > 0) ? 4u + (u32)cnt * lm : 0u; // llr = log2(12) + 1 = 4
        }
        const u32 inc0 = warp_incl_sum(glen[0], lane);
        const u32 tot0 = __shfl_sync(FULL_MASK, inc0, 31);
        const u32 inc1 = warp_incl_sum(glen[1], lane);
        const u32 tot1 = __shfl_sync(FULL_MASK, inc1, 31);
        u32 off[2] = { pos0 + inc0 - glen[0], pos0 + tot0 + inc1 - glen[1] };
#pragma unroll
        for (int u = 0; u < 2; u++) {
            if (gcnt[u] <= 0)
                continue;
            const int g = lane + 32 * u;
            sput(hw, off[u], glog[u], 4);
            u32 o = off[u] + 4;
            if (glog[u])
                for (int k = 0; k < gcnt[u]; k++) {
                    sput(hw, o, (u32)v16[g * chk + k], (int)glog[u]);
                    o += glog[u];
                }
        }
        bits = pos0 + tot0 + tot1;
    }
    __syncwarp();
    // ---- copy the header out (stream byte order); keep the partly filled last byte
    const int nw = (int)((bits + 31) >> 5);
    u32* sw = reinterpret_cast<u32*>(slot);
    for (int i = lane; i < nw; i += 32)
        sw[i] = bswap32(hw[i]);
    *hdrBits = bits;
    *partial = (hw[bits >> 5] >> (24 - (int)((bits >> 3) & 3) * 8)) & 0xFFu;
    __syncwarp();
    // ---- encoder entries over the whole region (cumulative frequencies by warp scan)
    u32 ls = 0;
#pragma unroll
    for (int t = 0; t < 8; t++)
        ls += c[t];
    u32 cum = warp_incl_sum(ls, lane) - ls;
    u64* ent = reinterpret_cast<u64*>(region);
    u64 e[8];
#pragma unroll
    for (int t = 0; t < 8; t++) {
        e[t] = (c[t] == 0 || asz <= 1) ? 0ull : make_enc_entry12(cum, c[t]);
        cum += c[t];
    }
#pragma unroll
    for (int t = 0; t < 8; t += 2)
        *reinterpret_cast<ulonglong2*>(ent + lane * 8 + t) = make_ulonglong2(e[t], e[t + 1]);
    return asz;
}

__global__ void __launch_bounds__(ENC_WARPS * 32)
ans0_encode_kernel(BufTable bt, const BlkState* __restrict__ st, int nBlocks, int maxChunks, u8* __restrict__ slots,
                   u32* __restrict__ hdrBits, u32* __restrict__ payBytes, u32* __restrict__ payOff, int dbgStop)
{
    __shared__ __align__(16) u64 s_sym[ENC_WARPS][8][256];

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int groupsPerBlk = (maxChunks + 7) >> 3;
    const int gw = blockIdx.x * ENC_WARPS + wib;
    const int b = gw / groupsPerBlk;
    if (b >= nBlocks)
        return;
    const int c0 = (gw - b * groupsPerBlk) << 3;
    const BlkState bs = st[b];
    const int m = bs.len;
    const u8* __restrict__ src = blk_src(bt, bs, b);
    const int nChunks = (m <= 32) ? 1 : ((m + ANS_CHUNK - 1) >> 14);
    if (c0 >= nChunks)
        return;

    if (m <= 32) { // stored raw (ANSRangeEncoder.cpp:160-163)
        if (lane == 0) {
            u8* slot = slots + ((i64)b * maxChunks) * ANS_SLOT;
            for (int i = 0; i < m; i++)
                slot[ANS_WEND - m + i] = src[i];
            hdrBits[(i64)b * maxChunks] = 0;
            payBytes[(i64)b * maxChunks] = (u32)m;
            payOff[(i64)b * maxChunks] = (u32)(ANS_WEND - m);
        }
        return;
    }

    u64(*sym)[256] = s_sym[wib];

    // ---- phase A: histograms (whole warp per chunk, 4-way privatised smem atomics)
    for (int j = 0; j < 8; j++) {
        const int c = c0 + j;
        if (c >= nChunks)
            break;
        const int len = min(ANS_CHUNK, m - c * ANS_CHUNK);
        const u8* __restrict__ p = src + (i64)c * ANS_CHUNK;
        // copy k lives in the upper KiB of table region (j + 1 + k) & 7
        u32* hk[4];
#pragma unroll
        for (int k = 0; k < 4; k++)
            hk[k] = reinterpret_cast<u32*>(sym[(j + 1 + k) & 7]) + 256;
#pragma unroll
        for (int k = 0; k < 4; k++)
            for (int i = lane; i < 256; i += 32)
                hk[k][i] = 0;
        __syncwarp();
        u32* h = reinterpret_cast<u32*>(sym[(j + 1 + (lane & 3)) & 7]) + 256;
        int i = lane * 16;
        if ((((size_t)p) & 15) == 0) {
            // eight 128-bit loads in flight per lane (4 KiB per warp) before the counting starts
            for (; i + 7 * 512 + 16 <= len; i += 8 * 512) {
                uint4 v[8];
#pragma unroll
                for (int u = 0; u < 8; u++)
                    v[u] = __ldg(reinterpret_cast<const uint4*>(p + i + u * 512));
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const u32 w[4] = { v[u].x, v[u].y, v[u].z, v[u].w };
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        atomicAdd(&h[w[q] & 0xFF], 1u);
                        atomicAdd(&h[(w[q] >> 8) & 0xFF], 1u);
                        atomicAdd(&h[(w[q] >> 16) & 0xFF], 1u);
                        atomicAdd(&h[w[q] >> 24], 1u);
                    }
                }
            }
        }
        for (; i < len; i += 512) {
            if (i + 16 <= len && (((size_t)p) & 15) == 0) {
                const uint4 v = *reinterpret_cast<const uint4*>(p + i);
                const u32 w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    atomicAdd(&h[w[q] & 0xFF], 1u);
                    atomicAdd(&h[(w[q] >> 8) & 0xFF], 1u);
                    atomicAdd(&h[(w[q] >> 16) & 0xFF], 1u);
                    atomicAdd(&h[w[q] >> 24], 1u);
                }
            } else {
                for (int t = i; t < min(i + 16, len); t++)
                    atomicAdd(&h[p[t]], 1u);
            }
        }
        __syncwarp();
        u32* f = reinterpret_cast<u32*>(sym[j]);
        for (int i = lane; i < 256; i += 32)
            f[i] = hk[0][i] + hk[1][i] + hk[2][i] + hk[3][i];
        __syncwarp();
    }

    if (dbgStop == 1)
        return;
    // ---- phase B: per chunk, whole warp: normalise, header, tables
    const int j = lane >> 2, k = lane & 3;
    const int c = c0 + j;
    const bool valid = c < nChunks;
    const int len = valid ? min(ANS_CHUNK, m - c * ANS_CHUNK) : 0;
    u8* slot = slots + ((i64)b * maxChunks + (valid ? c : 0)) * ANS_SLOT;
    int active = 0;
    u32 myHdrBits = 0, myPartial = 0;
    for (int jj = 0; jj < 8; jj++) {
        const int cc = c0 + jj;
        u32* region = reinterpret_cast<u32*>(sym[jj]);
        int asz = 0;
        u32 hb = 0, part = 0;
        if (cc < nChunks) {
            const int clen = min(ANS_CHUNK, m - cc * ANS_CHUNK);
            asz = ans0_build_chunk(region, clen, slots + ((i64)b * maxChunks + cc) * ANS_SLOT, lane, &hb, &part);
        } else {
#pragma unroll
            for (int t = 0; t < 4; t++) // no chunk: a zero table keeps the quad's lanes idle in the coding loop
                *reinterpret_cast<uint4*>(region + (t * 32 + lane) * 4) = make_uint4(0, 0, 0, 0);
        }
        if (j == jj) {
            active = (asz > 1) ? 1 : 0;
            myHdrBits = hb;
            myPartial = part;
        }
    }
    __syncwarp();

    if (dbgStop == 2)
        return;
    // ---- phase C: interleaved rANS, lane k of quad j owns state k
    const int end4 = len & ~3;
    const int steps = active ? (end4 >> 2) : 0;
    int maxSteps = steps, minSteps = active ? steps : 0x7FFFFFFF;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        maxSteps = max(maxSteps, __shfl_xor_sync(FULL_MASK, maxSteps, o));
        minSteps = min(minSteps, __shfl_xor_sync(FULL_MASK, minSteps, o));
    }
    // steps every active quad takes, in whole 16-byte groups; idle quads walk chunk c0's bytes
    // against their zero table (chunk c0 is never shorter than any other chunk of the warp)
    int fast = (maxSteps > 0) ? (minSteps & ~31) : 0;
    const u32* __restrict__ words = reinterpret_cast<const u32*>(src + (i64)(active ? c : c0) * ANS_CHUNK);
    const int wtop = active ? (end4 >> 2) - 1 : fast - 1; // step s consumes word wtop - s (the quad's 4 bytes)
    if (__any_sync(FULL_MASK, (((size_t)words) & 15) != 0 || (((wtop + 1) & 3) != 0)))
        fast = 0; // 128-bit groups need 16-byte alignment from the top
    const u64* __restrict__ tab = sym[j];
    u32 state = 1u << 15; // ANS_TOP
    u16* wend = reinterpret_cast<u16*>(slot + ANS_WEND);
    u16* const wlast = wend - 1;
    u32 cnt = 0; // renormalisation words of the quad so far
    const u32 selK = 0x4440u | (u32)(3 - k); // byte 3 - k of the quad's word
    const int qsh = lane & ~3;
    const u32 mQuad = 0xFu << qsh;
    const u32 mBelow = ((1u << k) - 1u) << qsh;
#define SYM_OF(W) __byte_perm((W), 0, selK)
    int s0 = 0;
    if (fast > 0) {
        // 32 steps (128 input bytes per quad) per trip: a ring of eight 128-bit groups;
        // table entries are fetched one group ahead.
        const uint4* __restrict__ gp = reinterpret_cast<const uint4*>(words + (wtop - 3)); // group g = gp[-g]
        const int nG = fast >> 2;
#define LOADG(G) __ldg(gp - min((G), nG - 1))
#define GROUP_STEPS(WN)                                                                                         \
    {                                                                                                           \
        const u64 n0 = tab[SYM_OF((WN).w)], n1 = tab[SYM_OF((WN).z)], n2 = tab[SYM_OF((WN).y)],                 \
                  n3 = tab[SYM_OF((WN).x)];                                                                     \
        ANS0_STEP(e0, true)                                                                                     \
        ANS0_STEP(e1, true)                                                                                     \
        ANS0_STEP(e2, true)                                                                                     \
        ANS0_STEP(e3, true)                                                                                     \
        e0 = n0, e1 = n1, e2 = n2, e3 = n3;                                                                     \
    }
        uint4 a0 = LOADG(0), a1 = LOADG(1), a2 = LOADG(2), a3 = LOADG(3);
        uint4 b0 = LOADG(4), b1 = LOADG(5), b2 = LOADG(6), b3 = LOADG(7);
        u64 e0 = tab[SYM_OF(a0.w)], e1 = tab[SYM_OF(a0.z)], e2 = tab[SYM_OF(a0.y)], e3 = tab[SYM_OF(a0.x)];
        for (int g = 0; g < nG; g += 8) {
#ifndef KNZ_SIM
            if (k == 0 && g + 40 < nG) // pull the line four trips ahead into L2
                asm volatile("prefetch.global.L2 [%0];" ::"l"(gp - (g + 40)));
#endif
            a0 = LOADG(g + 8); // group g's entries are already in e0..e3: every slot is refilled right
            GROUP_STEPS(a1)    // after its last use, seven groups (28 steps) before it is needed again
            a1 = LOADG(g + 9);
            GROUP_STEPS(a2)
            a2 = LOADG(g + 10);
            GROUP_STEPS(a3)
            a3 = LOADG(g + 11);
            GROUP_STEPS(b0)
            b0 = LOADG(g + 12);
            GROUP_STEPS(b1)
            b1 = LOADG(g + 13);
            GROUP_STEPS(b2)
            b2 = LOADG(g + 14);
            GROUP_STEPS(b3)
            b3 = LOADG(g + 15);
            GROUP_STEPS(a0)
        }
        s0 = fast;
#undef GROUP_STEPS
#undef LOADG
    }
    for (; s0 < maxSteps; s0++) { // ragged end: chunks of different lengths in one warp, unaligned chunks
        const bool live = s0 < steps;
        u64 e = 0;
        if (live)
            e = tab[SYM_OF(__ldg(&words[wtop - s0]))];
        ANS0_STEP(e, live)
    }
#undef SYM_OF
    // ---- epilogue: varint size, 4 states, tail bytes
    const u32 s1 = __shfl_sync(FULL_MASK, state, (lane & ~3) + 1);
    const u32 s2 = __shfl_sync(FULL_MASK, state, (lane & ~3) + 2);
    const u32 s3 = __shfl_sync(FULL_MASK, state, (lane & ~3) + 3);
    if (valid && k == 0) {
        const i64 ci = (i64)b * maxChunks + c;
        BitSink w;
        w.p = slot + (myHdrBits >> 3);
        w.n = (int)(myHdrBits & 7);
        w.acc = (u64)(myPartial >> (8 - w.n));
        w.total = myHdrBits;
        if (active) {
            const int tail = len & 3;
            u32 P = 2 * cnt + (u32)tail;
            payBytes[ci] = P;
            payOff[ci] = (u32)(ANS_WEND - 2 * cnt);
            while (P >= 128) { // EntropyUtils.cpp:247-259
                w.put(0x80 | (P & 0x7F), 8);
                P >>= 7;
            }
            w.put(P, 8);
            w.put(state, 32);
            w.put(s1, 32);
            w.put(s2, 32);
            w.put(s3, 32);
            const u8* p = src + (i64)c * ANS_CHUNK;
            for (int t = 0; t < tail; t++)
                slot[ANS_WEND + t] = p[end4 + t];
        } else {
            payBytes[ci] = 0;
            payOff[ci] = (u32)ANS_WEND;
        }
        w.finish();
        hdrBits[ci] = w.total;
    }
}

// Raw "entropy" (NullEntropyEncoder): one pseudo-chunk per block = the bytes themselves.
__global__ void raw_meta_kernel(const BlkState* __restrict__ st, int nBlocks, int maxChunks, u32* hdrBits,
                                u32* payBytes, u32* payOff)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nBlocks)
        return;
    hdrBits[(i64)b * maxChunks] = 0;
    payBytes[(i64)b * maxChunks] = (u32)st[b].len;
    payOff[(i64)b * maxChunks] = 0;
}

// Per block: exclusive scan of chunk bit lengths, block header bytes, total bits.
// Block header: mode byte, [skip-flag byte if > 4 transforms], post-transform
// length on 1..4 bytes (io/CompressedOutputStream.cpp:757-802).
__global__ void __launch_bounds__(256)
ans_scan_kernel(const BlkState* __restrict__ st, int nBlocks, int maxChunks, int eType, int nTransforms,
                const u32* __restrict__ hdrBits, const u32* __restrict__ payBytes, u64* __restrict__ chunkOff,
                u64* __restrict__ blockBits, u8* __restrict__ out, i64 outStride, int* __restrict__ errFlag)
{
    __shared__ u32 s_w[8];
    __shared__ u64 s_carry;
    const int b = blockIdx.x;
    const BlkState bs = st[b];
    const int m = bs.len;
    const int nChunks = (eType == E_RAW || m <= 32) ? 1 : ((m + ANS_CHUNK - 1) >> 14);
    const int hdrBytes = knz_hdr_bytes(m, nTransforms);
    if (threadIdx.x == 0)
        s_carry = (u64)hdrBytes * 8;
    __syncthreads();
    for (int base = 0; base < nChunks; base += 256) {
        const int c = base + threadIdx.x;
        const i64 ci = (i64)b * maxChunks + c;
        const u32 bits = (c < nChunks) ? hdrBits[ci] + 8u * payBytes[ci] : 0u;
        u32 tot;
        const u32 ex = block_excl_sum_256(bits, s_w, &tot);
        const u64 carry = s_carry;
        if (c < nChunks)
            chunkOff[ci] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0)
            s_carry = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const u64 total = s_carry;
        blockBits[b] = total;
        if ((i64)((total + 7) >> 3) + 8 > outStride)
            atomicExch(errFlag, KERR_OUT_OVERFLOW);
    }
}

// Zero the words each block's bit string will occupy, then write the block header.
__global__ void __launch_bounds__(256)
out_prepare_kernel(const BlkState* __restrict__ st, int nTransforms, const u64* __restrict__ blockBits,
                   u8* __restrict__ out, i64 outStride)
{
    const int b = blockIdx.y;
    u8* o = out + (i64)b * outStride;
    i64 words = (i64)((blockBits[b] + 31) >> 5) + 1;
    if (words * 4 > outStride)
        words = outStride >> 2;
    u32* ow = reinterpret_cast<u32*>(o);
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (i64)gridDim.x * blockDim.x)
        ow[i] = 0;
}

__global__ void block_header_kernel(const BlkState* __restrict__ st, int nBlocks, int hdrInfo,
                                    const u64* __restrict__ blockHash, u8* __restrict__ out, i64 outStride)
{
    const int nTransforms = hdrInfo & 0xFF, ckBytes = hdrInfo >> 8;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nBlocks)
        return;
    const BlkState bs = st[b];
    const int m = bs.len;
    const int dataSize = (m < 256) ? 1 : (ilog2_u32((u32)m) >> 3) + 1;
    u8* o = out + (i64)b * outStride;
    int mode = ((dataSize - 1) & 3) << 5;
    int p = 0;
    if (nTransforms <= 4) {
        mode |= (bs.flags >> 4) & 0x0F;
        o[p++] = (u8)mode;
    } else {
        mode |= 0x10;
        o[p++] = (u8)mode;
        o[p++] = (u8)bs.flags;
    }
    for (int sh = 8 * (dataSize - 1); sh >= 0; sh -= 8)
        o[p++] = (u8)(m >> sh);
    if (ckBytes) { // XXHash of the block before the transforms, 32 or 64 bits, MSB first (:804-807)
        const u64 h = blockHash[b];
        for (int sh = 8 * (ckBytes - 1); sh >= 0; sh -= 8)
            o[p++] = (u8)(h >> sh);
    }
}

// One CTA per (chunk, block): shift-merge the chunk's header bit string and its
// payload bytes into the block's bit string.
__global__ void __launch_bounds__(128)
ans_concat_kernel(BufTable bt, const BlkState* __restrict__ st, int maxChunks, int eType,
                  const u8* __restrict__ slots, const u32* __restrict__ hdrBits, const u32* __restrict__ payBytes,
                  const u32* __restrict__ payOff, const u64* __restrict__ chunkOff, u8* __restrict__ out,
                  i64 outStride)
{
    const int b = blockIdx.y, c = (eType == E_RAW) ? 0 : blockIdx.x;
    const BlkState bs = st[b];
    const int m = bs.len;
    const int nChunks = (eType == E_RAW || m <= 32) ? 1 : ((m + ANS_CHUNK - 1) >> 14);
    if (c >= nChunks)
        return;
    const i64 ci = (i64)b * maxChunks + c;
    u32* dst = reinterpret_cast<u32*>(out + (i64)b * outStride);
    const u64 off = chunkOff[ci];
    const u32 hb = hdrBits[ci], pb = payBytes[ci];
    if (eType == E_RAW) {
        // large copy: spread over the x dimension of the grid
        const u8* src = blk_src(bt, bs, b);
        const i64 nbits = (i64)pb * 8;
        if (nbits <= 0)
            return;
        const u64 w0 = off >> 5, w1 = (off + (u64)nbits - 1) >> 5;
        for (u64 w = w0 + (u64)blockIdx.x * blockDim.x + threadIdx.x; w <= w1; w += (u64)gridDim.x * blockDim.x) {
            const i64 s = (i64)(w << 5) - (i64)off;
            const u32 v = src_bits32(src, nbits, s);
            if (s >= 0 && s + 32 <= nbits)
                dst[w] = bswap32(v);
            else if (v)
                atomicOr(&dst[w], bswap32(v));
        }
        return;
    }
    const u8* slot = slots + ci * ANS_SLOT;
    bitcopy(dst, off, slot, (i64)hb, threadIdx.x, blockDim.x);
    bitcopy(dst, off + hb, slot + payOff[ci], (i64)pb * 8, threadIdx.x, blockDim.x);
}

// Stream body: per block `lw-3` (5 bits) | bits (lw bits) | payload
// (io/CompressedOutputStream.cpp:833, :852-864).  One CTA computes offsets.
__global__ void __launch_bounds__(256)
stream_scan_kernel(const u64* __restrict__ blockBits, int nBlocks, const u64* startBitPtr, u64* __restrict__ blockOff,
                   u64* endBit)
{
    __shared__ u64 s_part[256];
    const int t = threadIdx.x;
    const int per = (nBlocks + 255) / 256;
    const int lo = min(t * per, nBlocks), hi = min(lo + per, nBlocks);
    u64 sum = 0;
    for (int i = lo; i < hi; i++) {
        const u64 w = blockBits[i];
        const u32 lw = (w < 8) ? 3u : (u32)ilog2_u32((u32)(w >> 3)) + 4u;
        sum += 5 + lw + w;
    }
    s_part[t] = sum;
    __syncthreads();
    if (t == 0) {
        u64 run = *startBitPtr;
        for (int i = 0; i < 256; i++) {
            const u64 v = s_part[i];
            s_part[i] = run;
            run += v;
        }
        *endBit = run;
    }
    __syncthreads();
    u64 run = s_part[t];
    for (int i = lo; i < hi; i++) {
        blockOff[i] = run;
        const u64 w = blockBits[i];
        const u32 lw = (w < 8) ? 3u : (u32)ilog2_u32((u32)(w >> 3)) + 4u;
        run += 5 + lw + w;
    }
}

__global__ void __launch_bounds__(256)
stream_concat_kernel(const u8* __restrict__ blockOut, i64 outStride, const u64* __restrict__ blockBits,
                     const u64* __restrict__ blockOff, u8* __restrict__ stream, const int* __restrict__ srcIndex)
{
    const int b = blockIdx.y;
    u32* dst = reinterpret_cast<u32*>(stream);
    const u64 w = blockBits[b];
    const u32 lw = (w < 8) ? 3u : (u32)ilog2_u32((u32)(w >> 3)) + 4u;
    u64 off = blockOff[b];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        put_bits_atomic(dst, off, lw - 3, 5);
        // lw <= 35: split into two puts of <= 32 bits
        if (lw > 32) {
            put_bits_atomic(dst, off + 5, (u32)(w >> 32), (int)lw - 32);
            put_bits_atomic(dst, off + 5 + (lw - 32), (u32)w, 32);
        } else {
            put_bits_atomic(dst, off + 5, (u32)w, (int)lw);
        }
    }
    off += 5 + lw;
    // srcIndex: where block b's bytes sit in blockOut (blocks gathered from several GPUs arrive grouped
    // by rank, not in stream order); NULL = in order
    const u8* src = blockOut + (i64)(srcIndex ? srcIndex[b] : b) * outStride;
    const i64 nbits = (i64)w;
    if (nbits <= 0)
        return;
    const u64 w0 = off >> 5, w1 = (off + (u64)nbits - 1) >> 5;
    for (u64 x = w0 + (u64)blockIdx.x * blockDim.x + threadIdx.x; x <= w1; x += (u64)gridDim.x * blockDim.x) {
        const i64 s = (i64)(x << 5) - (i64)off;
        const u32 v = src_bits32(src, nbits, s);
        if (s >= 0 && s + 32 <= nbits)
            dst[x] = bswap32(v);
        else if (v)
            atomicOr(&dst[x], bswap32(v));
    }
}

// ------------------------------------------------------------------ host side
void launch_out_prepare_and_header(const EncodeLaunch& L, cudaStream_t s, u64* launches)
{
    const int nB = L.nBlocks;
    i64 zx64 = (L.outStride / 4 + 255) / 256;
    const int zx = (int)(zx64 < 64 ? zx64 : 64);
    KLAUNCH(out_prepare_kernel, dim3(zx, nB), 256, s, L.st, L.nTransforms, L.blockBits, L.out, L.outStride);
    KLAUNCH(block_header_kernel, (nB + 127) / 128, 128, s, L.st, nB, L.nTransforms, L.blockHash, L.out, L.outStride);
    *launches += 2;
}

void launch_entropy_encode(const EncodeLaunch& L, cudaStream_t s, u64* launches)
{
    const int nB = L.nBlocks;
    if (L.eType == E_ANS1) {
        launch_ans1_encode(L, s, launches);
        return;
    }
    if (L.eType == E_FPAQ) {
        launch_fpaq_encode(L, s, launches);
        return;
    }
    if (L.eType == E_RAW) {
        KLAUNCH(raw_meta_kernel, (nB + 127) / 128, 128, s, L.st, nB, L.maxChunks, L.hdrBits, L.payBytes, L.payOff);
    } else if (L.eType == E_HUF) {
        launch_huffman_encode_chunks(L, s, launches);
    } else {
        const int groups = (L.maxChunks + 7) / 8;
        const i64 warps = (i64)nB * groups;
        const int ctas = (int)((warps + ENC_WARPS - 1) / ENC_WARPS);
        if (L.evK0)
            cudaEventRecord(L.evK0, s);
        static int dbgStop = -1; // timing experiments only: KNZ_ANS_STOP=1|2 ends the kernel after phase A|B
        if (dbgStop < 0) {
            const char* e = getenv("KNZ_ANS_STOP");
            dbgStop = e ? atoi(e) : 0;
        }
        KLAUNCH(ans0_encode_kernel, ctas, ENC_WARPS * 32, s, L.bt, L.st, nB, L.maxChunks, L.slots, L.hdrBits,
                L.payBytes, L.payOff, dbgStop);
        if (L.evK1)
            cudaEventRecord(L.evK1, s);
    }
    KLAUNCH(ans_scan_kernel, nB, 256, s, L.st, nB, L.maxChunks, L.eType, L.nTransforms, L.hdrBits, L.payBytes,
                                       L.chunkOff, L.blockBits, L.out, L.outStride, L.errFlag);
    launch_out_prepare_and_header(L, s, launches);
    if (L.eType == E_RAW)
        KLAUNCH(ans_concat_kernel, dim3(64, nB), 128, s, L.bt, L.st, L.maxChunks, L.eType, L.slots, L.hdrBits,
                                                       L.payBytes, L.payOff, L.chunkOff, L.out, L.outStride);
    else
        KLAUNCH(ans_concat_kernel, dim3(L.maxChunks, nB), 128, s, L.bt, L.st, L.maxChunks, L.eType, L.slots,
                                                                L.hdrBits, L.payBytes, L.payOff, L.chunkOff,
                                                                L.out, L.outStride);
    *launches += 3;
}

void launch_stream_assemble(const u8* blockOut, i64 outStride, const u64* blockBits, int nBlocks,
                            const u64* startBit, u64* blockOff, u64* endBit, u8* stream, cudaStream_t s,
                            u64* launches, const int* srcIndex)
{
    KLAUNCH(stream_scan_kernel, 1, 256, s, blockBits, nBlocks, startBit, blockOff, endBit);
    KLAUNCH(stream_concat_kernel, dim3(32, nBlocks), 256, s, blockOut, outStride, blockBits, blockOff, stream,
            srcIndex);
    *launches += 2;
}

// ------------------------------------------------------------------ decoder
// MSB-first bit fetch (n <= 32) at an arbitrary bit position.
__device__ __forceinline__ u32 rd_bits(const u8* __restrict__ p, u64 pos, int n)
{
    const u64 b0 = pos >> 3;
    u64 w = 0;
#pragma unroll
    for (int k = 0; k < 5; k++)
        w = (w << 8) | p[b0 + k];
    const int sh = (int)(pos & 7);
    return (u32)((w >> (40 - sh - n)) & ((n == 32) ? 0xFFFFFFFFull : ((1ull << n) - 1)));
}

// Same fetch from a window of the bit string staged in shared memory as 32-bit words whose bit
// 31 is the first bit (rel = bit offset inside the window, n <= 25 + whatever fits: two words).
__device__ __forceinline__ u32 rd_win(const u32* __restrict__ w, u32 rel, int n)
{
    const u32 i = rel >> 5;
    const u64 v = ((u64)w[i] << 32) | (u64)w[i + 1];
    return (u32)((v >> (64 - (int)(rel & 31) - n)) & ((n == 32) ? 0xFFFFFFFFull : ((1ull << n) - 1)));
}

// Pass 1: where does every chunk start?  The positions depend on every previous chunk's header
// and payload size (ANSRangeDecoder.cpp:197-213,221), so a block is one serial walk -- done by
// one warp per block: the warp stages a 512-byte window of the bit string at the current position in
// shared memory with coalesced 32-bit loads, lane 0 parses the header out of it.
#define SCAN_WIN_WORDS 128
__global__ void __launch_bounds__(32)
ans0_dec_scan_kernel(DecodeLaunch L)
{
    __shared__ u32 s_win[SCAN_WIN_WORDS + 2];
    const int b = blockIdx.x;
    const int lane = threadIdx.x;
    const int m = L.preLen[b];
    u64* cp = L.chunkPos + (i64)b * L.maxChunks;
    u64 pos = L.payStart[b];
    const u64 endBits = L.inBits[b];
    if (L.eType == E_RAW || m <= 32) {
        if (lane == 0) {
            cp[0] = pos;
            if (pos + 8ull * (u64)m > endBits)
                atomicExch(L.errFlag, KERR_BAD_STREAM);
        }
        return;
    }
    const u8* __restrict__ p = L.in + (i64)b * L.inStride;
    const u32* __restrict__ pw = reinterpret_cast<const u32*>(p); // block bit strings start 4-byte aligned
    const bool aligned = (((size_t)p) & 3) == 0;
    const u64 lastWord = (endBits + 31) >> 5; // words holding valid bits: [0, lastWord)
    const int nChunks = (m + ANS_CHUNK - 1) >> 14;
    for (int c = 0; c < nChunks; c++) {
        // stage the window [pos & ~31, + 4096 bits + 2 words): a header is < 3600 bits
        const u64 w0 = pos >> 5;
        __syncwarp();
        for (int i = lane; i < SCAN_WIN_WORDS + 2; i += 32) {
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
        __syncwarp();
        int err = 0;
        u64 next = pos;
        if (lane == 0) {
            cp[c] = pos;
            u32 rel = (u32)(pos & 31);
            if (pos + 16 > endBits) {
                err = KERR_BAD_STREAM;
            } else {
                const int lr = 8 + (int)rd_win(s_win, rel, 3);
                rel += 3;
                int asz;
                if (rd_win(s_win, rel, 1) == 0) {
                    asz = (rd_win(s_win, rel + 1, 1) == 0) ? 256 : 0;
                    rel += 2;
                } else {
                    const int last = (int)rd_win(s_win, rel + 1, 5);
                    rel += 6;
                    asz = 0;
                    for (int i = 0; i <= last; i++) {
                        asz += __popc(rd_win(s_win, rel, 8));
                        rel += 8;
                    }
                }
                if (asz == 0 || lr > ANS0_LR) {
                    err = (asz == 0) ? KERR_BAD_STREAM : KERR_UNSUPPORTED;
                } else {
                    const int chk = (asz >= 64) ? 8 : 6;
                    const int llr = ilog2_u32((u32)lr) + 1;
                    // header <= 3 + 6 + 256 + 43 * 4 + 255 * lr bits < the window
                    for (int i = 1; i < asz && !err; i += chk) {
                        const int logMax = (int)rd_win(s_win, rel, llr);
                        rel += llr;
                        if (logMax > lr)
                            err = KERR_BAD_STREAM;
                        rel += (u32)logMax * (u32)min(chk, asz - i);
                    }
                    u64 q = (pos & ~31ull) + rel;
                    if (!err && asz > 1) {
                        u32 v = rd_win(s_win, rel, 8);
                        rel += 8;
                        u32 sz = v & 0x7F;
                        for (int shift = 7; v >= 128 && shift <= 28; shift += 7) {
                            v = rd_win(s_win, rel, 8);
                            rel += 8;
                            sz |= (v & 0x7F) << shift;
                        }
                        q = (pos & ~31ull) + rel + 128 + 8ull * sz;
                    }
                    if (q > endBits)
                        err = KERR_BAD_STREAM;
                    next = q;
                }
            }
        }
        err = __shfl_sync(FULL_MASK, err, 0);
        if (err) {
            if (lane == 0)
                atomicExch(L.errFlag, err);
            return;
        }
        pos = __shfl_sync(FULL_MASK, next, 0);
    }
}

// Pass 2: one warp per 8 chunks, one quad per chunk, lane k owns state k.
// smem per chunk: 4 KiB slot->symbol table, 256 x (freq | cum << 16), and a 128-word ring of the
// chunk's renormalisation words: the quad copies the (bit-misaligned) payload into it 32 words
// at a time, shifted into place, one refill ahead of the coding loop, so a renormalisation is a
// shared-memory read instead of a global one on the state's dependency chain.
#define DEC_RING 128
__global__ void __launch_bounds__(32)
ans0_decode_kernel(DecodeLaunch L)
{
    __shared__ __align__(16) u8 s_f2s[8][1 << ANS0_LR];
    __shared__ u32 s_fc[8][256];
    __shared__ __align__(16) u32 s_ring[8][DEC_RING / 2];

    const int lane = threadIdx.x;
    const int groupsPerBlk = (L.maxChunks + 7) >> 3;
    const int b = blockIdx.x / groupsPerBlk;
    const int c0 = (blockIdx.x - b * groupsPerBlk) << 3;
    const int m = L.preLen[b];
    const u8* __restrict__ p = L.in + (i64)b * L.inStride;
    u8* __restrict__ out = L.dst + (i64)b * L.dstStride;
    const u64* cp = L.chunkPos + (i64)b * L.maxChunks;
    const int nChunks = (L.eType == E_RAW || m <= 32) ? 1 : ((m + ANS_CHUNK - 1) >> 14);
    if (c0 >= nChunks)
        return;
    if (*L.errFlag != 0)
        return; // the header walk rejected a block: its chunk positions are not trustworthy
    if (L.eType == E_RAW || m <= 32) {
        const u64 pos = cp[0];
        for (int i = lane; i < m; i += 32)
            out[i] = (u8)rd_bits(p, pos + 8ull * i, 8);
        return;
    }

    const int j = lane >> 2, k = lane & 3;
    const int c = c0 + j;
    const bool valid = c < nChunks;
    const int len = valid ? min(ANS_CHUNK, m - c * ANS_CHUNK) : 0;
    u8* __restrict__ o = out + (i64)(valid ? c : 0) * ANS_CHUNK;

    // ---- header parse + tables (one lane per chunk, out of a shared-memory window of the bit string
    // that the whole warp stages first: the parse is a chain of ~300 dependent small reads)
    int asz = 0, lr = ANS0_LR, single = 0;
    u64 pos = 0;
    u64 hbase = 0; // bit position of s_hwin[j][0]
    {
        const u32* __restrict__ pwh = reinterpret_cast<const u32*>(p);
        const bool alh = (((size_t)p) & 3) == 0;
        const u64 lastW = (L.inBits[b] + 31) >> 5;
        const u64 mypos = (valid && k == 0) ? cp[c] : 0;
        hbase = (mypos >> 5) << 5;
        for (int jj = 0; jj < 8; jj++) {
            const u64 w0 = __shfl_sync(FULL_MASK, mypos, 4 * jj) >> 5;
            for (int i = lane; i < SCAN_WIN_WORDS + 2; i += 32) {
                const u64 wi = w0 + (u64)i;
                u32 v = 0;
                if (c0 + jj < nChunks && wi < lastW) {
                    if (alh) {
                        v = bswap32(__ldg(&pwh[wi]));
                    } else {
                        const u8* q = p + wi * 4;
                        v = ((u32)q[0] << 24) | ((u32)q[1] << 16) | ((u32)q[2] << 8) | (u32)q[3];
                    }
                }
                reinterpret_cast<u32*>(s_f2s[jj])[i] = v; // the window lives in the (not yet built) slot table
            }
        }
        __syncwarp();
    }
#define HRD(POS, N) rd_win(reinterpret_cast<const u32*>(s_f2s[j]), (u32)((POS) - hbase), (N))
    u32 psz = 0;
    u32 st0 = 0, st1 = 0, st2 = 0, st3 = 0;
    if (valid && k == 0) {
        pos = cp[c];
        lr = 8 + (int)HRD(pos, 3);
        pos += 3;
        u32 pm[8];
        if (HRD(pos, 1) == 0) {
            pos += 2; // "00": the scan pass rejected "01"
            for (int i = 0; i < 8; i++)
                pm[i] = 0xFFFFFFFFu;
            asz = 256;
        } else {
            const int last = (int)HRD(pos + 1, 5);
            pos += 6;
            for (int i = 0; i < 8; i++)
                pm[i] = 0;
            for (int i = 0; i <= last; i++) {
                const u32 mk = HRD(pos, 8);
                pos += 8;
                pm[i >> 2] |= mk << (8 * (i & 3)); // bit (8i+j) = symbol 8i+j
                asz += __popc(mk);
            }
        }
        u32* fc = s_fc[j];
        for (int i = 0; i < 256; i++)
            fc[i] = 0;
        const int chk = (asz >= 64) ? 8 : 6;
        const int llr = ilog2_u32((u32)lr) + 1;
        const u32 scale = 1u << lr;
        // first present symbol
        int sym = 0;
        while (!((pm[sym >> 5] >> (sym & 31)) & 1))
            sym++;
        const int firstSym = sym;
        sym++;
        u32 sum = 0;
        int left = asz - 1;
        bool bad = false;
        while (left > 0) {
            const int cnt = min(left, chk);
            const int logMax = (int)HRD(pos, llr);
            pos += llr;
            int got = 0;
            while (got < cnt) {
                if ((pm[sym >> 5] >> (sym & 31)) & 1) {
                    u32 fr = 1;
                    if (logMax != 0) {
                        fr = HRD(pos, logMax) + 1;
                        pos += logMax;
                    }
                    if (fr >= scale)
                        bad = true;
                    fc[sym] = fr;
                    sum += fr;
                    got++;
                }
                sym++;
            }
            left -= cnt;
        }
        if (bad || sum >= scale) {
            atomicExch(L.errFlag, KERR_BAD_STREAM);
            asz = 0;
        } else {
            fc[firstSym] = scale - sum;
            if (asz == 1) {
                single = 1 + firstSym;
            } else {
                u32 v = HRD(pos, 8);
                pos += 8;
                psz = v & 0x7F;
                for (int shift = 7; v >= 128 && shift <= 28; shift += 7) {
                    v = HRD(pos, 8);
                    pos += 8;
                    psz |= (v & 0x7F) << shift;
                }
                st0 = HRD(pos, 32);
                st1 = HRD(pos + 32, 32);
                st2 = HRD(pos + 64, 32);
                st3 = HRD(pos + 96, 32);
                pos += 128;
            }
        }
    }
#undef HRD
    __syncwarp();
    const int qlead = lane & ~3;
    asz = __shfl_sync(FULL_MASK, asz, qlead);
    single = __shfl_sync(FULL_MASK, single, qlead);
    lr = __shfl_sync(FULL_MASK, lr, qlead);
    psz = __shfl_sync(FULL_MASK, psz, qlead);
    pos = __shfl_sync(FULL_MASK, pos, qlead);
    u32 state;
    {
        const u32 a = __shfl_sync(FULL_MASK, st0, qlead), b1 = __shfl_sync(FULL_MASK, st1, qlead);
        const u32 c2 = __shfl_sync(FULL_MASK, st2, qlead), d3 = __shfl_sync(FULL_MASK, st3, qlead);
        state = (k == 0) ? a : (k == 1) ? b1 : (k == 2) ? c2 : d3;
    }
    const bool active = valid && asz > 1;
    // slot -> symbol table and cumulative frequencies: the quad's four lanes fill it together
    // (lane k takes the symbols k, k+4, ...; each writes its symbol's run of slots)
    {
        u32* fc = s_fc[j];
        u8* f2s = s_f2s[j];
        // exclusive prefix of the frequencies: lane k sums its 64 consecutive symbols first
        u32 part = 0;
        if (active)
            for (int i = 64 * k; i < 64 * k + 64; i++)
                part += fc[i];
        u32 run = 0;
#pragma unroll
        for (int t = 0; t < 3; t++) {
            const u32 v = __shfl_sync(FULL_MASK, part, qlead + t);
            if (t < k)
                run += v;
        }
        if (active) {
            for (int i = 64 * k; i < 64 * k + 64; i++) {
                const u32 fr = fc[i];
                if (fr == 0)
                    continue;
                for (u32 t = 0; t < fr; t++)
                    f2s[run + t] = (u8)i;
                fc[i] = fr | (run << 16);
                run += fr;
            }
        }
    }
    __syncwarp();

    // single-symbol chunks: fill (ANSRangeDecoder.cpp:203-205)
    if (single) {
        for (int i = k; i < len; i += 4)
            o[i] = (u8)(single - 1);
    }

    const int count4 = len & ~3;
    const int steps = active ? (count4 >> 2) : 0;
    int maxSteps = steps;
#pragma unroll
    for (int x = 16; x > 0; x >>= 1)
        maxSteps = max(maxSteps, __shfl_xor_sync(FULL_MASK, maxSteps, x));
    const u32 mask = (1u << lr) - 1;
    const u8* f2s = s_f2s[j];
    const u32* fc = s_fc[j];

    // ---- renormalisation ring.  Word w of the payload = bits [pos + 16 w, + 16) of the block's bit string.
    // A refill = 32 words = 64 bytes; lane k produces bytes [16 k, 16 k + 16) of it from five aligned
    // source words, funnel-shifted by the payload's misalignment; ring word pairs are stored first
    // word in the high half, so word w is the 16-bit value at index w ^ 1.  The loads of a refill
    // are issued one period (8 steps) before their data is shifted into the ring.
    const u32* __restrict__ pw = reinterpret_cast<const u32*>(p);
    const bool aligned = (((size_t)p) & 3) == 0;
    const u32 lastWord = (u32)((L.inBits[b] + 31) >> 5);
    u32* ring = s_ring[j];
    const u16* ring16 = reinterpret_cast<const u16*>(ring);
    u32 filled = 0; // words staged so far (multiple of 32)
    const u32 wordsTotal = active ? (psz >> 1) : 0;
    const int fsh = (int)((pos + 128ull * (u32)k) & 31);     // the same for every refill (64 bytes apart)
    u32 srcw = (u32)((pos + 128ull * (u32)k) >> 5);          // first source word of this lane's next refill
    u32 pend[5];
    bool havePending = false;
    auto issue = [&]() {
#pragma unroll
        for (int t = 0; t < 5; t++) {
            const u32 wi = srcw + (u32)t;
            u32 v = 0;
            if (wi < lastWord) {
                if (aligned) {
                    v = __ldg(&pw[wi]);
                } else {
                    const u8* q = p + (u64)wi * 4;
                    v = (u32)q[0] | ((u32)q[1] << 8) | ((u32)q[2] << 16) | ((u32)q[3] << 24);
                }
            }
            pend[t] = v; // memory byte order; swapped when it lands in the ring
        }
        srcw += 16;
    };
    auto store = [&]() {
        const u32 w0 = bswap32(pend[0]), w1 = bswap32(pend[1]), w2 = bswap32(pend[2]), w3 = bswap32(pend[3]),
                  w4 = bswap32(pend[4]);
        uint4 v;
        v.x = __funnelshift_l(w1, w0, fsh);
        v.y = __funnelshift_l(w2, w1, fsh);
        v.z = __funnelshift_l(w3, w2, fsh);
        v.w = __funnelshift_l(w4, w3, fsh);
        *reinterpret_cast<uint4*>(ring + (((filled & (DEC_RING - 1)) >> 1) + 4 * k)) = v;
        filled += 32;
    };
    // initial fill: the whole ring
    for (int r = 0; r < DEC_RING / 32; r++) {
        if (active) {
            issue();
            store();
        }
    }
    __syncwarp();

    u32 cnt = 0; // 16-bit words consumed by the quad so far
    // every 8 steps (<= 32 words consumed): land the refill issued last time, issue the next one
#define DEC_REFILL()                                                                             \
    {                                                                                            \
        if (havePending) {                                                                       \
            store();                                                                             \
            havePending = false;                                                                 \
        }                                                                                        \
        if (active && filled - cnt <= DEC_RING - 32 && filled < wordsTotal + 32) {               \
            issue();                                                                             \
            havePending = true;                                                                  \
        }                                                                                        \
        __syncwarp();                                                                            \
    }
    // one step; LIVE masks lanes whose chunk has no step here (their table reads stay in bounds)
#define DEC_STEP(LIVE, OUT)                                                                      \
    {                                                                                            \
        const u32 slot_ = state & mask;                                                          \
        const u32 cur_ = f2s[slot_];                                                             \
        const u32 e_ = fc[cur_];                                                                 \
        const u32 nst_ = (e_ & 0xFFFF) * (state >> lr) + slot_ - (e_ >> 16);                     \
        const bool need_ = (LIVE) && nst_ < (1u << 15);                                          \
        if (LIVE) {                                                                              \
            state = nst_;                                                                        \
            (OUT) = (u8)cur_; /* st3 -> i, st2 -> i+1, st1 -> i+2, st0 -> i+3 */                 \
        }                                                                                        \
        const u32 bal_ = __ballot_sync(FULL_MASK, need_);                                        \
        if (need_) { /* consumption order inside a step: st3, st2, st1, st0 (ANSRangeDecoder.cpp:245-258) */ \
            const u32 idx_ = cnt + (u32)__popc(bal_ & mAbove);                                   \
            state = (state << 16) | (u32)ring16[(idx_ & (DEC_RING - 1)) ^ 1];                    \
        }                                                                                        \
        cnt += (u32)__popc(bal_ & mQuad);                                                        \
    }
    const u32 mQuad = 0xFu << qlead;
    const u32 mAbove = (0xFu << (k + 1) & 0xFu) << qlead; // states consumed before this lane's in a step
    int minSteps = active ? steps : 0x7FFFFFFF;
#pragma unroll
    for (int x = 16; x > 0; x >>= 1)
        minSteps = min(minSteps, __shfl_xor_sync(FULL_MASK, minSteps, x));
    const int fast = (maxSteps > 0) ? (minSteps & ~7) : 0; // steps every active quad takes
    u8* op = o + (3 - k);
    int s = 0;
    for (; s < fast; s += 8) {
        DEC_REFILL()
#pragma unroll
        for (int x = 0; x < 8; x++)
            DEC_STEP(active, op[4 * x])
        op += 32;
    }
    for (; s < maxSteps; s++) {
        if ((s & 7) == 0)
            DEC_REFILL()
        DEC_STEP(s < steps, op[0])
        op += 4;
    }
#undef DEC_STEP
#undef DEC_REFILL
    if (active && k == 0) {
        const int tail = len & 3;
        // the tail bytes follow the consumed words: only read them when the payload size agrees
        // (a corrupted payload may report any word count; the scan bounded pos + 8*psz, not cnt)
        if (2 * cnt + (u32)tail != psz) {
            atomicExch(L.errFlag, KERR_BAD_STREAM);
        } else {
            for (int t = 0; t < tail; t++)
                o[count4 + t] = (u8)rd_bits(p, pos + 16ull * cnt + 8ull * t, 8);
        }
    }
}

void launch_entropy_decode(const DecodeLaunch& L, cudaStream_t s, u64* launches)
{
    if (L.eType == E_HUF) {
        launch_huffman_decode(L, s, launches);
        return;
    }
    if (L.eType == E_ANS1) {
        launch_ans1_decode(L, s, launches);
        return;
    }
    if (L.eType == E_FPAQ) {
        launch_fpaq_decode(L, s, launches);
        return;
    }
    KLAUNCH(ans0_dec_scan_kernel, L.nBlocks, 32, s, L);
    const int groups = (L.maxChunks + 7) / 8;
    if (L.evK0)
        cudaEventRecord(L.evK0, s);
    KLAUNCH(ans0_decode_kernel, L.nBlocks * groups, 32, s, L);
    if (L.evK1)
        cudaEventRecord(L.evK1, s);
    *launches += 2;
}
