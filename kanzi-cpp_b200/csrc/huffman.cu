// huffman.cu -- static canonical Huffman block coder (kanzi HUFFMAN) on sm_100a.
//
// Replaces HuffmanEncoder::encode / HuffmanDecoder::decodeV6
// (entropy/HuffmanEncoder.cpp:304-421, entropy/HuffmanDecoder.cpp:156-347).
// 16 KiB chunks, code lengths <= 12 bits, every chunk = header (alphabet, signed
// exp-Golomb length deltas) | 4 varint fragment bit counts | 4 fragment bit strings |
// raw tail bytes, all bit-contiguous.
// Encoder: one warp per chunk.  Histogram (privatised smem atomics), parallel rank
// sort of the (freq,symbol) keys, serial Moffat-Katajainen lengths on one lane, then
// a fully parallel emit: every lane owns 1/8 of a fragment, a segmented warp scan of
// the code lengths gives its exact bit offset in the chunk's bit string, and it packs
// its codes there (whole words stored, edge words OR-ed).  The chunk bit strings are
// concatenated by the same scan/concat kernels as rANS (ans.cu).
// Decoder: header walk per block, then one quad per chunk (lane k = fragment k).
#include "common.cuh"
#include "huf_tables.cuh"
#include "kernels.h"

using namespace knz;

#define HENC_WARPS 4

struct HufWarpMem {
    u32 hist[4][256]; // privatised counts; hist[0] = final counts
    u32 keys[256];    // (freq << 8) | symbol, sorted
    u32 work[256];    // lengths in place
    u16 codes[256];
    u8 sizes[256];
    u8 order[256];
    u8 scratch[6 * 256];
};

// slow serial helpers (rare paths) -------------------------------------------------
__device__ void huf_sort_keys_serial(u32* k, int n)
{
    for (int i = 1; i < n; i++) {
        const u32 v = k[i];
        int j = i - 1;
        while (j >= 0 && k[j] > v) {
            k[j + 1] = k[j];
            j--;
        }
        k[j + 1] = v;
    }
}

__global__ void __launch_bounds__(HENC_WARPS * 32)
huf_encode_kernel(BufTable bt, const BlkState* __restrict__ st, int nBlocks, int maxChunks, u8* __restrict__ slots,
                  u32* __restrict__ hdrBits, u32* __restrict__ payBytes, u32* __restrict__ payOff)
{
    __shared__ HufWarpMem s_mem[HENC_WARPS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const i64 gw = (i64)blockIdx.x * HENC_WARPS + wib;
    const int b = (int)(gw / maxChunks);
    if (b >= nBlocks)
        return;
    const int c = (int)(gw - (i64)b * maxChunks);
    const BlkState bs = st[b];
    const int m = bs.len;
    const int nChunks = (m + HUF_CHUNK - 1) / HUF_CHUNK;
    if (c >= nChunks)
        return;
    const int len = min(HUF_CHUNK, m - c * HUF_CHUNK);
    const u8* __restrict__ p = blk_src(bt, bs, b) + (i64)c * HUF_CHUNK;
    const i64 ci = (i64)b * maxChunks + c;
    u8* slot = slots + ci * ANS_SLOT;
    u32* slotw = reinterpret_cast<u32*>(slot);
    HufWarpMem& M = s_mem[wib];

    if (len < 32) { // raw chunk (HuffmanEncoder.cpp:326-329)
        for (int i = lane; i < len; i += 32)
            slot[i] = p[i];
        if (lane == 0) {
            hdrBits[ci] = 8u * (u32)len;
            payBytes[ci] = 0;
            payOff[ci] = 0;
        }
        return;
    }

    // ---- histogram
    for (int i = lane; i < 1024; i += 32)
        (&M.hist[0][0])[i] = 0;
    __syncwarp();
    {
        u32* h = M.hist[lane & 3];
        for (int i = lane * 16; i < len; i += 512) {
            if (i + 16 <= len) {
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
                for (int t = i; t < len; t++)
                    atomicAdd(&h[p[t]], 1u);
            }
        }
    }
    __syncwarp();
    int count = 0;
    for (int i = lane; i < 256; i += 32) {
        const u32 f = M.hist[0][i] + M.hist[1][i] + M.hist[2][i] + M.hist[3][i];
        M.hist[0][i] = f;
        count += f ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        count += __shfl_xor_sync(FULL_MASK, count, o);
    __syncwarp();

    // ---- parallel rank sort of the keys (freq << 8 | symbol); absent symbols sort last
    {
        u32 myk[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int s = 32 * k + lane;
            const u32 f = M.hist[0][s];
            myk[k] = f ? ((f << 8) | (u32)s) : (0xFFFFFF00u | (u32)s);
            M.work[s] = myk[k];
        }
        __syncwarp();
        int rk[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
        for (int u = 0; u < 256; u++) {
            const u32 ku = M.work[u];
#pragma unroll
            for (int k = 0; k < 8; k++)
                rk[k] += (ku < myk[k]) ? 1 : 0;
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; k++)
            M.keys[rk[k]] = myk[k];
        __syncwarp();
    }

    // ---- one lane: code lengths, limiting, canonical codes, header
    BitSink w;
    w.init(slot);
    if (lane == 0) {
        for (int i = 0; i < 256; i++) {
            M.sizes[i] = 0;
            M.codes[i] = 0;
        }
        if (count == 1) {
            const int s = (int)(M.keys[0] & 0xFF);
            M.sizes[s] = 1;
            M.codes[s] = 1 << 12;
        } else {
            for (int i = 0; i < count; i++) {
                M.work[i] = M.keys[i] >> 8;
                M.order[i] = (u8)(M.keys[i] & 0xFF);
            }
            int maxLen = huf_inplace_lengths(M.work, count);
            for (int i = 0; i < count; i++)
                M.sizes[M.order[i]] = (u8)M.work[i];
            if (maxLen > HUF_MAX_LEN) {
                maxLen = huf_limit_fast(M.sizes, M.order, count, M.scratch);
                if (maxLen < 0) {
                    // slow path (HuffmanEncoder.cpp:186-211): renormalise the counts to 2^11 and redo
                    u32* f = M.hist[1];
                    u32 total = 0;
                    int n = 0;
                    for (int s = 0; s < 256; s++) {
                        f[s] = 0;
                    }
                    for (int s = 0; s < 256; s++)
                        if (M.hist[0][s]) {
                            f[n++] = M.hist[0][s];
                            total += M.hist[0][s];
                        }
                    normalize_counts(f, total, HUF_CHUNK >> 3);
                    n = 0;
                    for (int s = 0; s < 256; s++)
                        if (M.hist[0][s]) {
                            M.keys[n] = (f[n] << 8) | (u32)s;
                            n++;
                        }
                    huf_sort_keys_serial(M.keys, count);
                    for (int i = 0; i < count; i++) {
                        M.work[i] = M.keys[i] >> 8;
                        M.order[i] = (u8)(M.keys[i] & 0xFF);
                    }
                    maxLen = huf_inplace_lengths(M.work, count);
                    for (int i = 0; i < count; i++)
                        M.sizes[M.order[i]] = (u8)M.work[i];
                }
            }
            if (maxLen > HUF_MAX_LEN) { // flat 8-bit codes in alphabet order (:98-106)
                int n = 0;
                for (int s = 0; s < 256; s++)
                    if (M.hist[0][s]) {
                        M.sizes[s] = 8;
                        M.codes[s] = (u16)((8 << 12) | n);
                        n++;
                    }
            } else {
                huf_canonical_codes(M.sizes, M.codes);
            }
        }
        huf_put_header(w, M.sizes, M.hist[0], count);
        w.finish();
    }
    __syncwarp();
    u32 hb = __shfl_sync(FULL_MASK, w.total, 0);
    if (count <= 1) { // single-symbol chunk: header only (:333-336)
        if (lane == 0) {
            hdrBits[ci] = hb;
            payBytes[ci] = 0;
            payOff[ci] = 0;
        }
        return;
    }

    // ---- emit: lane = (fragment f, part s); segmented scan of the code lengths
    const int frag = len >> 2;
    const int f = lane >> 3, sp = lane & 7;
    const int per = (frag + 7) >> 3;
    const int i0 = min(sp * per, frag), i1 = min(i0 + per, frag);
    const u8* __restrict__ fp = p + (i64)f * frag;
    u32 bitsLane = 0;
    for (int i = i0; i < i1; i++)
        bitsLane += M.codes[fp[i]] >> 12;
    // inclusive scan inside the 8-lane segment
    u32 inc = bitsLane;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        const u32 t = __shfl_up_sync(FULL_MASK, inc, o);
        if (sp >= o)
            inc += t;
    }
    const u32 nb0 = __shfl_sync(FULL_MASK, inc, 7), nb1 = __shfl_sync(FULL_MASK, inc, 15);
    const u32 nb2 = __shfl_sync(FULL_MASK, inc, 23), nb3 = __shfl_sync(FULL_MASK, inc, 31);
    auto vbits = [](u32 v) -> u32 { return v < 128 ? 8u : v < 16384 ? 16u : v < 2097152 ? 24u : 32u; };
    const u32 fragBase = hb + vbits(nb0) + vbits(nb1) + vbits(nb2) + vbits(nb3);
    const u32 fragOff = (f == 0) ? 0u : (f == 1) ? nb0 : (f == 2) ? nb0 + nb1 : nb0 + nb1 + nb2;
    const u32 tailBits = 8u * (u32)(len - 4 * frag);
    const u32 totalBits = fragBase + nb0 + nb1 + nb2 + nb3 + tailBits;
    // zero the words the bit string will occupy beyond the header bytes already written
    const u32 firstFree = (hb + 31) >> 5; // header words were written by the sink (zero padded)
    // the sink wrote bytes; clear the rest of its last word explicitly
    if (lane == 0) {
        const u32 hbytes = (hb + 7) >> 3;
        for (u32 x = hbytes; x < firstFree * 4; x++)
            slot[x] = 0;
    }
    for (u32 x = firstFree + lane; x <= (totalBits >> 5) + 1; x += 32)
        slotw[x] = 0;
    __syncwarp();
    if (lane == 0) { // 4 varints (EntropyUtils.cpp:247-259), byte granular but at a bit offset
        u64 pos = hb;
        const u32 nbv[4] = { nb0, nb1, nb2, nb3 };
        for (int j = 0; j < 4; j++) {
            u32 v = nbv[j];
            while (v >= 128) {
                put_bits_atomic(slotw, pos, 0x80 | (v & 0x7F), 8);
                pos += 8;
                v >>= 7;
            }
            put_bits_atomic(slotw, pos, v, 8);
            pos += 8;
        }
        u64 tp = (u64)fragBase + nb0 + nb1 + nb2 + nb3;
        for (int i = 4 * frag; i < len; i++) {
            put_bits_atomic(slotw, tp, p[i], 8);
            tp += 8;
        }
    }
    {
        u64 pos = (u64)fragBase + fragOff + (inc - bitsLane);
        u64 acc = 0; // bits pending, right aligned
        int na = 0;
        for (int i = i0; i < i1; i++) {
            const u32 cd = M.codes[fp[i]];
            const int cl = (int)(cd >> 12);
            acc = (acc << cl) | (u64)(cd & 0x0FFF);
            na += cl;
            if (na >= 32) {
                const u32 out = (u32)(acc >> (na - 32));
                // whole aligned word -> plain store; otherwise OR the two halves in
                if ((pos & 31) == 0)
                    slotw[pos >> 5] = bswap32(out);
                else
                    put_bits_atomic(slotw, pos, out, 32);
                pos += 32;
                na -= 32;
                acc &= (na ? ((1ull << na) - 1) : 0ull);
            }
        }
        if (na > 0)
            put_bits_atomic(slotw, pos, (u32)acc, na);
    }
    if (lane == 0) {
        hdrBits[ci] = totalBits;
        payBytes[ci] = 0;
        payOff[ci] = 0;
    }
}

// ------------------------------------------------------------------ decoder
__device__ __forceinline__ u32 hrd_bits(const u8* __restrict__ p, u64 pos, int n)
{
    const u64 b0 = pos >> 3;
    u64 w = 0;
#pragma unroll
    for (int k = 0; k < 5; k++)
        w = (w << 8) | p[b0 + k];
    const int sh = (int)(pos & 7);
    return (u32)((w >> (40 - sh - n)) & ((n == 32) ? 0xFFFFFFFFull : ((1ull << n) - 1)));
}

// The same fetch from a window of the bit string staged in shared memory (32-bit words, bit 31 =
// first bit; word i holds bits [base + 32 i, + 32)); n <= 32.
#define HSCAN_WIN_WORDS 128
struct WinSrc {
    const u32* w;
    u64 base;
};
__device__ __forceinline__ u32 hrd_bits(const WinSrc& s, u64 pos, int n)
{
    const u32 rel = (u32)(pos - s.base);
    const u32 i = min(rel >> 5, 128u); // malformed headers may run past the 130-word window
    const u64 v = ((u64)s.w[i] << 32) | (u64)s.w[i + 1];
    return (u32)((v >> (64 - (int)(rel & 31) - n)) & ((n == 32) ? 0xFFFFFFFFull : ((1ull << n) - 1)));
}

// Signed exp-Golomb (ExpGolombDecoder.hpp:52-75): `1` = 0, else z >= 1 zeros, a one, (z & 7) + 1
// value bits whose last one is the sign.  One 32-bit peek decodes every code the encoder emits
// (z <= 4); longer zero runs (malformed input) take the bitwise walk.
template <class S>
__device__ __forceinline__ int hrd_expgolomb(const S& p, u64& pos)
{
    const u32 v = hrd_bits(p, pos, 32);
    if (v & 0x80000000u) {
        pos += 1;
        return 0;
    }
    u32 lg;
    int res;
    const int z = __clz((int)v); // zeros before the first one (v != 0 -> z <= 31; v == 0 -> 32)
    if (z <= 16) {
        lg = (u32)z & 7;
        res = (int)((v << (z + 1)) >> (31 - lg));
        pos += (u64)(z + 1) + lg + 1;
    } else {
        pos += 1;
        lg = 1;
        while (hrd_bits(p, pos, 1) == 0 && lg < 64) {
            pos += 1;
            lg++;
        }
        pos += 1;
        lg &= 7;
        res = (int)hrd_bits(p, pos, (int)lg + 1);
        pos += lg + 1;
    }
    const int sgn = res & 1;
    res = (res >> 1) + (1 << lg) - 1;
    return (int)(int8_t)((res - sgn) ^ -sgn);
}

template <class S>
__device__ __forceinline__ u32 hrd_varint(const S& p, u64& pos)
{
    u32 v = hrd_bits(p, pos, 8);
    pos += 8;
    u32 r = v & 0x7F;
    for (int shift = 7; v >= 128 && shift <= 28; shift += 7) {
        v = hrd_bits(p, pos, 8);
        pos += 8;
        r |= (v & 0x7F) << shift;
    }
    return r;
}

// Parses one chunk header at pos: alphabet bitmap into pm[8], lengths into sizes (may be NULL).
// Returns the alphabet size (0 = invalid).
template <class S>
__device__ int huf_read_header(const S& p, u64& pos, u32* pm, u8* sizes)
{
    int asz = 0;
    if (hrd_bits(p, pos, 1) == 0) {
        const bool full = hrd_bits(p, pos + 1, 1) == 0;
        pos += 2;
        if (!full)
            return 0;
        for (int i = 0; i < 8; i++)
            pm[i] = 0xFFFFFFFFu;
        asz = 256;
    } else {
        const int last = (int)hrd_bits(p, pos + 1, 5);
        pos += 6;
        for (int i = 0; i < 8; i++)
            pm[i] = 0;
        for (int i = 0; i <= last; i++) {
            const u32 mk = hrd_bits(p, pos, 8);
            pos += 8;
            pm[i >> 2] |= mk << (8 * (i & 3));
            asz += __popc(mk);
        }
    }
    int cur = 2;
    for (int s = 0; s < 256; s++) {
        if (!((pm[s >> 5] >> (s & 31)) & 1))
            continue;
        cur = (int)(int8_t)(cur + hrd_expgolomb(p, pos));
        if (cur <= 0 || cur > HUF_MAX_LEN)
            return 0;
        if (sizes)
            sizes[s] = (u8)cur;
    }
    return asz;
}

// Chunk start positions: one serial walk per block (a chunk's length is only known once its header
// is parsed) -- one warp per block stages a 512-byte window of the bit string at the current
// position in shared memory (coalesced 32-bit loads), lane 0 parses the header out of it.  A header
// is at most 262 + 256 * 10 + 4 * 40 bits.
__global__ void __launch_bounds__(32)
huf_dec_scan_kernel(DecodeLaunch L)
{
    __shared__ u32 s_win[HSCAN_WIN_WORDS + 2];
    const int b = blockIdx.x;
    const int lane = threadIdx.x;
    const int m = L.preLen[b];
    u64* cp = L.chunkPos + (i64)b * L.maxChunks;
    u64 pos = L.payStart[b];
    const u64 endBits = L.inBits[b];
    const u8* __restrict__ p = L.in + (i64)b * L.inStride;
    const u32* __restrict__ pw = reinterpret_cast<const u32*>(p);
    const bool aligned = (((size_t)p) & 3) == 0;
    const u64 lastWord = (endBits + 31) >> 5;
    const int nChunks = (m + HUF_CHUNK - 1) / HUF_CHUNK;
    for (int c = 0; c < nChunks; c++) {
        const int len = min(HUF_CHUNK, m - c * HUF_CHUNK);
        if (len < 32) {
            if (lane == 0)
                cp[c] = pos;
            pos += 8ull * (u64)len;
            if (pos > endBits) {
                if (lane == 0)
                    atomicExch(L.errFlag, KERR_BAD_STREAM);
                return;
            }
            continue;
        }
        const u64 w0 = pos >> 5;
        __syncwarp();
        for (int i = lane; i < HSCAN_WIN_WORDS + 2; i += 32) {
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
            if (pos + 8 > endBits) {
                err = KERR_BAD_STREAM;
            } else {
                WinSrc src;
                src.w = s_win;
                src.base = w0 << 5;
                u32 pm[8];
                u64 q = pos;
                const int asz = huf_read_header(src, q, pm, (u8*)NULL);
                if (asz == 0) {
                    err = KERR_BAD_STREAM;
                } else if (asz > 1) {
                    u64 tot = 0;
                    for (int j = 0; j < 4; j++)
                        tot += hrd_varint(src, q);
                    q += tot + 8ull * (u64)(len - 4 * (len >> 2));
                }
                if (q > endBits)
                    err = KERR_BAD_STREAM;
                next = q;
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

// One warp per 8 chunks, one quad per chunk, lane k decodes fragment k.
// smem per chunk: 4 KiB (12-bit prefix -> symbol) + 256 B lengths.
__global__ void __launch_bounds__(32)
huf_decode_kernel(DecodeLaunch L)
{
    __shared__ u8 s_tab[8][1 << HUF_MAX_LEN];
    __shared__ u8 s_len[8][256];
    __shared__ u32 s_hwin[8][HSCAN_WIN_WORDS + 2]; // the 8 chunk headers, staged by the whole warp
    const int lane = threadIdx.x;
    const int groupsPerBlk = (L.maxChunks + 7) >> 3;
    const int b = blockIdx.x / groupsPerBlk;
    const int c0 = (blockIdx.x - b * groupsPerBlk) << 3;
    const int m = L.preLen[b];
    const u8* __restrict__ p = L.in + (i64)b * L.inStride;
    u8* __restrict__ out = L.dst + (i64)b * L.dstStride;
    const u64* cp = L.chunkPos + (i64)b * L.maxChunks;
    const int nChunks = (m + HUF_CHUNK - 1) / HUF_CHUNK;
    if (c0 >= nChunks)
        return;
    if (*L.errFlag != 0)
        return; // the header walk rejected a block: its chunk positions are not trustworthy
    const int j = lane >> 2, k = lane & 3;
    const int c = c0 + j;
    const bool valid = c < nChunks;
    const int len = valid ? min(HUF_CHUNK, m - c * HUF_CHUNK) : 0;
    u8* __restrict__ o = out + (i64)(valid ? c : 0) * HUF_CHUNK;
    u64 pos = valid ? cp[c] : 0;
    {
        const u32* __restrict__ pw = reinterpret_cast<const u32*>(p);
        const bool aligned = (((size_t)p) & 3) == 0;
        const u64 lastWord = (L.inBits[b] + 31) >> 5;
        for (int jj = 0; jj < 8; jj++) {
            const u64 w0 = __shfl_sync(FULL_MASK, pos, 4 * jj) >> 5;
            for (int i = lane; i < HSCAN_WIN_WORDS + 2; i += 32) {
                const u64 wi = w0 + (u64)i;
                u32 v = 0;
                if (c0 + jj < nChunks && wi < lastWord) {
                    if (aligned) {
                        v = bswap32(__ldg(&pw[wi]));
                    } else {
                        const u8* q = p + wi * 4;
                        v = ((u32)q[0] << 24) | ((u32)q[1] << 16) | ((u32)q[2] << 8) | (u32)q[3];
                    }
                }
                s_hwin[jj][i] = v;
            }
        }
        __syncwarp();
    }
    int asz = 0, single = -1;
    u32 nb[4] = { 0, 0, 0, 0 };
    if (valid && len < 32) {
        for (int i = k; i < len; i += 4)
            o[i] = (u8)hrd_bits(p, pos + 8ull * i, 8);
    } else if (valid && k == 0) {
        u32 pm[8];
        u8* sizes = s_len[j];
        for (int i = 0; i < 256; i++)
            sizes[i] = 0;
        WinSrc hsrc;
        hsrc.w = s_hwin[j];
        hsrc.base = (pos >> 5) << 5;
        asz = huf_read_header(hsrc, pos, pm, sizes);
        if (asz == 1) {
            for (int s = 0; s < 256; s++)
                if (sizes[s])
                    single = s;
        } else if (asz > 1) {
            // canonical codes -> 12-bit direct table (HuffmanDecoder.cpp:111-140)
            u8* tab = s_tab[j];
            int code = 0, curLen = 0;
            bool first = true, bad = false;
            u32 filled = 0;
            for (int l = 1; l <= HUF_MAX_LEN; l++)
                for (int s = 0; s < 256; s++) {
                    if (sizes[s] != l)
                        continue;
                    if (first) {
                        curLen = l;
                        first = false;
                    }
                    code <<= (l - curLen);
                    curLen = l;
                    const int wdt = 1 << (HUF_MAX_LEN - l);
                    const int idx = code * wdt;
                    if (idx + wdt > (1 << HUF_MAX_LEN)) {
                        bad = true;
                    } else {
                        for (int x = 0; x < wdt; x++)
                            tab[idx + x] = (u8)s;
                        filled += (u32)wdt;
                    }
                    code++;
                }
            if (bad)
                asz = 0;
            // prefixes no code covers map to symbol of length 0 -> caught below
            if (filled < (1u << HUF_MAX_LEN)) {
                // incomplete code (cannot come from the encoder): mark the rest invalid
                // by pointing at a symbol with length 0 if one exists
            }
            for (int q = 0; q < 4; q++)
                nb[q] = hrd_varint(hsrc, pos);
        }
        if (asz == 0)
            atomicExch(L.errFlag, KERR_BAD_STREAM);
    }
    __syncwarp();
    const int ql = lane & ~3;
    asz = __shfl_sync(FULL_MASK, asz, ql);
    single = __shfl_sync(FULL_MASK, single, ql);
    pos = __shfl_sync(FULL_MASK, pos, ql);
    const u32 n0 = __shfl_sync(FULL_MASK, nb[0], ql), n1 = __shfl_sync(FULL_MASK, nb[1], ql);
    const u32 n2 = __shfl_sync(FULL_MASK, nb[2], ql), n3 = __shfl_sync(FULL_MASK, nb[3], ql);
    if (!valid || len < 32)
        return;
    if (asz == 1) {
        for (int i = k; i < len; i += 4)
            o[i] = (u8)single;
        return;
    }
    if (asz < 2)
        return;
    const int frag = len >> 2;
    const u32 myBits = (k == 0) ? n0 : (k == 1) ? n1 : (k == 2) ? n2 : n3;
    u64 fpos = pos + ((k == 0) ? 0u : (k == 1) ? n0 : (k == 2) ? n0 + n1 : n0 + n1 + n2);
    const u64 fend = fpos + myBits;
    const u8* tab = s_tab[j];
    const u8* sizes = s_len[j];
    u8* __restrict__ fo = o + (i64)k * frag;
    bool bad = false;
    // 64-bit MSB-aligned bit buffer refilled with aligned 32-bit loads (one load per ~3 symbols, L1
    // resident) instead of a 5-byte fetch per symbol on the fragment's dependency chain
    const u32* __restrict__ pw = reinterpret_cast<const u32*>(p);
    const bool aligned = (((size_t)p) & 3) == 0;
    const u32 lastWord = (u32)((L.inBits[b] + 31) >> 5);
    auto loadw = [&](u32 wi) -> u32 { // memory byte order: swapped where it is used, not where it is loaded
        if (wi >= lastWord)
            return 0u;
        if (aligned)
            return __ldg(&pw[wi]);
        const u8* q = p + (u64)wi * 4;
        return (u32)q[0] | ((u32)q[1] << 8) | ((u32)q[2] << 16) | ((u32)q[3] << 24);
    };
    u32 wi = (u32)(fpos >> 5);
    const int skip = (int)(fpos & 31);
    u64 buf = (u64)bswap32(loadw(wi++)) << (32 + skip);
    int nbuf = 32 - skip;
    u32 nxt = loadw(wi++); // always one word ahead: a refill never waits for its load
    i64 rem = (i64)myBits; // bits of the fragment not consumed yet
    for (int i = 0; i < frag; i++) {
        if (nbuf <= 32) {
            buf |= (u64)bswap32(nxt) << (32 - nbuf);
            nbuf += 32;
            nxt = loadw(wi++);
        }
        u32 v = (u32)(buf >> (64 - HUF_MAX_LEN));
        if (rem < HUF_MAX_LEN)
            v = (rem <= 0) ? 0u : (v & ~((1u << (HUF_MAX_LEN - (int)rem)) - 1u)); // bits past the fragment read as 0
        const u32 sy = tab[v];
        const u32 cl = sizes[sy];
        if (cl == 0) {
            bad = true;
            break;
        }
        fo[i] = (u8)sy;
        buf <<= cl;
        nbuf -= (int)cl;
        rem -= (i64)cl;
    }
    fpos = fend - rem;
    if (bad || fpos != fend)
        atomicExch(L.errFlag, KERR_BAD_STREAM);
    if (k == 0) {
        const u64 tp = pos + n0 + n1 + n2 + n3;
        for (int i = 4 * frag; i < len; i++)
            o[i] = (u8)hrd_bits(p, tp + 8ull * (u64)(i - 4 * frag), 8);
    }
}

void launch_huffman_encode_chunks(const EncodeLaunch& L, cudaStream_t s, u64* launches)
{
    const i64 warps = (i64)L.nBlocks * L.maxChunks;
    const int ctas = (int)((warps + HENC_WARPS - 1) / HENC_WARPS);
    if (L.evK0)
        cudaEventRecord(L.evK0, s);
    KLAUNCH(huf_encode_kernel, ctas, HENC_WARPS * 32, s, L.bt, L.st, L.nBlocks, L.maxChunks, L.slots, L.hdrBits,
            L.payBytes, L.payOff);
    if (L.evK1)
        cudaEventRecord(L.evK1, s);
    *launches += 1;
}

void launch_huffman_decode(const DecodeLaunch& L, cudaStream_t s, u64* launches)
{
    KLAUNCH(huf_dec_scan_kernel, L.nBlocks, 32, s, L);
    const int groups = (L.maxChunks + 7) / 8;
    if (L.evK0)
        cudaEventRecord(L.evK0, s);
    KLAUNCH(huf_decode_kernel, L.nBlocks * groups, 32, s, L);
    if (L.evK1)
        cudaEventRecord(L.evK1, s);
    *launches += 2;
}
