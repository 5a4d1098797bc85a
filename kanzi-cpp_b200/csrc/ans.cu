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
#include "common.cuh"
#include "ans_tables.cuh"
#include "kernels.h"

using namespace knz;

// ------------------------------------------------------------------ encoder
// smem per warp: 8 tables x 256 x 8 B = 16 KiB.  While the histograms are built the
// second KiB of every table region is free: four of them hold the 4 privatised
// copies of the chunk being counted.
#define ENC_WARPS 2

__global__ void __launch_bounds__(ENC_WARPS * 32)
ans0_encode_kernel(BufTable bt, const BlkState* __restrict__ st, int nBlocks, int maxChunks, u8* __restrict__ slots,
                   u32* __restrict__ hdrBits, u32* __restrict__ payBytes, u32* __restrict__ payOff)
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
        __syncwarp();
        u32* f = reinterpret_cast<u32*>(sym[j]);
        for (int i = lane; i < 256; i += 32)
            f[i] = hk[0][i] + hk[1][i] + hk[2][i] + hk[3][i];
        __syncwarp();
    }

    // ---- phase B: one lane per chunk: normalise, header, tables
    const int j = lane >> 2, k = lane & 3;
    const int c = c0 + j;
    const bool valid = c < nChunks;
    const int len = valid ? min(ANS_CHUNK, m - c * ANS_CHUNK) : 0;
    u8* slot = slots + ((i64)b * maxChunks + (valid ? c : 0)) * ANS_SLOT;
    BitSink w;
    w.init(slot);
    int active = 0;
    if (valid && k == 0) {
        u32* f = reinterpret_cast<u32*>(sym[j]);
        const int asz = normalize_counts(f, (u32)len, 1u << ANS0_LR);
        put_chunk_header(w, f, asz, ANS0_LR);
        if (asz > 1) {
            active = 1;
            u32 total = 0;
            for (int i = 0; i < 256; i++)
                total += f[i];
            // descending: entry i overwrites histogram words 2i, 2i+1 (both >= i, already consumed);
            // cumulative frequency of i = total - sum of the frequencies >= i
            u32 run = total;
            for (int i = 255; i >= 0; i--) {
                const u32 fr = f[i];
                run -= fr;
                sym[j][i] = (fr == 0) ? 0ull : make_enc_entry((int)run, (int)fr, ANS0_LR);
            }
        }
    }
    __syncwarp();
    active = __shfl_sync(FULL_MASK, active, lane & ~3);

    // ---- phase C: interleaved rANS, lane k of quad j owns state k
    const int end4 = len & ~3;
    const int steps = active ? (end4 >> 2) : 0;
    int maxSteps = steps;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        maxSteps = max(maxSteps, __shfl_xor_sync(FULL_MASK, maxSteps, o));

    const u32* __restrict__ words = reinterpret_cast<const u32*>(src + (i64)(valid ? c : 0) * ANS_CHUNK);
    const u64* __restrict__ tab = sym[j];
    u32 state = 1u << 15; // ANS_TOP
    u32 cnt = 0;
    u16* wend = reinterpret_cast<u16*>(slot + ANS_WEND);
    const int bsh = 8 * (3 - k);
    const u32 below = (1u << k) - 1u;
    const int qsh = lane & ~3;
    const int wtop = (end4 >> 2) - 1; // step s consumes word wtop - s (the quad's 4 bytes)
    // Software pipeline, 4 steps (16 input bytes per quad) per group, two groups ahead.
    // Full chunks are 16-byte aligned from the top, so a group is one 128-bit load.
    const bool vec = (end4 & 15) == 0;
    uint4 ga = make_uint4(0, 0, 0, 0), gb = ga, gc = ga;
    auto fetch = [&](int s0) -> uint4 {
        uint4 r = make_uint4(0, 0, 0, 0);
        if (s0 + 3 < steps && vec) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(words + (wtop - s0 - 3)));
            r = make_uint4(v.w, v.z, v.y, v.x); // .x = word of step s0
        } else {
            if (s0 < steps)
                r.x = __ldg(&words[wtop - s0]);
            if (s0 + 1 < steps)
                r.y = __ldg(&words[wtop - s0 - 1]);
            if (s0 + 2 < steps)
                r.z = __ldg(&words[wtop - s0 - 2]);
            if (s0 + 3 < steps)
                r.w = __ldg(&words[wtop - s0 - 3]);
        }
        return r;
    };
    ga = fetch(0);
    gb = fetch(4);
    for (int s0 = 0; s0 < maxSteps; s0 += 4) {
        gc = fetch(s0 + 8);
        const u32 wv[4] = { ga.x, ga.y, ga.z, ga.w };
#pragma unroll
        for (int x = 0; x < 4; x++) {
            bool did = false;
            u32 word = 0;
            if (s0 + x < steps) {
                const u32 cb = (wv[x] >> bsh) & 0xFF;
                state = enc_step(state, tab[cb], ANS0_LR, &did, &word);
            }
            const u32 qb = (__ballot_sync(FULL_MASK, did) >> qsh) & 0xF;
            if (did) // memory order [hi][lo] (ANSRangeEncoder.hpp:122-126)
                wend[-1 - (int)(cnt + __popc(qb & below))] = (u16)__byte_perm(word, 0, 0x4401);
            cnt += __popc(qb);
        }
        ga = gb;
        gb = gc;
    }

    // ---- epilogue: varint size, 4 states, tail bytes
    const u32 s1 = __shfl_sync(FULL_MASK, state, (lane & ~3) + 1);
    const u32 s2 = __shfl_sync(FULL_MASK, state, (lane & ~3) + 2);
    const u32 s3 = __shfl_sync(FULL_MASK, state, (lane & ~3) + 3);
    if (valid && k == 0) {
        const i64 ci = (i64)b * maxChunks + c;
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
    const int dataSize = (m < 256) ? 1 : (ilog2_u32((u32)m) >> 3) + 1;
    const int hdrBytes = 1 + ((nTransforms > 4) ? 1 : 0) + dataSize;
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

__global__ void block_header_kernel(const BlkState* __restrict__ st, int nBlocks, int nTransforms,
                                    u8* __restrict__ out, i64 outStride)
{
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
                     const u64* __restrict__ blockOff, u8* __restrict__ stream)
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
    const u8* src = blockOut + (i64)b * outStride;
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
void launch_entropy_encode(const EncodeLaunch& L, cudaStream_t s, u64* launches)
{
    const int nB = L.nBlocks;
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
        KLAUNCH(ans0_encode_kernel, ctas, ENC_WARPS * 32, s, L.bt, L.st, nB, L.maxChunks, L.slots, L.hdrBits,
                L.payBytes, L.payOff);
        if (L.evK1)
            cudaEventRecord(L.evK1, s);
    }
    KLAUNCH(ans_scan_kernel, nB, 256, s, L.st, nB, L.maxChunks, L.eType, L.nTransforms, L.hdrBits, L.payBytes,
                                       L.chunkOff, L.blockBits, L.out, L.outStride, L.errFlag);
    i64 zx64 = (L.outStride / 4 + 255) / 256;
    const int zx = (int)(zx64 < 64 ? zx64 : 64);
    KLAUNCH(out_prepare_kernel, dim3(zx, nB), 256, s, L.st, L.nTransforms, L.blockBits, L.out, L.outStride);
    KLAUNCH(block_header_kernel, (nB + 127) / 128, 128, s, L.st, nB, L.nTransforms, L.out, L.outStride);
    if (L.eType == E_RAW)
        KLAUNCH(ans_concat_kernel, dim3(64, nB), 128, s, L.bt, L.st, L.maxChunks, L.eType, L.slots, L.hdrBits,
                                                       L.payBytes, L.payOff, L.chunkOff, L.out, L.outStride);
    else
        KLAUNCH(ans_concat_kernel, dim3(L.maxChunks, nB), 128, s, L.bt, L.st, L.maxChunks, L.eType, L.slots,
                                                                L.hdrBits, L.payBytes, L.payOff, L.chunkOff,
                                                                L.out, L.outStride);
    *launches += 5;
}

void launch_stream_assemble(const u8* blockOut, i64 outStride, const u64* blockBits, int nBlocks,
                            const u64* startBit, u64* blockOff, u64* endBit, u8* stream, cudaStream_t s,
                            u64* launches)
{
    KLAUNCH(stream_scan_kernel, 1, 256, s, blockBits, nBlocks, startBit, blockOff, endBit);
    KLAUNCH(stream_concat_kernel, dim3(32, nBlocks), 256, s, blockOut, outStride, blockBits, blockOff, stream);
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

// Pass 1: one thread per block walks the chunk headers to find where each chunk
// starts (the positions depend on every previous chunk's header and payload
// size: ANSRangeDecoder.cpp:197-213,221).
__global__ void ans0_dec_scan_kernel(DecodeLaunch L)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= L.nBlocks)
        return;
    const int m = L.preLen[b];
    u64* cp = L.chunkPos + (i64)b * L.maxChunks;
    u64 pos = L.payStart[b];
    const u64 endBits = L.inBits[b];
    if (L.eType == E_RAW || m <= 32) {
        cp[0] = pos;
        if (pos + 8ull * (u64)m > endBits)
            atomicExch(L.errFlag, KERR_BAD_STREAM);
        return;
    }
    const u8* __restrict__ p = L.in + (i64)b * L.inStride;
    const int nChunks = (m + ANS_CHUNK - 1) >> 14;
    for (int c = 0; c < nChunks; c++) {
        cp[c] = pos;
        if (pos + 16 > endBits) {
            atomicExch(L.errFlag, KERR_BAD_STREAM);
            return;
        }
        const int lr = 8 + (int)rd_bits(p, pos, 3);
        pos += 3;
        int asz;
        if (rd_bits(p, pos, 1) == 0) {
            asz = (rd_bits(p, pos + 1, 1) == 0) ? 256 : 0;
            pos += 2;
        } else {
            const int last = (int)rd_bits(p, pos + 1, 5);
            pos += 6;
            asz = 0;
            for (int i = 0; i <= last; i++) {
                asz += __popc(rd_bits(p, pos, 8));
                pos += 8;
            }
        }
        if (asz == 0 || lr > ANS0_LR) {
            atomicExch(L.errFlag, (asz == 0) ? KERR_BAD_STREAM : KERR_UNSUPPORTED);
            return;
        }
        const int chk = (asz >= 64) ? 8 : 6;
        const int llr = ilog2_u32((u32)lr) + 1;
        for (int i = 1; i < asz; i += chk) {
            const int logMax = (int)rd_bits(p, pos, llr);
            pos += llr;
            if (logMax > lr) {
                atomicExch(L.errFlag, KERR_BAD_STREAM);
                return;
            }
            pos += (u64)logMax * (u64)min(chk, asz - i);
            if (pos > endBits) {
                atomicExch(L.errFlag, KERR_BAD_STREAM);
                return;
            }
        }
        if (asz > 1) {
            u32 v = rd_bits(p, pos, 8);
            pos += 8;
            u32 sz = v & 0x7F;
            for (int shift = 7; v >= 128 && shift <= 28; shift += 7) {
                v = rd_bits(p, pos, 8);
                pos += 8;
                sz |= (v & 0x7F) << shift;
            }
            pos += 128 + 8ull * sz;
        }
        if (pos > endBits) {
            atomicExch(L.errFlag, KERR_BAD_STREAM);
            return;
        }
    }
}

// Pass 2: one warp per 8 chunks, one quad per chunk.
// smem per chunk: 4 KiB slot->symbol table + 256 x (freq | cum<<16).
__global__ void __launch_bounds__(32)
ans0_decode_kernel(DecodeLaunch L)
{
    __shared__ __align__(16) u8 s_f2s[8][1 << ANS0_LR];
    __shared__ u32 s_fc[8][256];

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

    // ---- header parse + tables (one lane per chunk)
    int asz = 0, lr = ANS0_LR, single = 0;
    u64 pos = 0;
    u32 psz = 0;
    u32 st0 = 0, st1 = 0, st2 = 0, st3 = 0;
    if (valid && k == 0) {
        pos = cp[c];
        lr = 8 + (int)rd_bits(p, pos, 3);
        pos += 3;
        u32 pm[8];
        if (rd_bits(p, pos, 1) == 0) {
            pos += 2; // "00": the scan pass rejected "01"
            for (int i = 0; i < 8; i++)
                pm[i] = 0xFFFFFFFFu;
            asz = 256;
        } else {
            const int last = (int)rd_bits(p, pos + 1, 5);
            pos += 6;
            for (int i = 0; i < 8; i++)
                pm[i] = 0;
            for (int i = 0; i <= last; i++) {
                const u32 mk = rd_bits(p, pos, 8);
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
            const int logMax = (int)rd_bits(p, pos, llr);
            pos += llr;
            int got = 0;
            while (got < cnt) {
                if ((pm[sym >> 5] >> (sym & 31)) & 1) {
                    u32 fr = 1;
                    if (logMax != 0) {
                        fr = rd_bits(p, pos, logMax) + 1;
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
                u8* f2s = s_f2s[j];
                u32 run = 0;
                for (int i = 0; i < 256; i++) {
                    const u32 fr = fc[i];
                    if (fr == 0)
                        continue;
                    for (u32 t = 0; t < fr; t++)
                        f2s[run + t] = (u8)i;
                    fc[i] = fr | (run << 16);
                    run += fr;
                }
                u32 v = rd_bits(p, pos, 8);
                pos += 8;
                psz = v & 0x7F;
                for (int shift = 7; v >= 128 && shift <= 28; shift += 7) {
                    v = rd_bits(p, pos, 8);
                    pos += 8;
                    psz |= (v & 0x7F) << shift;
                }
                st0 = rd_bits(p, pos, 32);
                st1 = rd_bits(p, pos + 32, 32);
                st2 = rd_bits(p, pos + 64, 32);
                st3 = rd_bits(p, pos + 96, 32);
                pos += 128;
            }
        }
    }
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

    // single-symbol chunks: fill (ANSRangeDecoder.cpp:203-205)
    if (single) {
        for (int i = k; i < len; i += 4)
            o[i] = (u8)(single - 1);
    }

    const bool active = valid && asz > 1;
    const int count4 = len & ~3;
    const int steps = active ? (count4 >> 2) : 0;
    int maxSteps = steps;
#pragma unroll
    for (int x = 16; x > 0; x >>= 1)
        maxSteps = max(maxSteps, __shfl_xor_sync(FULL_MASK, maxSteps, x));
    const u32 mask = (1u << lr) - 1;
    const u8* f2s = s_f2s[j];
    const u32* fc = s_fc[j];
    u32 cnt = 0; // 16-bit words consumed by the quad so far
    for (int s = 0; s < maxSteps; s++) {
        bool need = false;
        u32 cur = 0;
        if (s < steps) {
            cur = f2s[state & mask];
            const u32 e = fc[cur];
            state = (e & 0xFFFF) * (state >> lr) + (state & mask) - (e >> 16);
            need = state < (1u << 15);
            o[4 * s + 3 - k] = (u8)cur; // st3 -> i, st2 -> i+1, st1 -> i+2, st0 -> i+3
        }
        const u32 bal = __ballot_sync(FULL_MASK, need);
        const u32 qb = (bal >> qlead) & 0xF;
        if (need) {
            // consumption order inside a step: st3, st2, st1, st0 (ANSRangeDecoder.cpp:245-258)
            const u32 idx = cnt + __popc(qb >> (k + 1));
            state = (state << 16) | rd_bits(p, pos + 16ull * idx, 16);
        }
        cnt += __popc(qb);
    }
    if (active && k == 0) {
        const int tail = len & 3;
        for (int t = 0; t < tail; t++)
            o[count4 + t] = (u8)rd_bits(p, pos + 16ull * cnt + 8ull * t, 8);
        if (2 * cnt + (u32)tail != psz)
            atomicExch(L.errFlag, KERR_BAD_STREAM);
    }
}

void launch_entropy_decode(const DecodeLaunch& L, cudaStream_t s, u64* launches)
{
    if (L.eType == E_HUF) {
        launch_huffman_decode(L, s, launches);
        return;
    }
    KLAUNCH(ans0_dec_scan_kernel, (L.nBlocks + 31) / 32, 32, s, L);
    const int groups = (L.maxChunks + 7) / 8;
    if (L.evK0)
        cudaEventRecord(L.evK0, s);
    KLAUNCH(ans0_decode_kernel, L.nBlocks * groups, 32, s, L);
    if (L.evK1)
        cudaEventRecord(L.evK1, s);
    *launches += 2;
}
