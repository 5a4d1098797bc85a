// fpaq.cu -- kanzi FPAQ (order-0 binary arithmetic coder with 4 x 256 adaptive probabilities).
//
// Reference: entropy/FPAQEncoder.cpp:58-103 (encode), FPAQEncoder.hpp:72-94 (encodeBit, flush),
// entropy/FPAQDecoder.cpp:61-120 (decode), FPAQDecoder.hpp:74-118 (decodeBit, read).
// The 56-bit interval [low, high] and the probabilities are carried from the first bit of a block
// to its last (they are not reset between the 4 MiB chunks), every bit's split depends on the
// interval left by the bit before it: a block is ONE dependency chain of 8 * n steps.  There is
// nothing to spread over a warp, so one lane per block runs the recurrence with the 1024
// probabilities in shared memory; parallelism is the number of blocks of the batch.  Output is byte
// aligned inside the block: per chunk  varint(bytes) | 32-bit words flushed by the interval | 56
// bits of low | 0xFFFFFF  (the last one written by dispose()).
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

#define FPAQ_CHUNK (4u << 20)
#define FPAQ_TOP 0x00FFFFFFFFFFFFFFull

// Move n bytes down by d (1..3) bytes, whole warp, source ahead of destination.
__device__ __forceinline__ void fpaq_shift_down(u8* dst, const u8* src, u32 n, int lane)
{
    for (u32 i = 0; i < n; i += 32) {
        u8 v = 0;
        if (i + lane < n)
            v = src[i + lane];
        __syncwarp();
        if (i + lane < n)
            dst[i + lane] = v;
        __syncwarp();
    }
}

// One warp per block, lane 0 codes.  stage[b] receives the block's entropy bytes, stageLen[b]
// their count.
__global__ void __launch_bounds__(32)
fpaq_encode_kernel(BufTable bt, const BlkState* __restrict__ st, u8* __restrict__ stage, i64 stageStride,
                   u32* __restrict__ stageLen, int* __restrict__ errFlag)
{
    __shared__ u16 s_p[4][256];
    __shared__ u32 s_idx;
    const int b = blockIdx.x, lane = threadIdx.x;
    const BlkState bs = st[b];
    const u32 count = (u32)bs.len;
    const u8* __restrict__ src = blk_src(bt, bs, b);
    u8* __restrict__ out = stage + (i64)b * stageStride;
    for (int i = lane; i < 1024; i += 32)
        (&s_p[0][0])[i] = 32768; // PSCALE >> 1
    __syncwarp();
    u64 low = 0, high = FPAQ_TOP;
    u32 pos = 0; // bytes of the block's output so far
    u32 start = 0;
    while (start < count) {
        const u32 chunk = min(FPAQ_CHUNK, count - start);
        u8* __restrict__ buf = out + pos + 4; // room for the largest varint in front
        if (lane == 0) {
            u32 idx = 0;
            const u32 limit = (u32)min((i64)0x7FFFFFFF, stageStride - (i64)pos - 4 - 16);
            u16* p = s_p[0];
            for (u32 i = start; i < start + chunk; i++) {
                const int val = src[i];
                const int bits = val + 256;
#pragma unroll
                for (int k = 7; k >= 0; k--) {
                    u16* pr = &p[(k == 7) ? 1 : (bits >> (k + 1))];
                    const u32 pv = *pr;
                    const u64 split = (((high - low) >> 8) * (u64)pv) >> 8;
                    if (((val >> k) & 1) == 0) {
                        low = low + split + 1;
                        *pr = (u16)(pv - (pv >> 6));
                    } else {
                        high = low + split;
                        *pr = (u16)(pv - (u32)(((int)pv - 65536 + 64) >> 6));
                    }
                    if (((low ^ high) >> 24) == 0) { // flush the 32 settled bits
                        if (idx + 4 <= limit) {
                            const u32 v = (u32)(high >> 24);
                            buf[idx] = (u8)(v >> 24);
                            buf[idx + 1] = (u8)(v >> 16);
                            buf[idx + 2] = (u8)(v >> 8);
                            buf[idx + 3] = (u8)v;
                        }
                        idx += 4;
                        low <<= 32;
                        high = (high << 32) | 0xFFFFFFFFull;
                    }
                }
                p = s_p[val >> 6];
            }
            if (idx + 4 > limit) {
                atomicExch(errFlag, KERR_OUT_OVERFLOW);
                idx = 0;
            }
            s_idx = idx;
        }
        __syncwarp();
        const u32 idx = s_idx;
        // varint(idx) then the bytes, contiguous
        int vl = 1;
        for (u32 v = idx; v >= 128; v >>= 7)
            vl++;
        if (vl < 4)
            fpaq_shift_down(out + pos + vl, buf, idx, lane);
        if (lane == 0) {
            u32 v = idx;
            u8* q = out + pos;
            while (v >= 128) {
                *q++ = (u8)(0x80 | (v & 0x7F));
                v >>= 7;
            }
            *q = (u8)v;
        }
        pos += (u32)vl + idx;
        start += chunk;
        if (lane == 0) { // between chunks and at dispose(): 56 bits of low | MASK_0_24
            const u64 w = (low | 0xFFFFFFull) & FPAQ_TOP;
            for (int k = 0; k < 7; k++)
                out[pos + k] = (u8)(w >> (48 - 8 * k));
        }
        pos += 7;
        __syncwarp();
    }
    if (lane == 0)
        stageLen[b] = pos;
}

// blockBits = header bytes + staged bytes
__global__ void fpaq_bits_kernel(const BlkState* __restrict__ st, int nBlocks, int nTransforms,
                                 const u32* __restrict__ stageLen, u64* __restrict__ blockBits, i64 outStride,
                                 int* __restrict__ errFlag)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nBlocks)
        return;
    const int m = st[b].len;
    const int hdrBytes = knz_hdr_bytes(m, nTransforms);
    const u64 total = 8ull * ((u64)hdrBytes + stageLen[b]);
    blockBits[b] = total;
    if ((i64)(total >> 3) + 8 > outStride)
        atomicExch(errFlag, KERR_OUT_OVERFLOW);
}

__global__ void __launch_bounds__(256)
fpaq_copy_kernel(const BlkState* __restrict__ st, int nTransforms, const u8* __restrict__ stage, i64 stageStride,
                 const u32* __restrict__ stageLen, u8* __restrict__ out, i64 outStride)
{
    const int b = blockIdx.y;
    const int m = st[b].len;
    const int hdrBytes = knz_hdr_bytes(m, nTransforms);
    const u32 n = stageLen[b];
    if ((i64)hdrBytes + n + 8 > outStride)
        return;
    const u8* __restrict__ s = stage + (i64)b * stageStride;
    u8* __restrict__ d = out + (i64)b * outStride + hdrBytes;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        d[i] = s[i];
}

void launch_fpaq_encode(const EncodeLaunch& L, cudaStream_t s, u64* launches)
{
    const int nB = L.nBlocks;
    const i64 stageStride = (i64)L.maxChunks * ANS_SLOT; // the rANS chunk slots double as the staging area
    u32* stageLen = L.payBytes;
    if (L.evK0)
        cudaEventRecord(L.evK0, s);
    KLAUNCH(fpaq_encode_kernel, nB, 32, s, L.bt, L.st, L.slots, stageStride, stageLen, L.errFlag);
    if (L.evK1)
        cudaEventRecord(L.evK1, s);
    KLAUNCH(fpaq_bits_kernel, (nB + 127) / 128, 128, s, L.st, nB, L.nTransforms, stageLen, L.blockBits, L.outStride,
            L.errFlag);
    launch_out_prepare_and_header(L, s, launches);
    KLAUNCH(fpaq_copy_kernel, dim3(64, nB), 256, s, L.st, L.nTransforms, L.slots, stageStride, stageLen, L.out,
            L.outStride);
    *launches += 3;
}

// ------------------------------------------------------------------ decoder
__device__ __forceinline__ u32 fpaq_rd(const u8* __restrict__ p, u64 pos, int n)
{
    const u64 b0 = pos >> 3;
    u64 w = 0;
#pragma unroll
    for (int k = 0; k < 5; k++)
        w = (w << 8) | p[b0 + k];
    const int sh = (int)(pos & 7);
    return (u32)((w >> (40 - sh - n)) & ((n == 32) ? 0xFFFFFFFFull : ((1ull << n) - 1)));
}

__global__ void __launch_bounds__(32)
fpaq_decode_kernel(DecodeLaunch L)
{
    __shared__ u16 s_p[4][256];
    const int b = blockIdx.x, lane = threadIdx.x;
    const u32 count = (u32)L.preLen[b];
    const u64 endBits = L.inBits[b];
    const u8* __restrict__ in = L.in + (i64)b * L.inStride;
    u8* __restrict__ out = L.dst + (i64)b * L.dstStride;
    for (int i = lane; i < 1024; i += 32)
        (&s_p[0][0])[i] = 32768;
    __syncwarp();
    if (lane != 0)
        return;
    u64 pos = L.payStart[b];
    u64 low = 0, high = FPAQ_TOP, current = 0;
    bool bad = false;
    u32 start = 0;
    while (start < count && !bad) {
        // szBytes = readVarInt (EntropyUtils.cpp:261-286)
        if (pos + 8 > endBits) {
            bad = true;
            break;
        }
        u32 v = fpaq_rd(in, pos, 8);
        pos += 8;
        u32 sz = v & 0x7F;
        for (int shift = 7; v >= 128 && shift <= 28; shift += 7) {
            if (pos + 8 > endBits) {
                bad = true;
                break;
            }
            v = fpaq_rd(in, pos, 8);
            pos += 8;
            sz |= (v & 0x7F) << shift;
        }
        if (bad || sz >= 2 * count || pos + 56 + 8ull * sz > endBits) { // FPAQDecoder.cpp:76-77 + bounds
            bad = true;
            break;
        }
        current = ((u64)fpaq_rd(in, pos, 24) << 32) | (u64)fpaq_rd(in, pos + 24, 32);
        pos += 56;
        const u64 bufPos = pos; // sz bytes follow
        u32 idx = 0;
        const u32 chunk = min(FPAQ_CHUNK, count - start);
        u16* p = s_p[0];
        for (u32 i = start; i < start + chunk; i++) {
            int ctx = 1;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const u32 pv = p[ctx];
                const u64 split = ((((high - low) >> 8) * (u64)pv) >> 8) + low;
                if (split >= current) {
                    high = split;
                    p[ctx] = (u16)(pv - (u32)(((int)pv - 65536 + 64) >> 6));
                    ctx += ctx + 1;
                } else {
                    low = split + 1;
                    p[ctx] = (u16)(pv - (pv >> 6));
                    ctx += ctx;
                }
                if (((low ^ high) >> 24) == 0) { // read(): 32 more bits
                    low = (low << 32) & FPAQ_TOP;
                    high = ((high << 32) | 0xFFFFFFFFull) & FPAQ_TOP;
                    if (idx + 4 > sz) {
                        current = (current << 32) & FPAQ_TOP;
                        idx = sz + 1;
                    } else {
                        const u64 val = fpaq_rd(in, bufPos + 8ull * idx, 32);
                        current = ((current << 32) | val) & FPAQ_TOP;
                        idx += 4;
                    }
                }
            }
            out[i] = (u8)ctx;
            if (idx > sz) {
                bad = true;
                break;
            }
            p = s_p[(ctx & 0xFF) >> 6];
        }
        pos = bufPos + 8ull * sz;
        start += chunk;
    }
    if (bad)
        atomicExch(L.errFlag + 1, KERR_BAD_STREAM);
}

void launch_fpaq_decode(const DecodeLaunch& L, cudaStream_t s, u64* launches)
{
    if (L.evK0)
        cudaEventRecord(L.evK0, s);
    KLAUNCH(fpaq_decode_kernel, L.nBlocks, 32, s, L);
    if (L.evK1)
        cudaEventRecord(L.evK1, s);
    *launches += 1;
}
