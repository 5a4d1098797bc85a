// blockscan.cu -- the `skipBlocks` test of EncodingTask<T>::run (io/CompressedOutputStream.cpp:
// 697-715) on the device: a block is stored as a COPY block (mode 0x80, NullTransform, raw bytes)
// when its first four bytes are the signature of an already compressed format (Magic.hpp) or when
// its order-0 entropy, in the reference's fixed-point arithmetic (Global::computeFirstOrderEntropy1024,
// Global.cpp:313-329, log2 from Global::log2_1024 :136-150), reaches 973 / 1024 bits per byte.
//
//   skip_decide_kernel   one CTA per block: 128-bit loads, per-warp shared-memory histograms,
//                        entropy sum by a block reduction, decision into skip[b]
//   copy_frame_kernel    rewrites the private buffer of every flagged block as the reference's copy
//                        block: mode | length | [checksum] | the block's bytes; exact bit count
//
// The transform and entropy stages of a batch run for every block regardless (their launches are batch
// wide); the framing kernel runs last and overrides the flagged ones.  Incompressible blocks are the
// cheap case of every stage here (no repeats for the suffix sort, ZRLT refuses), so what the reference
// saves on the CPU by skipping is small on this path; the stream is what has to match.
#include "kernels.h"

// 1024 * log2(x) from the 257-entry table of round(4096 * log2(i)) (uploaded by the host)
__device__ __forceinline__ int log2_1024_dev(u32 x, const int* __restrict__ tab)
{
    if (x < 256)
        return (tab[x] + 2) >> 2;
    const int lg = ilog2_u32(x);
    if ((x & (x - 1)) == 0)
        return lg << 10;
    return ((lg - 7) << 10) + ((tab[x >> (lg - 7)] + 2) >> 2);
}

// Signatures of formats that are compressed already (Magic.hpp:27-56, isCompressed :108-131):
// JPG (low nibble free), GIF, PNG, 7z/LZMA, ZSTD, BROTLI, CAB, ZIP, FLAC, XZ, KNZ, RAR on 32 bits,
// BZIP2 and MP3/ID3 on 24 bits, GZIP on 16 bits.
__device__ __forceinline__ bool magic_compressed(u32 key)
{
    if ((key & ~0x0Fu) == 0xFFD8FFE0u)
        return true;
    const u32 k24 = key >> 8;
    if (k24 == 0x425A68u || k24 == 0x494433u)
        return true;
    switch (key) {
    case 0x47494638u: case 0x504B0304u: case 0x377ABCAFu: case 0x89504E47u: case 0x28B52FFDu:
    case 0x81CFB2CEu: case 0x4D534346u: case 0x664C6143u: case 0xFD377A58u: case 0x4B414E5Au:
    case 0x52617221u:
        return true;
    default:
        break;
    }
    // 32-bit signatures that are NOT compressed formats still win over the 16-bit ones (getType :69-86)
    switch (key) {
    case 0x25504446u: case 0x7F454C46u: case 0xFEEDFACEu: case 0xCEFAEDFEu: case 0xFEEDFACFu:
    case 0xCFFAEDFEu: case 0x52494646u:
        return false;
    default:
        break;
    }
    return (key >> 16) == 0x1F8Bu;
}

__global__ void __launch_bounds__(256)
skip_decide_kernel(BufTable bt, const BlkState* __restrict__ st, int nBlocks, const int* __restrict__ log2tab,
                   int* __restrict__ skip)
{
    __shared__ u32 hist[8][256];
    __shared__ u64 part[256];
    const int b = blockIdx.x;
    if (b >= nBlocks)
        return;
    const BlkState bs = st[b];
    const int n = bs.len;
    const u8* __restrict__ src = blk_src(bt, bs, b);
    const int tid = threadIdx.x, w = tid >> 5;
    for (int i = tid; i < 8 * 256; i += 256)
        (&hist[0][0])[i] = 0;
    __syncthreads();
    const bool al = (((uintptr_t)src) & 15) == 0;
    const int body = al ? (n & ~15) : 0;
    const uint4* __restrict__ src16 = reinterpret_cast<const uint4*>(src);
    for (int i = tid; i < (body >> 4); i += 256) {
        const uint4 v = __ldg(src16 + i);
        const u32 x[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
        for (int k = 0; k < 4; k++) {
            atomicAdd(&hist[w][x[k] & 0xFF], 1u);
            atomicAdd(&hist[w][(x[k] >> 8) & 0xFF], 1u);
            atomicAdd(&hist[w][(x[k] >> 16) & 0xFF], 1u);
            atomicAdd(&hist[w][x[k] >> 24], 1u);
        }
    }
    for (int i = body + tid; i < n; i += 256)
        atomicAdd(&hist[w][src[i]], 1u);
    __syncthreads();
    u32 c = 0;
#pragma unroll
    for (int k = 0; k < 8; k++)
        c += hist[k][tid];
    u64 term = 0;
    if (c != 0 && n > 0) {
        const int logN = log2_1024_dev((u32)n, log2tab);
        term = ((u64)c * (u64)(i64)(logN - log2_1024_dev(c, log2tab))) >> 3;
    }
    part[tid] = term;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o)
            part[tid] += part[tid + o];
        __syncthreads();
    }
    if (tid == 0) {
        bool sk = false;
        if (n >= 4) {
            const u32 key = ((u32)src[0] << 24) | ((u32)src[1] << 16) | ((u32)src[2] << 8) | (u32)src[3];
            sk = magic_compressed(key);
        }
        if (!sk && n > 0)
            sk = (int)(part[0] / (u64)n) >= 973; // EntropyUtils::INCOMPRESSIBLE_THRESHOLD
        skip[b] = sk ? 1 : 0;
    }
}

// Copy block of a flagged block: mode byte 0x80 | (length bytes - 1) << 5 | 0x07 (one NullTransform,
// applied: skip flags 0x7F >> 4), the length, the checksum when enabled, the bytes themselves
// (io/CompressedOutputStream.cpp:757-807 with tType = eType = NONE).
__global__ void __launch_bounds__(256)
copy_frame_kernel(BufTable bt, const BlkState* __restrict__ st0, int nBlocks, const int* __restrict__ skip,
                  int ckBytes, const u64* __restrict__ blockHash, u8* __restrict__ out, i64 outStride,
                  u64* __restrict__ blockBits)
{
    const int b = blockIdx.y;
    if (b >= nBlocks || !skip[b])
        return;
    const BlkState bs = st0[b];
    const int n = bs.len;
    const u8* __restrict__ src = blk_src(bt, bs, b);
    const int dataSize = knz_len_bytes(n);
    const int hdr = 1 + dataSize + ckBytes;
    const u64 hash = ckBytes ? blockHash[b] : 0;
    const i64 total = (i64)hdr + n;
    const i64 words = (total + 3) >> 2;
    u32* __restrict__ dst = reinterpret_cast<u32*>(out + (i64)b * outStride);
    for (i64 wd = (i64)blockIdx.x * blockDim.x + threadIdx.x; wd < words; wd += (i64)gridDim.x * blockDim.x) {
        u32 v = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const i64 p = 4 * wd + j;
            u32 x = 0;
            if (p == 0)
                x = 0x80u | (u32)(((dataSize - 1) & 3) << 5) | 0x07u;
            else if (p <= dataSize)
                x = ((u32)n >> (8 * (dataSize - (int)p))) & 0xFF;
            else if (p < hdr)
                x = (u32)(hash >> (8 * (hdr - 1 - (int)p))) & 0xFF;
            else if (p < total)
                x = src[p - hdr];
            v |= x << (8 * j);
        }
        dst[wd] = v;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        blockBits[b] = 8ull * (u64)total;
}

void launch_skip_decide(const BufTable& bt, const BlkState* st, int nBlocks, const int* log2tab, int* skip,
                        cudaStream_t s, u64* launches)
{
    KLAUNCH(skip_decide_kernel, nBlocks, 256, s, bt, st, nBlocks, log2tab, skip);
    *launches += 1;
}

void launch_copy_frame(const BufTable& bt, const BlkState* st0, int nBlocks, const int* skip, int ckBytes,
                       const u64* blockHash, u8* out, i64 outStride, u64* blockBits, cudaStream_t s, u64* launches)
{
    KLAUNCH(copy_frame_kernel, dim3(64, nBlocks), 256, s, bt, st0, nBlocks, skip, ckBytes, blockHash, out, outStride,
            blockBits);
    *launches += 1;
}
