// common.cuh -- shared device/host helpers for libknzgpu (sm_100a).
#pragma once
#include <stdint.h>
#ifdef KNZ_SIM
// tests/sim/cusim.h is force-included (CPU emulation of the execution model, tests only)
#else
#include <cuda_runtime.h>
#define KLAUNCH(kernel, grid, block, stream, ...) kernel<<<grid, block, 0, stream>>>(__VA_ARGS__)
#define KLAUNCH_DYN(kernel, grid, block, smem, stream, ...) kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#define KNZ_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int64_t i64;

#define KNZ_WARP 32
#define FULL_MASK 0xFFFFFFFFu

// Per-block state entering transform stage s (array [stage][block]); written by
// the previous stage's decide step, read-only for the stage's heavy kernels.
struct BlkState {
    int len;    // bytes of current data
    int cur;    // which buffer holds it: 0 = A, 1 = B, 2 = caller input
    int swaps;  // successful stages so far (reference ping-pong parity)
    int flags;  // TransformSequence skip flags so far (bit 7-i set = stage i skipped)
};

// Buffer table for a batch: bufs[0]=A, bufs[1]=B (stride bstride), bufs[2]=input (stride inStride)
struct BufTable {
    u8* base[3];
    i64 stride[3];
};

__device__ __forceinline__ const u8* blk_src(const BufTable& t, const BlkState& s, int b)
{
    return t.base[s.cur] + (i64)b * t.stride[s.cur];
}

__device__ __forceinline__ u8* blk_dst(const BufTable& t, const BlkState& s, int b)
{
    const int d = (s.cur == 0) ? 1 : 0; // input (2) and B (1) both write to A
    return t.base[d] + (i64)b * t.stride[d];
}

__host__ __device__ __forceinline__ int next_cur(int cur) { return (cur == 0) ? 1 : 0; }

// Block header written by EncodingTask::run (io/CompressedOutputStream.cpp:757-807): mode byte,
// [skip-flag byte when the sequence has more than 4 transforms], the post-transform length on 1..4
// bytes, [the block checksum on 4 or 8 bytes].  hdrInfo = nTransforms | (checksumBytes << 8).
__host__ __device__ __forceinline__ int knz_len_bytes(int m)
{
    int r = 1;
    for (unsigned v = (unsigned)m; v >= 256; v >>= 8)
        r++;
    return r;
}
__host__ __device__ __forceinline__ int knz_hdr_bytes(int m, int hdrInfo)
{
    return 1 + (((hdrInfo & 0xFF) > 4) ? 1 : 0) + knz_len_bytes(m) + (hdrInfo >> 8);
}

__host__ __device__ __forceinline__ int ilog2_u32(u32 x) // floor(log2 x), x >= 1
{
#ifdef __CUDA_ARCH__
    return 31 - __clz((int)x);
#else
    return 31 - __builtin_clz(x);
#endif
}

__device__ __forceinline__ u32 bswap32(u32 x) { return __byte_perm(x, 0, 0x0123); }

__device__ __forceinline__ u32 lanemask_lt()
{
#ifdef KNZ_SIM
    return (1u << (threadIdx.x & 31)) - 1u;
#else
    u32 m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
#endif
}

__device__ __forceinline__ u32 warp_incl_sum(u32 v, int lane)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 t = __shfl_up_sync(FULL_MASK, v, o);
        if (lane >= o)
            v += t;
    }
    return v;
}

// Block-wide exclusive sum for blockDim.x == 256 (8 warps); smem >= 8 u32.
// Returns exclusive prefix of v; *total gets the block total (all threads).
__device__ __forceinline__ u32 block_excl_sum_256(u32 v, u32* smem8, u32* total)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const u32 inc = warp_incl_sum(v, lane);
    if (lane == 31)
        smem8[w] = inc;
    __syncthreads();
    u32 base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const u32 t = smem8[i];
        if (i < w)
            base += t;
        tot += t;
    }
    __syncthreads();
    *total = tot;
    return base + inc - v;
}

// ---------------------------------------------------------------- bit copy
// Destination streams are MSB-first bit strings stored in memory byte order; the
// destination base must be 4-byte aligned and zero-initialised.  Interior 32-bit
// words are stored whole, boundary words are OR-ed in atomically (pieces never
// overlap in bits, so OR is exact).
__device__ __forceinline__ u32 src_bits32(const u8* __restrict__ src, i64 nbits, i64 s)
{
    // bits [s, s+32) of the MSB-first source bit string; bits outside [0,nbits) read as 0
    const i64 nbytes = (nbits + 7) >> 3;
    const i64 b0 = s >> 3; // floor (s may be negative)
    u64 w = 0;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const i64 bi = b0 + k;
        const u32 v = (bi >= 0 && bi < nbytes) ? (u32)src[bi] : 0u;
        w = (w << 8) | v;
    }
    const int sh = (int)(s - (b0 << 3)); // 0..7
    u32 r = (u32)(w >> (8 - sh));
    // mask bits beyond the ends
    if (s < 0) {
        const int k = (int)(-s); // first k bits invalid
        r = (k >= 32) ? 0u : (r & (0xFFFFFFFFu >> k));
    }
    const i64 rem = nbits - s; // valid bits from s
    if (rem < 32)
        r = (rem <= 0) ? 0u : (r & ~(0xFFFFFFFFu >> (int)rem));
    return r;
}

__device__ __forceinline__ void bitcopy(u32* __restrict__ dstWords, u64 dstBit, const u8* __restrict__ src,
                                        i64 nbits, int tid, int nthreads)
{
    if (nbits <= 0)
        return;
    const u64 w0 = dstBit >> 5, w1 = (dstBit + (u64)nbits - 1) >> 5;
    for (u64 w = w0 + tid; w <= w1; w += nthreads) {
        const i64 s = (i64)(w << 5) - (i64)dstBit;
        const u32 v = src_bits32(src, nbits, s);
        const bool whole = (s >= 0) && (s + 32 <= nbits);
        if (whole)
            dstWords[w] = bswap32(v);
        else if (v != 0)
            atomicOr(&dstWords[w], bswap32(v));
    }
}

// OR `n` (<= 32) bits (right-aligned in v) into the stream at bit position pos.
__device__ __forceinline__ void put_bits_atomic(u32* dstWords, u64 pos, u32 v, int n)
{
    if (n <= 0)
        return;
    const u64 w = pos >> 5;
    const int off = (int)(pos & 31);
    const u64 x = ((u64)v << (64 - n)) >> off; // MSB-aligned 64-bit window starting at word w
    const u32 hi = (u32)(x >> 32), lo = (u32)x;
    if (hi)
        atomicOr(&dstWords[w], bswap32(hi));
    if (lo)
        atomicOr(&dstWords[w + 1], bswap32(lo));
}
