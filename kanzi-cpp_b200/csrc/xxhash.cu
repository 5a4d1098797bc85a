// xxhash.cu -- block checksums (kanzi's XXHash32 / XXHash64, util/XXHash.hpp:61-121,163-231) on sm_100a.
//
// EncodingTask::run hashes every block before the transforms and writes the value behind the block
// length (io/CompressedOutputStream.cpp:674-682, :804-807); DecodingTask::run hashes the decoded
// block and compares (io/CompressedInputStream.cpp:1003-1022).  Seed = the bitstream type "KANZ".
// A hash is four serial accumulator chains over 16-byte (32-bit variant) or 32-byte (64-bit variant)
// stripes: one quad per block, lane k owns accumulator k, eight blocks per warp; ~2 ms for a batch
// of 4 MiB blocks whatever its size (the chains of all blocks run side by side).
// Note the 64-bit variant's merge step rotates by (1, 7, 12, 18) with 32-bit complements
// (`(v << 1) | (v >> 31)` on a 64-bit value): kanzi's own arithmetic, reproduced as written.
#include "common.cuh"
#include "kernels.h"

#define XXH_SEED 0x4B414E5Au

#define P32_1 0x9E3779B1u
#define P32_2 0x85EBCA77u
#define P32_3 0xC2B2AE3Du
#define P32_4 0x27D4EB2Fu
#define P32_5 0x165667B1u
#define P64_1 0x9E3779B185EBCA87ull
#define P64_2 0xC2B2AE3D27D4EB4Full
#define P64_3 0x165667B19E3779F9ull
#define P64_4 0x85EBCA77C2B2AE63ull
#define P64_5 0x27D4EB2F165667C5ull

__host__ __device__ __forceinline__ u32 xxh_rotl32(u32 v, int r) { return (v << r) | (v >> (32 - r)); }
__host__ __device__ __forceinline__ u32 xxh32_round(u32 acc, u32 val) { return xxh_rotl32(acc + val * P32_2, 13) * P32_1; }
__host__ __device__ __forceinline__ u64 xxh64_round(u64 acc, u64 val)
{
    acc += val * P64_2;
    return ((acc << 31) | (acc >> 33)) * P64_1;
}
__host__ __device__ __forceinline__ u64 xxh64_merge(u64 acc, u64 val) { return (acc ^ xxh64_round(0, val)) * P64_1 + P64_4; }

__host__ __device__ __forceinline__ u32 xxh_le32(const u8* p)
{
    return (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24);
}
__host__ __device__ __forceinline__ u64 xxh_le64(const u8* p) { return (u64)xxh_le32(p) | ((u64)xxh_le32(p + 4) << 32); }

// tail + avalanche of the 32-bit hash (h already holds the merged accumulators or seed + P5)
__host__ __device__ __forceinline__ u32 xxh32_finish(u32 h, const u8* data, int idx, int length)
{
    h += (u32)length;
    while (idx <= length - 4) {
        h += xxh_le32(data + idx) * P32_3;
        h = xxh_rotl32(h, 17) * P32_4;
        idx += 4;
    }
    while (idx < length) {
        h += (u32)data[idx] * P32_5;
        h = xxh_rotl32(h, 11) * P32_1;
        idx++;
    }
    h ^= h >> 15;
    h *= P32_2;
    h ^= h >> 13;
    h *= P32_3;
    return h ^ (h >> 16);
}

__host__ __device__ __forceinline__ u64 xxh64_finish(u64 h, const u8* data, int idx, int length)
{
    h += (u64)(i64)length;
    while (idx + 8 <= length) {
        h ^= xxh64_round(0, xxh_le64(data + idx));
        h = ((h << 27) | (h >> 37)) * P64_1 + P64_4;
        idx += 8;
    }
    while (idx + 4 <= length) {
        h ^= (u64)xxh_le32(data + idx) * P64_1;
        h = ((h << 23) | (h >> 41)) * P64_2 + P64_3;
        idx += 4;
    }
    while (idx < length) {
        h ^= (u64)data[idx] * P64_5;
        h = ((h << 11) | (h >> 53)) * P64_1;
        idx++;
    }
    h ^= h >> 33;
    h *= P64_2;
    h ^= h >> 29;
    h *= P64_3;
    return h ^ (h >> 32);
}

// Host versions (small blocks are framed on the host; copy blocks are verified there).
u64 knz_xxhash_host(const u8* data, int length, int bits)
{
    if (bits == 32) {
        u32 h;
        int idx = 0;
        if (length >= 16) {
            u32 v1 = XXH_SEED + P32_1 + P32_2, v2 = XXH_SEED + P32_2, v3 = XXH_SEED, v4 = XXH_SEED - P32_1;
            do {
                v1 = xxh32_round(v1, xxh_le32(data + idx));
                v2 = xxh32_round(v2, xxh_le32(data + idx + 4));
                v3 = xxh32_round(v3, xxh_le32(data + idx + 8));
                v4 = xxh32_round(v4, xxh_le32(data + idx + 12));
                idx += 16;
            } while (idx <= length - 16);
            h = xxh_rotl32(v1, 1) + xxh_rotl32(v2, 7) + xxh_rotl32(v3, 12) + xxh_rotl32(v4, 18);
        } else {
            h = XXH_SEED + P32_5;
        }
        return (u64)xxh32_finish(h, data, idx, length);
    }
    u64 h;
    int idx = 0;
    const u64 seed = (u64)(i64)(int)XXH_SEED;
    if (length >= 32) {
        u64 v1 = seed + P64_1 + P64_2, v2 = seed + P64_2, v3 = seed, v4 = seed - P64_1;
        do {
            v1 = xxh64_round(v1, xxh_le64(data + idx));
            v2 = xxh64_round(v2, xxh_le64(data + idx + 8));
            v3 = xxh64_round(v3, xxh_le64(data + idx + 16));
            v4 = xxh64_round(v4, xxh_le64(data + idx + 24));
            idx += 32;
        } while (idx <= length - 32);
        h = ((v1 << 1) | (v1 >> 31)) + ((v2 << 7) | (v2 >> 25)) + ((v3 << 12) | (v3 >> 20)) + ((v4 << 18) | (v4 >> 14));
        h = xxh64_merge(h, v1);
        h = xxh64_merge(h, v2);
        h = xxh64_merge(h, v3);
        h = xxh64_merge(h, v4);
    } else {
        h = seed + P64_5;
    }
    return xxh64_finish(h, data, idx, length);
}

// One quad per block.  expect == NULL: store the hash; else compare and raise KERR_CRC in errFlag[2].
template <int BITS>
__global__ void __launch_bounds__(32)
xxh_kernel(BufTable bt, const BlkState* __restrict__ st, int nBlocks, u64* __restrict__ hash,
           const u64* __restrict__ expect, int* __restrict__ errFlag)
{
    const int lane = threadIdx.x, k = lane & 3, qbase = lane & ~3;
    const int g = blockIdx.x * 8 + (lane >> 2);
    const bool valid = g < nBlocks;
    BlkState bs;
    bs.len = 0, bs.cur = 0, bs.swaps = 0, bs.flags = 0;
    if (valid)
        bs = st[g];
    const int length = bs.len;
    const u8* __restrict__ data = valid ? blk_src(bt, bs, g) : NULL;
    const bool al = valid && ((((size_t)data) & 7) == 0);
    if (BITS == 32) {
        const int stripes = (length >= 16) ? (length >> 4) : 0;
        u32 v = (k == 0) ? XXH_SEED + P32_1 + P32_2 : (k == 1) ? XXH_SEED + P32_2 : (k == 2) ? XXH_SEED : XXH_SEED - P32_1;
        if (al) {
            const u32* __restrict__ w = reinterpret_cast<const u32*>(data) + k;
            int s = 0;
            for (; s + 8 <= stripes; s += 8) { // eight loads in flight per chain
                u32 x[8];
#pragma unroll
                for (int u = 0; u < 8; u++)
                    x[u] = __ldg(w + 4 * (s + u));
#pragma unroll
                for (int u = 0; u < 8; u++)
                    v = xxh32_round(v, x[u]);
            }
            for (; s < stripes; s++)
                v = xxh32_round(v, __ldg(w + 4 * s));
        } else {
            for (int s = 0; s < stripes; s++)
                v = xxh32_round(v, xxh_le32(data + 16 * s + 4 * k));
        }
        __syncwarp();
        const u32 v1 = __shfl_sync(FULL_MASK, v, qbase), v2 = __shfl_sync(FULL_MASK, v, qbase + 1),
                  v3 = __shfl_sync(FULL_MASK, v, qbase + 2), v4 = __shfl_sync(FULL_MASK, v, qbase + 3);
        if (valid && k == 0) {
            u32 h = stripes ? xxh_rotl32(v1, 1) + xxh_rotl32(v2, 7) + xxh_rotl32(v3, 12) + xxh_rotl32(v4, 18)
                            : XXH_SEED + P32_5;
            h = xxh32_finish(h, data, stripes << 4, length);
            if (expect == NULL)
                hash[g] = (u64)h;
            else if ((u64)h != expect[g])
                atomicExch(errFlag + 2, KERR_CRC);
        }
    } else {
        const int stripes = (length >= 32) ? (length >> 5) : 0;
        const u64 seed = (u64)(i64)(int)XXH_SEED;
        u64 v = (k == 0) ? seed + P64_1 + P64_2 : (k == 1) ? seed + P64_2 : (k == 2) ? seed : seed - P64_1;
        if (al) {
            const u64* __restrict__ w = reinterpret_cast<const u64*>(data) + k;
            int s = 0;
            for (; s + 8 <= stripes; s += 8) {
                u64 x[8];
#pragma unroll
                for (int u = 0; u < 8; u++)
                    x[u] = __ldg(w + 4 * (s + u));
#pragma unroll
                for (int u = 0; u < 8; u++)
                    v = xxh64_round(v, x[u]);
            }
            for (; s < stripes; s++)
                v = xxh64_round(v, __ldg(w + 4 * s));
        } else {
            for (int s = 0; s < stripes; s++)
                v = xxh64_round(v, xxh_le64(data + 32 * s + 8 * k));
        }
        __syncwarp();
        const u64 v1 = __shfl_sync(FULL_MASK, v, qbase), v2 = __shfl_sync(FULL_MASK, v, qbase + 1),
                  v3 = __shfl_sync(FULL_MASK, v, qbase + 2), v4 = __shfl_sync(FULL_MASK, v, qbase + 3);
        if (valid && k == 0) {
            u64 h;
            if (stripes) {
                h = ((v1 << 1) | (v1 >> 31)) + ((v2 << 7) | (v2 >> 25)) + ((v3 << 12) | (v3 >> 20)) +
                    ((v4 << 18) | (v4 >> 14));
                h = xxh64_merge(h, v1);
                h = xxh64_merge(h, v2);
                h = xxh64_merge(h, v3);
                h = xxh64_merge(h, v4);
            } else {
                h = seed + P64_5;
            }
            h = xxh64_finish(h, data, stripes << 5, length);
            if (expect == NULL)
                hash[g] = h;
            else if (h != expect[g])
                atomicExch(errFlag + 2, KERR_CRC);
        }
    }
}

void launch_xxhash(const BufTable& bt, const BlkState* st, int nBlocks, int bits, u64* hash, const u64* expect,
                   int* errFlag, cudaStream_t s, u64* launches)
{
    if (bits == 32)
        KLAUNCH((xxh_kernel<32>), (nBlocks + 7) / 8, 32, s, bt, st, nBlocks, hash, expect, errFlag);
    else
        KLAUNCH((xxh_kernel<64>), (nBlocks + 7) / 8, 32, s, bt, st, nBlocks, hash, expect, errFlag);
    *launches += 1;
}
