// bwt.cu -- Burrows-Wheeler transform (kanzi BWT block codec) on sm_100a.
//
// Replaces BWTBlockCodec::forward/inverse (transform/BWTBlockCodec.cpp:32-168),
// BWT::forward (transform/BWT.cpp:92-134) + DivSufSort::computeBWT
// (transform/DivSufSort.cpp:171-295) and BWT::inverse (transform/BWT.cpp:136-657).
// The BWT is canonical, so the induced-sorting suffix sorter of the reference is
// replaced by batched prefix doubling built on a segmented LSD radix sort:
//   * initial order: 8-byte big-endian prefixes (zero padded), radix sorted
//   * each round h: only suffixes whose group is still ambiguous stay in play;
//     they are re-keyed (group rank, rank[i+h]) and radix sorted; equal-key runs
//     become the new groups; singleton groups leave the working set
//   * only the inverse suffix array is materialised; the output is scattered as
//     out[rank(p) + (rank(p) < rank(0))] = in[p-1], out[0] = in[n-1], primary
//     index k = rank(k*ceil(n/8)) + 1 (same contract as constructBWT).
// Many blocks are sorted in one launch sequence (segment = block).
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

#define RS_THREADS 256
#define RS_ITEMS 8
#define RS_TILE (RS_THREADS * RS_ITEMS)

__device__ __forceinline__ int bwt_chunks(int n) { return (n < 256) ? 1 : 8; }

__device__ __forceinline__ int bwt_pisz(int n)
{
    int lg = ilog2_u32((u32)n);
    if (n & (n - 1))
        lg++;
    return (lg + 7) >> 3;
}

// ------------------------------------------------------------------ radix sort
struct SortArrays {
    u64* key[2];
    u32* val[2];
    u32* hist;      // [nBlocks][maxTiles][256]
    int* ticket;    // [8][maxBlocks]: next tile index of block b in pass p (one-sweep)
    u32* totals;    // [nBlocks][8][256]
    int* which;     // [9][maxBlocks]: buffer index holding block b's data before pass p
    int* trivial;   // [8][maxBlocks]
    const int* cnt; // [nBlocks]
    int capN, maxTiles, maxBlocks;
};

// Digit totals for all 8 byte positions in one read of the keys.
__global__ void __launch_bounds__(RS_THREADS)
rs_totals_kernel(SortArrays A)
{
    __shared__ u32 s_h[8][256];
    const int b = blockIdx.y;
    const int cnt = A.cnt[b];
    const int base = blockIdx.x * RS_TILE;
    if (base >= cnt)
        return;
    for (int i = threadIdx.x; i < 2048; i += RS_THREADS)
        (&s_h[0][0])[i] = 0;
    __syncthreads();
    const u64* __restrict__ k = A.key[A.which[b]] + (i64)b * A.capN;
#pragma unroll
    for (int it = 0; it < RS_ITEMS; it++) {
        const int j = base + it * RS_THREADS + threadIdx.x;
        if (j < cnt) {
            const u64 v = k[j];
#pragma unroll
            for (int p = 0; p < 8; p++)
                atomicAdd(&s_h[p][(v >> (8 * p)) & 0xFF], 1u);
        }
    }
    __syncthreads();
    u32* tot = A.totals + (i64)b * 2048;
    for (int i = threadIdx.x; i < 2048; i += RS_THREADS) {
        const u32 v = (&s_h[0][0])[i];
        if (v)
            atomicAdd(&tot[i], v);
    }
}

// Per block: which passes are trivial (all keys share the digit) and where the
// data lives before every pass.
__global__ void rs_plan_kernel(SortArrays A, int nBlocks)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nBlocks)
        return;
    const u32 cnt = (u32)A.cnt[b];
    int w = A.which[b];
    const u32* tot = A.totals + (i64)b * 2048;
    for (int p = 0; p < 8; p++) {
        bool triv = (cnt <= 1);
        if (!triv)
            for (int d = 0; d < 256; d++)
                if (tot[p * 256 + d] == cnt) {
                    triv = true;
                    break;
                }
        A.trivial[p * A.maxBlocks + b] = triv ? 1 : 0;
        if (!triv)
            w ^= 1;
        A.which[(p + 1) * A.maxBlocks + b] = w;
    }
}

// One-sweep pass (Adinets & Merrill): histogram, tile offsets and scatter in ONE kernel.
// The tile's digit counts are chained to its predecessors with a decoupled look-back over
// status words (2 flag bits + 30-bit count) instead of a separate histogram pass and a
// scan kernel, so a pass reads 12 and writes 12 bytes per element.  The sorted tile is
// staged in shared memory and written out in sorted order: consecutive threads store
// consecutive addresses inside a digit run.  Tiles take their index from a per-block
// ticket, so every predecessor of a running tile is itself running or finished.
#define OS_AGG 0x40000000u
#define OS_INC 0x80000000u
#define OS_VAL 0x3FFFFFFFu
#define OS_SMEM(items) (RS_THREADS * (items) * 12 + (RS_THREADS / 32) * 256 * 4 + 2 * 256 * 4 + 64)

#define OS_LOOK 8
// status words travel on their own (the count is in the word): relaxed device-scope accesses
__device__ __forceinline__ void os_publish(u32* p, u32 v)
{
#ifndef KNZ_SIM
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#else
    *(volatile u32*)p = v;
#endif
}
__device__ __forceinline__ u32 os_peek(const u32* p)
{
#ifndef KNZ_SIM
    u32 v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
#else
    return *(const volatile u32*)p;
#endif
}

// CTA -> block mapping of the one-sweep kernels (1-D grid of tilesMax * nBlocks CTAs): consecutive
// CTAs serve different blocks of a group of OS_GROUP blocks, so the CTAs in flight are spread over
// OS_GROUP look-back chains (a chain link costs an L2 round trip; with one block at a time every
// tile polled a long run of unfinished predecessors) while the group's text stays L2 resident.
#define OS_GROUP 16
__device__ __forceinline__ int os_block_of_cta(int nBlocks, int tilesMax)
{
    const int id = (int)blockIdx.x;
    const int per = OS_GROUP * tilesMax;
    const int g = id / per;
    const int r = id - g * per;
    const int nbg = min(OS_GROUP, nBlocks - g * OS_GROUP);
    return g * OS_GROUP + (r % nbg);
}

template <int OS_ITEMS>
__global__ void __launch_bounds__(RS_THREADS, (OS_ITEMS <= 8) ? 4 : 2)
rs_onesweep_kernel(SortArrays A, int pass, int nBlocks, int tilesMax)
{
    constexpr int OS_TILE = RS_THREADS * OS_ITEMS;
    KNZ_DYN_SMEM(os_smem);
    u64* s_key = reinterpret_cast<u64*>(os_smem);
    u32* s_val = reinterpret_cast<u32*>(os_smem + OS_TILE * 8);
    u32(*s_cnt)[256] = reinterpret_cast<u32(*)[256]>(os_smem + OS_TILE * 12);
    u32* s_delta = reinterpret_cast<u32*>(os_smem + OS_TILE * 12 + (RS_THREADS / 32) * 1024);
    u32* s_toff = s_delta + 256;
    u32* s_w = s_toff + 256; // 8 words for the block scan + 1 for the ticket
    const int b = os_block_of_cta(nBlocks, tilesMax);
    const int cnt = A.cnt[b];
    if (cnt <= 0 || A.trivial[pass * A.maxBlocks + b])
        return;
    const int tiles = (cnt + OS_TILE - 1) / OS_TILE;
    if (threadIdx.x == 0)
        s_w[8] = (u32)atomicAdd(&A.ticket[pass * A.maxBlocks + b], 1);
    for (int i = threadIdx.x; i < (RS_THREADS / 32) * 256; i += RS_THREADS)
        (&s_cnt[0][0])[i] = 0;
    __syncthreads();
    const int tile = (int)s_w[8];
    if (tile >= tiles)
        return; // every block has tilesMax CTAs: the surplus ones of a short block leave here
    const int tbase = tile * OS_TILE;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int src = A.which[pass * A.maxBlocks + b];
    const u64* __restrict__ kin = (src ? A.key[1] : A.key[0]) + (i64)b * A.capN;
    const u32* __restrict__ vin = (src ? A.val[1] : A.val[0]) + (i64)b * A.capN;
    u64* __restrict__ kout = (src ? A.key[0] : A.key[1]) + (i64)b * A.capN;
    u32* __restrict__ vout = (src ? A.val[0] : A.val[1]) + (i64)b * A.capN;
    const int sh = 8 * pass;
    u64 key[OS_ITEMS];
    u32 val[OS_ITEMS];
    u16 rnk[OS_ITEMS];
#pragma unroll
    for (int it = 0; it < OS_ITEMS; it++) {
        const int j = tbase + w * (32 * OS_ITEMS) + it * 32 + lane;
        key[it] = (j < cnt) ? __ldg(&kin[j]) : 0;
    }
#pragma unroll
    for (int it = 0; it < OS_ITEMS; it++) {
        const int j = tbase + w * (32 * OS_ITEMS) + it * 32 + lane;
        val[it] = (j < cnt) ? __ldg(&vin[j]) : 0;
    }
    // rank inside the warp's 512 elements (load order = stable order)
#pragma unroll
    for (int it = 0; it < OS_ITEMS; it++) {
        const int j = tbase + w * (32 * OS_ITEMS) + it * 32 + lane;
        const u32 d = (j < cnt) ? (u32)((key[it] >> sh) & 0xFF) : 256u; // 256 = padding class
        const u32 peers = __match_any_sync(FULL_MASK, d);
        const u32 prior = (d < 256) ? s_cnt[w][d] : 0;
        __syncwarp();
        if (d < 256 && (peers & lanemask_lt()) == 0)
            s_cnt[w][d] = prior + __popc(peers);
        __syncwarp();
        rnk[it] = (u16)(prior + __popc(peers & lanemask_lt()));
    }
    __syncthreads();
    // thread = digit: warp prefixes, tile count, look-back, offsets
    const int d = threadIdx.x;
    u32 tcount = 0;
#pragma unroll
    for (int x = 0; x < RS_THREADS / 32; x++) {
        const u32 v = s_cnt[x][d];
        s_cnt[x][d] = tcount;
        tcount += v;
    }
    u32* st = A.hist + ((i64)b * A.maxTiles) * 256 + d;
    os_publish(st + (i64)tile * 256, (tile == 0) ? (OS_INC | tcount) : (OS_AGG | tcount));
    u32 before = 0; // elements with this digit in earlier tiles
    // look-back, OS_LOOK predecessors per round trip (the loads are independent)
    for (int t = tile - 1; t >= 0;) {
        u32 v[OS_LOOK];
#pragma unroll
        for (int k = 0; k < OS_LOOK; k++)
            v[k] = (t - k >= 0) ? os_peek(st + (i64)(t - k) * 256) : OS_INC; // nothing precedes tile 0
        bool done = false;
#pragma unroll
        for (int k = 0; k < OS_LOOK; k++) {
            if (done || !(v[k] & (OS_INC | OS_AGG)))
                break; // not published yet: poll again from this tile
            before += v[k] & OS_VAL;
            t--;
            done = (v[k] & OS_INC) != 0;
        }
        if (done)
            break;
    }
    if (tile > 0)
        os_publish(st + (i64)tile * 256, OS_INC | (before + tcount));
    u32 tot;
    const u32 dbase = block_excl_sum_256(A.totals[(i64)b * 2048 + pass * 256 + d], s_w, &tot); // digits < d, whole block
    const u32 toff = block_excl_sum_256(tcount, s_w, &tot);                                    // digits < d, this tile
    s_delta[d] = dbase + before - toff;
    const u32 mine = toff; // first staged slot of digit d
    // stage in sorted order
    s_toff[d] = mine;
    __syncthreads();
#pragma unroll
    for (int it = 0; it < OS_ITEMS; it++) {
        const int j = tbase + w * (32 * OS_ITEMS) + it * 32 + lane;
        if (j < cnt) {
            const u32 dg = (u32)((key[it] >> sh) & 0xFF);
            const u32 q = s_toff[dg] + s_cnt[w][dg] + rnk[it];
            s_key[q] = key[it];
            s_val[q] = val[it];
        }
    }
    __syncthreads();
    const int valid = min(OS_TILE, cnt - tbase);
#pragma unroll
    for (int it = 0; it < OS_ITEMS; it++) {
        const int q = it * RS_THREADS + threadIdx.x;
        if (q < valid) {
            const u64 k = s_key[q];
            const u32 pos = s_delta[(u32)((k >> sh) & 0xFF)] + (u32)q;
            kout[pos] = k;
            vout[pos] = s_val[q];
        }
    }
}

static void radix_sort(const SortArrays& A, int nBlocks, int maxCnt, u32 passMask, cudaStream_t s, u64* launches)
{
    const int tiles = (maxCnt + RS_TILE - 1) / RS_TILE;
    if (tiles <= 0)
        return;
    static int items = 0;
    if (!items) {
        const char* e = getenv("KNZ_OS_ITEMS"); // experiments: 8 (4 CTAs/SM) or 16 (2 CTAs/SM)
        items = (e && atoi(e) == 16) ? 16 : 8;
        cudaFuncSetAttribute(rs_onesweep_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, OS_SMEM(8));
        cudaFuncSetAttribute(rs_onesweep_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, OS_SMEM(16));
    }
    cudaMemsetAsync(A.totals, 0, sizeof(u32) * 2048 * (size_t)nBlocks, s);
    cudaMemsetAsync(A.ticket, 0, sizeof(int) * 8 * (size_t)A.maxBlocks, s);
    KLAUNCH(rs_totals_kernel, dim3(tiles, nBlocks), RS_THREADS, s, A);
    KLAUNCH(rs_plan_kernel, (nBlocks + 63) / 64, 64, s, A, nBlocks);
    *launches += 2;
    const int otile = RS_THREADS * items;
    const int otiles = (maxCnt + otile - 1) / otile;
    for (int p = 0; p < 8; p++) {
        if (!((passMask >> p) & 1))
            continue; // digit statically zero for every key: the plan marks it trivial as well
        // status words of the tiles this pass can touch: [block][tile][256]
        cudaMemset2DAsync(A.hist, sizeof(u32) * 256 * (size_t)A.maxTiles, 0, sizeof(u32) * 256 * (size_t)otiles,
                          (size_t)nBlocks, s);
        if (items == 16)
            KLAUNCH_DYN(rs_onesweep_kernel<16>, otiles * nBlocks, RS_THREADS, OS_SMEM(16), s, A, p, nBlocks, otiles);
        else
            KLAUNCH_DYN(rs_onesweep_kernel<8>, otiles * nBlocks, RS_THREADS, OS_SMEM(8), s, A, p, nBlocks, otiles);
        *launches += 1;
    }
}


// ------------------------------------------------------------------ initial sort, text mode
// For blocks of at most 4 MiB the first round does not move keys at all: the element is the
// suffix index (22 bits) and the digit of pass p is the text byte at index + 7 - p, which is
// either read in order (pass 0), carried in the top bits of the element from the pass before
// (odd passes) or gathered from the block's text, L2 resident while its tiles are in flight
// (passes 2, 4, 6; the gather also fetches the byte the next pass needs).  A pass then moves
// 4 + 4 bytes per element instead of 12 + 12.
#define TX_IDX_BITS 22
#define TX_IDX_MASK ((1u << TX_IDX_BITS) - 1u)
#define TX_SMEM(items) (RS_THREADS * (items) * 4 + (RS_THREADS / 32) * 256 * 4 + 2 * 256 * 4 + 64 + RS_THREADS * (items))

struct TextSort {
    BufTable bt;
    const BlkState* st;
};

// H[d] = occurrences of byte d in the block -> totals row 0 (the plan derives all 8 rows)
__global__ void __launch_bounds__(RS_THREADS)
rs_text_hist_kernel(SortArrays A, TextSort T)
{
    __shared__ u32 s_h[256];
    const int b = blockIdx.y;
    const int cnt = A.cnt[b];
    const int base = blockIdx.x * (RS_TILE * 4);
    if (base >= cnt)
        return;
    s_h[threadIdx.x] = 0;
    __syncthreads();
    const u8* __restrict__ text = blk_src(T.bt, T.st[b], b);
    const int end = min(cnt, base + RS_TILE * 4);
    const int lane = threadIdx.x & 31;
    for (int j0 = base + (threadIdx.x & ~31); j0 < end; j0 += RS_THREADS) { // one atomic per distinct byte per warp row
        const int j = j0 + lane;
        const u32 c = (j < end) ? (u32)text[j] : (256u + (u32)lane);
        const u32 peers = __match_any_sync(FULL_MASK, c);
        if (j < end && (peers >> lane) == 1u)
            atomicAdd(&s_h[c], (u32)__popc(peers));
    }
    __syncthreads();
    const u32 v = s_h[threadIdx.x];
    if (v)
        atomicAdd(&A.totals[(i64)b * 2048 + threadIdx.x], v);
}

// One CTA per block: totals of pass p = histogram of text[s .. n) plus min(s, n) padding zeros,
// s = 7 - p; no pass is skipped (an odd pass needs the byte its predecessor carried).
__global__ void __launch_bounds__(256)
rs_text_plan_kernel(SortArrays A, TextSort T, int nBlocks)
{
    const int b = blockIdx.x;
    const int d = threadIdx.x;
    const int n = A.cnt[b];
    u32* tot = A.totals + (i64)b * 2048;
    const u32 H = tot[d];
    const u8* __restrict__ text = blk_src(T.bt, T.st[b], b);
    for (int p = 0; p < 8; p++) {
        const int s = min(7 - p, n);
        u32 v = H;
        for (int x = 0; x < s; x++)
            v -= (text[x] == d) ? 1u : 0u;
        if (d == 0)
            v += (u32)s;
        tot[p * 256 + d] = v;
    }
    if (d == 0) {
        int w = A.which[b];
        for (int p = 0; p < 8; p++) {
            const int triv = (n <= 0) ? 1 : 0;
            A.trivial[p * A.maxBlocks + b] = triv;
            if (!triv)
                w ^= 1;
            A.which[(p + 1) * A.maxBlocks + b] = w;
        }
    }
}

// KIND 0: pass 0 (implicit index, text read in order); 1: odd pass (digit carried by the
// element); 2: even pass >= 2 (digit and the next pass's digit gathered from the text).
template <int KIND, int ITEMS>
__global__ void __launch_bounds__(RS_THREADS, (ITEMS <= 8) ? 4 : 3)
rs_onesweep_text_kernel(SortArrays A, TextSort T, int pass, int nBlocks, int tilesMax)
{
    constexpr int TILE = RS_THREADS * ITEMS;
    KNZ_DYN_SMEM(os_smem);
    u32* s_val = reinterpret_cast<u32*>(os_smem);
    u32(*s_cnt)[256] = reinterpret_cast<u32(*)[256]>(os_smem + TILE * 4);
    u32* s_delta = reinterpret_cast<u32*>(os_smem + TILE * 4 + (RS_THREADS / 32) * 1024);
    u32* s_toff = s_delta + 256;
    u32* s_w = s_toff + 256; // 8 words for the block scan + 1 for the ticket
    u8* s_dig = reinterpret_cast<u8*>(s_w + 16); // digit of every staged slot
    const int b = os_block_of_cta(nBlocks, tilesMax);
    const int cnt = A.cnt[b];
    if (cnt <= 0)
        return;
    const int tiles = (cnt + TILE - 1) / TILE;
    if (threadIdx.x == 0)
        s_w[8] = (u32)atomicAdd(&A.ticket[pass * A.maxBlocks + b], 1);
    for (int i = threadIdx.x; i < (RS_THREADS / 32) * 256; i += RS_THREADS)
        (&s_cnt[0][0])[i] = 0;
    __syncthreads();
    const int tile = (int)s_w[8];
    if (tile >= tiles)
        return;
    const int tbase = tile * TILE;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int src = A.which[pass * A.maxBlocks + b];
    const u32* __restrict__ vin = (src ? A.val[1] : A.val[0]) + (i64)b * A.capN;
    u32* __restrict__ vout = (src ? A.val[0] : A.val[1]) + (i64)b * A.capN;
    const u8* __restrict__ text = blk_src(T.bt, T.st[b], b);
    const int sft = 7 - pass; // digit = text[index + sft]
    const bool textAl = (((size_t)text) & 7) == 0;
    u32 val[ITEMS];
    u32 dgt[ITEMS]; // digit (9 bits) | rank inside the warp's elements << 9

    if (KIND != 0) {
#pragma unroll
        for (int it = 0; it < ITEMS; it++) {
            const int j = tbase + w * (32 * ITEMS) + it * 32 + lane;
            val[it] = (j < cnt) ? __ldg(&vin[j]) : 0;
        }
    }
#pragma unroll
    for (int it = 0; it < ITEMS; it++) {
        const int j = tbase + w * (32 * ITEMS) + it * 32 + lane;
        u32 d = 256u; // padding class
        if (j < cnt) {
            if (KIND == 1) {
                d = val[it] >> TX_IDX_BITS;
                val[it] &= TX_IDX_MASK;
            } else {
                const int idx = (KIND == 0) ? j : (int)(val[it] & TX_IDX_MASK);
                const int a = idx + sft;
                const int q = a - 1; // the next pass's digit sits right before this pass's
                u32 c;
                if (KIND == 2 && textAl && q >= 0 && (q & 7) != 7 && (q | 7) < cnt) {
                    // both bytes from one aligned 8-byte word: one L1 wavefront per lane instead of two
                    const u64 w8 = __ldg(reinterpret_cast<const u64*>(text + (q & ~7)));
                    const u32 two = (u32)(w8 >> ((q & 7) * 8));
                    c = two & 0xFFu;
                    d = (two >> 8) & 0xFFu;
                } else {
                    d = (a < cnt) ? (u32)__ldg(&text[a]) : 0u;
                    c = (q >= 0 && q < cnt) ? (u32)__ldg(&text[q]) : 0u;
                }
                val[it] = (u32)idx | (c << TX_IDX_BITS);
            }
        }
        dgt[it] = d;
    }
    // rank inside the warp's 256 elements (load order = stable order)
#pragma unroll
    for (int it = 0; it < ITEMS; it++) {
        const u32 d = dgt[it];
        const u32 peers = __match_any_sync(FULL_MASK, d);
        const u32 prior = (d < 256) ? s_cnt[w][d] : 0;
        __syncwarp();
        if (d < 256 && (peers & lanemask_lt()) == 0)
            s_cnt[w][d] = prior + __popc(peers);
        __syncwarp();
        dgt[it] = d | ((prior + (u32)__popc(peers & lanemask_lt())) << 9);
    }
    __syncthreads();
    // thread = digit: warp prefixes, tile count, look-back, offsets
    const int d = threadIdx.x;
    u32 tcount = 0;
#pragma unroll
    for (int x = 0; x < RS_THREADS / 32; x++) {
        const u32 v = s_cnt[x][d];
        s_cnt[x][d] = tcount;
        tcount += v;
    }
    u32* stw = A.hist + ((i64)b * A.maxTiles) * 256 + d;
    os_publish(stw + (i64)tile * 256, (tile == 0) ? (OS_INC | tcount) : (OS_AGG | tcount));
    u32 before = 0;
    for (int t = tile - 1; t >= 0;) {
        u32 v[OS_LOOK];
#pragma unroll
        for (int k = 0; k < OS_LOOK; k++)
            v[k] = (t - k >= 0) ? os_peek(stw + (i64)(t - k) * 256) : OS_INC;
        bool done = false;
#pragma unroll
        for (int k = 0; k < OS_LOOK; k++) {
            if (done || !(v[k] & (OS_INC | OS_AGG)))
                break;
            before += v[k] & OS_VAL;
            t--;
            done = (v[k] & OS_INC) != 0;
        }
        if (done)
            break;
    }
    if (tile > 0)
        os_publish(stw + (i64)tile * 256, OS_INC | (before + tcount));
    u32 tot;
    const u32 dbase = block_excl_sum_256(A.totals[(i64)b * 2048 + pass * 256 + d], s_w, &tot);
    const u32 toff = block_excl_sum_256(tcount, s_w, &tot);
    s_delta[d] = dbase + before - toff;
    s_toff[d] = toff;
    __syncthreads();
    // stage in sorted order (the element no longer holds this pass's digit: stage it alongside)
#pragma unroll
    for (int it = 0; it < ITEMS; it++) {
        const u32 dg = dgt[it] & 0x1FFu;
        if (dg < 256) {
            const u32 q = s_toff[dg] + s_cnt[w][dg] + (dgt[it] >> 9);
            s_val[q] = val[it];
            s_dig[q] = (u8)dg;
        }
    }
    __syncthreads();
    const int valid = min(TILE, cnt - tbase);
#pragma unroll
    for (int it = 0; it < ITEMS; it++) {
        const int q = it * RS_THREADS + threadIdx.x;
        if (q < valid)
            vout[s_delta[s_dig[q]] + (u32)q] = s_val[q];
    }
}

// Keys of the sorted suffixes for the grouping kernels: 8 text bytes, big-endian, zero padded.
__global__ void __launch_bounds__(256)
bwt_regen_keys_kernel(SortArrays A, TextSort T, int slowOnly)
{
    const int b = blockIdx.y;
    const int n = A.cnt[b];
    if (n <= 0)
        return;
    const int w8 = A.which[8 * A.maxBlocks + b];
    const u32* __restrict__ v = A.val[w8] + (i64)b * A.capN;
    u64* __restrict__ k = A.key[w8] + (i64)b * A.capN;
    const u8* __restrict__ text = blk_src(T.bt, T.st[b], b);
    const bool al = (((size_t)text) & 7) == 0;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const int i = (int)v[j];
        u64 key;
        const int b8 = i & ~7;
        if (al && b8 + 16 <= n && !slowOnly) {
            // two aligned 8-byte words cover text[i .. i+8)
#ifdef KNZ_REGEN_PLAIN
            const u64 w0 = *reinterpret_cast<const volatile u64*>(text + b8);
            const u64 w1 = *reinterpret_cast<const volatile u64*>(text + b8 + 8);
#else
            const u64 w0 = __ldg(reinterpret_cast<const u64*>(text + b8));
            const u64 w1 = __ldg(reinterpret_cast<const u64*>(text + b8 + 8));
#endif
            const int sh = (i & 7) * 8;
            const u64 le = sh ? ((w0 >> sh) | (w1 << (64 - sh))) : w0; // byte x of le = text[i + x]
            const u32 lo = (u32)le, hi = (u32)(le >> 32);
            key = ((u64)bswap32(lo) << 32) | (u64)bswap32(hi);
        } else {
            // the last few suffixes of the block (zero padded) and unaligned text: byte by byte
            u32 hi = 0, lo = 0;
#pragma unroll 1
            for (int x = 0; x < 8; x++) {
                const int p = i + x;
                u32 c = 0;
                if (p < n)
                    c = text[p];
                hi = (hi << 8) | (lo >> 24);
                lo = (lo << 8) | c;
            }
            key = ((u64)hi << 32) | lo;
        }
#ifdef KNZ_DEBUG_REGEN
        {
            u64 key2 = 0;
            for (int x = 0; x < 8; x++)
                key2 = (key2 << 8) | ((i + x < n) ? (u64)text[i + x] : 0ull);
            if (key2 != key)
                printf("regen mismatch j=%d i=%d n=%d fast=%llx slow=%llx\n", j, i, n, (unsigned long long)key,
                       (unsigned long long)key2);
        }
#endif
        k[j] = key;
    }
}

// debug (KNZ_TX_CHECK=1): recompute every key bytewise and report mismatches / order violations
__global__ void bwt_regen_check_kernel(SortArrays A, TextSort T)
{
    const int b = blockIdx.y;
    const int n = A.cnt[b];
    if (n <= 0)
        return;
    const int w8 = A.which[8 * A.maxBlocks + b];
    const u32* v = A.val[w8] + (i64)b * A.capN;
    const u64* k = A.key[w8] + (i64)b * A.capN;
    const u8* text = blk_src(T.bt, T.st[b], b);
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const int i = (int)v[j];
        u64 key2 = 0;
        for (int x = 0; x < 8; x++)
            key2 = (key2 << 8) | ((i + x < n) ? (u64)text[i + x] : 0ull);
        if (key2 != k[j])
            printf("KEY mismatch b=%d j=%d i=%d got=%llx want=%llx\n", b, j, i, (unsigned long long)k[j],
                   (unsigned long long)key2);
        if (j > 0 && k[j - 1] > k[j])
            printf("ORDER violation b=%d j=%d i=%d prev=%llx cur=%llx\n", b, j, i, (unsigned long long)k[j - 1],
                   (unsigned long long)k[j]);
        if (j > 0 && k[j - 1] == k[j] && v[j - 1] > v[j])
            printf("STABILITY violation b=%d j=%d i=%d prev i=%d\n", b, j, i, (int)v[j - 1]);
    }
}

static void radix_sort_text(const SortArrays& A, const TextSort& T, int nBlocks, int maxCnt, cudaStream_t s,
                            u64* launches)
{
    static int items = 0;
    if (!items) {
        const char* e = getenv("KNZ_TX_ITEMS"); // elements per thread of the index-only passes: 8 or 16
        items = (e && atoi(e) == 8) ? 8 : 16;
        cudaFuncSetAttribute(rs_onesweep_text_kernel<0, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, TX_SMEM(8));
        cudaFuncSetAttribute(rs_onesweep_text_kernel<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, TX_SMEM(8));
        cudaFuncSetAttribute(rs_onesweep_text_kernel<2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, TX_SMEM(8));
        cudaFuncSetAttribute(rs_onesweep_text_kernel<0, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, TX_SMEM(16));
        cudaFuncSetAttribute(rs_onesweep_text_kernel<1, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, TX_SMEM(16));
        cudaFuncSetAttribute(rs_onesweep_text_kernel<2, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, TX_SMEM(16));
    }
    const int tile = RS_THREADS * items;
    const int tiles = (maxCnt + tile - 1) / tile;
    if (tiles <= 0)
        return;
    cudaMemsetAsync(A.totals, 0, sizeof(u32) * 2048 * (size_t)nBlocks, s);
    cudaMemsetAsync(A.ticket, 0, sizeof(int) * 8 * (size_t)A.maxBlocks, s);
    KLAUNCH(rs_text_hist_kernel, dim3((maxCnt + RS_TILE * 4 - 1) / (RS_TILE * 4), nBlocks), RS_THREADS, s, A, T);
    KLAUNCH(rs_text_plan_kernel, nBlocks, 256, s, A, T, nBlocks);
    *launches += 2;
    for (int p = 0; p < 8; p++) {
        cudaMemset2DAsync(A.hist, sizeof(u32) * 256 * (size_t)A.maxTiles, 0, sizeof(u32) * 256 * (size_t)tiles,
                          (size_t)nBlocks, s);
        const int kind = (p == 0) ? 0 : (p & 1) ? 1 : 2;
#define TX_LAUNCH(K, I) KLAUNCH_DYN((rs_onesweep_text_kernel<K, I>), tiles * nBlocks, RS_THREADS, TX_SMEM(I), s, A, T, p, nBlocks, tiles)
        if (items == 16) {
            if (kind == 0)
                TX_LAUNCH(0, 16);
            else if (kind == 1)
                TX_LAUNCH(1, 16);
            else
                TX_LAUNCH(2, 16);
        } else {
            if (kind == 0)
                TX_LAUNCH(0, 8);
            else if (kind == 1)
                TX_LAUNCH(1, 8);
            else
                TX_LAUNCH(2, 8);
        }
#undef TX_LAUNCH
        *launches += 1;
    }
    const int gblocks = min((maxCnt + 255) / 256, 2048);
    static int dbgSlow = -1;
    if (dbgSlow < 0) {
        const char* e = getenv("KNZ_TX_REGEN_SLOW");
        dbgSlow = (e && atoi(e)) ? 1 : 0;
    }
    KLAUNCH(bwt_regen_keys_kernel, dim3(gblocks, nBlocks), 256, s, A, T, dbgSlow);
#ifndef KNZ_SIM
    static int dbgCheck = -1;
    if (dbgCheck < 0) {
        const char* e = getenv("KNZ_TX_CHECK");
        dbgCheck = (e && atoi(e)) ? 1 : 0;
    }
    if (dbgCheck) {
        bwt_regen_check_kernel<<<dim3(gblocks, nBlocks), 256, 0, s>>>(A, T);
        cudaStreamSynchronize(s);
    }
#endif
    *launches += 1;
}

// ------------------------------------------------------------------ forward
__global__ void bwt_decide_kernel(StageLaunch L, int* __restrict__ cnt, int* __restrict__ which0,
                                  int* __restrict__ bwtOk)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= L.nBlocks)
        return;
    const BlkState bs = L.stIn[b];
    BlkState ns = bs;
    const int n = bs.len;
    const int cap = (bs.swaps & 1) ? L.capOdd[b] : L.capEven[b];
    const int pisz = (n >= 1) ? bwt_pisz(n) : 0;
    // BWTBlockCodec.cpp:44-60: needs room for n + 33 and 1..4 index bytes
    const bool ok = (n >= 1) && (cap >= n + 33) && (pisz >= 1) && (pisz <= 4);
    if (ok) {
        ns.len = n + 1 + bwt_chunks(n) * pisz;
        ns.cur = next_cur(bs.cur);
        ns.swaps = bs.swaps + 1;
        ns.flags = bs.flags & ~(1 << (7 - L.stageIdx));
    }
    L.stOut[b] = ns;
    cnt[b] = ok ? n : 0;
    which0[b] = 0;
    bwtOk[b] = ok ? 1 : 0;
}

__global__ void __launch_bounds__(256)
bwt_init_keys_kernel(BufTable bt, const BlkState* __restrict__ st, const int* __restrict__ bwtOk, int capN,
                     u64* __restrict__ keyOut, u32* __restrict__ valOut)
{
    const int b = blockIdx.y;
    if (!bwtOk[b])
        return;
    const BlkState bs = st[b];
    const int n = bs.len;
    const u8* __restrict__ src = blk_src(bt, bs, b);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        u64 k = 0;
#pragma unroll
        for (int x = 0; x < 8; x++)
            k = (k << 8) | ((i + x < n) ? (u64)src[i + x] : 0ull);
        keyOut[(i64)b * capN + i] = k;
        valOut[(i64)b * capN + i] = (u32)i;
    }
}

// Flags of sorted element j:  head = starts a new equal-key run; ghead = starts
// a new OLD group (high key word changes; the initial round has one old group);
// resolved = its run has length 1.
struct GrpCtx {
    const u64* key[2];
    const u32* val[2];
    u32* valOut[2];
    u32* grpOut;
    u32* isa;
    const int* which; // [9][maxBlocks]; row 8 = after the sort
    int* whichNext;   // row 0
    const int* cnt;
    int* cntNext;
    u32* part; // [nBlocks][maxTiles][4]: maxH, maxG, sumU (then exclusive carries)
    int capN, maxTiles, maxBlocks, initial;
};

__device__ __forceinline__ void grp_flags(const u64* __restrict__ k, int j, int cnt, int initial, bool& head,
                                          bool& ghead, bool& resolved)
{
    const u64 kj = k[j];
    if (j == 0) {
        head = true;
        ghead = true;
    } else {
        const u64 kp = k[j - 1];
        head = kj != kp;
        ghead = !initial && ((kj >> 32) != (kp >> 32));
    }
    const bool nexthead = (j + 1 >= cnt) || (k[j + 1] != kj);
    resolved = head && nexthead;
}

// The grouping kernels give every thread RS_ITEMS consecutive sorted positions (the scans run in
// that order); the keys are brought in with coalesced loads and re-read from shared memory in the
// blocked order.  Slot i holds k[tbase - 1 + i] (one halo key on either side), padded one slot in
// eight so the blocked 64-bit reads are bank-conflict free.
#define GRP_PAD(i) ((i) + ((i) >> 3))
#define GRP_SK_SLOTS (GRP_PAD(RS_TILE + 2) + 1)
__device__ __forceinline__ void grp_stage_keys(const u64* __restrict__ k, int tbase, int cnt, u64* sk)
{
    for (int i = threadIdx.x; i < RS_TILE + 2; i += RS_THREADS) {
        const int j = tbase - 1 + i;
        sk[GRP_PAD(i)] = (j >= 0 && j < cnt) ? __ldg(&k[j]) : 0ull;
    }
    __syncthreads();
}
// flags of position j from its key and its neighbours' (kp = k[j-1], kn = k[j+1])
__device__ __forceinline__ void grp_flags_k(u64 kp, u64 kj, u64 kn, int j, int cnt, int initial, bool& head, bool& ghead,
                                            bool& resolved)
{
    if (j == 0) {
        head = true;
        ghead = true;
    } else {
        head = kj != kp;
        ghead = !initial && ((kj >> 32) != (kp >> 32));
    }
    const bool nexthead = (j + 1 >= cnt) || (kn != kj);
    resolved = head && nexthead;
}

__global__ void __launch_bounds__(RS_THREADS)
bwt_grp_partials_kernel(GrpCtx G)
{
    __shared__ u32 s_a[8], s_b[8], s_c[8];
    const int b = blockIdx.y;
    const int cnt = G.cnt[b];
    const int tbase = blockIdx.x * RS_TILE;
    if (tbase >= cnt)
        return;
    const u64* __restrict__ k = (G.which[8 * G.maxBlocks + b] ? G.key[1] : G.key[0]) + (i64)b * G.capN;
    __shared__ u64 s_k[GRP_SK_SLOTS];
    grp_stage_keys(k, tbase, cnt, s_k);
    u32 mh = 0, mg = 0, su = 0;
    const int j0 = tbase + threadIdx.x * RS_ITEMS;
    {
        const int i0 = threadIdx.x * RS_ITEMS + 1; // slot of j0
        u64 kp = s_k[GRP_PAD(i0 - 1)], kj = s_k[GRP_PAD(i0)];
#pragma unroll
        for (int x = 0; x < RS_ITEMS; x++) {
            const int j = j0 + x;
            const u64 kn = s_k[GRP_PAD(i0 + x + 1)];
            if (j < cnt) {
                bool head, ghead, res;
                grp_flags_k(kp, kj, kn, j, cnt, G.initial, head, ghead, res);
                if (head)
                    mh = (u32)j + 1;
                if (ghead)
                    mg = (u32)j + 1;
                su += res ? 0u : 1u;
            }
            kp = kj;
            kj = kn;
        }
    }
    // block reduce: max, max, sum
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mh = max(mh, __shfl_xor_sync(FULL_MASK, mh, o));
        mg = max(mg, __shfl_xor_sync(FULL_MASK, mg, o));
        su += __shfl_xor_sync(FULL_MASK, su, o);
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) {
        s_a[w] = mh;
        s_b[w] = mg;
        s_c[w] = su;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 a = 0, bb = 0, c = 0;
        for (int i = 0; i < 8; i++) {
            a = max(a, s_a[i]);
            bb = max(bb, s_b[i]);
            c += s_c[i];
        }
        u32* p = G.part + ((i64)b * G.maxTiles + blockIdx.x) * 4;
        p[0] = a;
        p[1] = bb;
        p[2] = c;
    }
}

// One warp per block: exclusive carries over tiles (max, max, sum); survivors; next source buffer.
__global__ void __launch_bounds__(32)
bwt_grp_scan_kernel(GrpCtx G, int nBlocks)
{
    const int b = blockIdx.x;
    const int lane = threadIdx.x;
    if (b >= nBlocks)
        return;
    const int cnt = G.cnt[b];
    const int after = G.which[8 * G.maxBlocks + b];
    if (lane == 0)
        G.whichNext[b] = after ^ 1; // survivors are compacted into the other buffer
    if (cnt <= 0) {
        if (lane == 0)
            G.cntNext[b] = 0;
        return;
    }
    const int tiles = (cnt + RS_TILE - 1) / RS_TILE;
    u32 ch = 0, cg = 0, cu = 0; // carries of everything before the current group of 32 tiles
    u32* p = G.part + (i64)b * G.maxTiles * 4;
    for (int t0 = 0; t0 < tiles; t0 += 32) {
        const int t = t0 + lane;
        u32 a = 0, bb = 0, c = 0;
        if (t < tiles) {
            a = p[4 * t];
            bb = p[4 * t + 1];
            c = p[4 * t + 2];
        }
        u32 ia = a, ib = bb, ic = c; // inclusive scans over the lanes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 ta = __shfl_up_sync(FULL_MASK, ia, o), tb = __shfl_up_sync(FULL_MASK, ib, o);
            const u32 tc = __shfl_up_sync(FULL_MASK, ic, o);
            if (lane >= o) {
                ia = max(ia, ta);
                ib = max(ib, tb);
                ic += tc;
            }
        }
        u32 ea = __shfl_up_sync(FULL_MASK, ia, 1), eb = __shfl_up_sync(FULL_MASK, ib, 1);
        u32 ec = __shfl_up_sync(FULL_MASK, ic, 1);
        if (lane == 0)
            ea = eb = ec = 0;
        if (t < tiles) {
            p[4 * t] = max(ch, ea);
            p[4 * t + 1] = max(cg, eb);
            p[4 * t + 2] = cu + ec;
        }
        ch = max(ch, __shfl_sync(FULL_MASK, ia, 31));
        cg = max(cg, __shfl_sync(FULL_MASK, ib, 31));
        cu += __shfl_sync(FULL_MASK, ic, 31);
    }
    if (lane == 0)
        G.cntNext[b] = (int)cu;
}

__global__ void __launch_bounds__(RS_THREADS, 4)
bwt_grp_apply_kernel(GrpCtx G)
{
    __shared__ u32 s_w[8], s_mh[8], s_mg[8];
    const int b = blockIdx.y;
    const int cnt = G.cnt[b];
    const int tbase = blockIdx.x * RS_TILE;
    if (tbase >= cnt)
        return;
    const int src = G.which[8 * G.maxBlocks + b];
    const u64* __restrict__ k = (src ? G.key[1] : G.key[0]) + (i64)b * G.capN;
    const u32* __restrict__ v = (src ? G.val[1] : G.val[0]) + (i64)b * G.capN;
    u32* __restrict__ vout = (src ? G.valOut[0] : G.valOut[1]) + (i64)b * G.capN;
    u32* __restrict__ gout = G.grpOut + (i64)b * G.capN;
    u32* __restrict__ isa = G.isa + (i64)b * G.capN;
    const u32* carry = G.part + ((i64)b * G.maxTiles + blockIdx.x) * 4;
    const int j0 = tbase + threadIdx.x * RS_ITEMS;
    __shared__ u64 s_k[GRP_SK_SLOTS];
    __shared__ u32 s_v[GRP_PAD(RS_TILE) + 1];
    for (int i = threadIdx.x; i < RS_TILE; i += RS_THREADS)
        s_v[GRP_PAD(i)] = (tbase + i < cnt) ? __ldg(&v[tbase + i]) : 0u;
    grp_stage_keys(k, tbase, cnt, s_k);
    bool head[RS_ITEMS], ghead[RS_ITEMS], res[RS_ITEMS];
    u32 oldHi[RS_ITEMS];
    u32 mh = 0, mg = 0, su = 0;
    {
        const int i0 = threadIdx.x * RS_ITEMS + 1; // slot of j0
        u64 kp = s_k[GRP_PAD(i0 - 1)], kj = s_k[GRP_PAD(i0)];
#pragma unroll
        for (int x = 0; x < RS_ITEMS; x++) {
            const int j = j0 + x;
            const u64 kn = s_k[GRP_PAD(i0 + x + 1)];
            head[x] = ghead[x] = false;
            res[x] = true;
            oldHi[x] = (u32)(kj >> 32);
            if (j < cnt) {
                grp_flags_k(kp, kj, kn, j, cnt, G.initial, head[x], ghead[x], res[x]);
                if (head[x])
                    mh = (u32)j + 1;
                if (ghead[x])
                    mg = (u32)j + 1;
                su += res[x] ? 0u : 1u;
            }
            kp = kj;
            kj = kn;
        }
    }
    // exclusive scans across threads: max (H), max (G), sum (U)
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    u32 ih = mh, ig = mg;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 th = __shfl_up_sync(FULL_MASK, ih, o), tg = __shfl_up_sync(FULL_MASK, ig, o);
        if (lane >= o) {
            ih = max(ih, th);
            ig = max(ig, tg);
        }
    }
    if (lane == 31) {
        s_mh[w] = ih;
        s_mg[w] = ig;
    }
    u32 eh = __shfl_up_sync(FULL_MASK, ih, 1), eg = __shfl_up_sync(FULL_MASK, ig, 1);
    if (lane == 0)
        eh = eg = 0;
    u32 totU;
    const u32 eu = block_excl_sum_256(su, s_w, &totU); // contains __syncthreads
    u32 bh = carry[0], bg = carry[1];
    for (int i = 0; i < w; i++) {
        bh = max(bh, s_mh[i]);
        bg = max(bg, s_mg[i]);
    }
    u32 H = max(bh, eh), Gm = max(bg, eg), U = carry[2] + eu;
#pragma unroll
    for (int x = 0; x < RS_ITEMS; x++) {
        const int j = j0 + x;
        if (j >= cnt)
            break;
        if (head[x])
            H = (u32)j + 1;
        if (ghead[x])
            Gm = (u32)j + 1;
        const u32 old = G.initial ? 0u : oldHi[x];
        const u32 rank = old + (H - Gm);
        const u32 sfx = s_v[GRP_PAD(threadIdx.x * RS_ITEMS + x)];
        // a survivor whose group did not split keeps the rank the round before wrote: skip the scattered
        // store (one 32 B sector each; in the late rounds, which only serve long repeats, that is most of them)
        if (G.initial || H != Gm)
            isa[sfx] = rank;
        if (!res[x]) {
            vout[U] = sfx;
            gout[U] = rank;
            U++;
        }
    }
}

// Re-key the survivors for doubling step h: (group rank, order of suffix i+h).
// Past-the-end neighbours sort before every real rank, longer overshoot first
// (end-of-string is the smallest symbol).
__global__ void __launch_bounds__(256)
bwt_gather_kernel(const BlkState* __restrict__ st, const int* __restrict__ cnt, const int* __restrict__ which0,
                  u64* key0, u64* key1, const u32* val0, const u32* val1, const u32* __restrict__ grp,
                  const u32* __restrict__ isa, int capN, int h)
{
    const int b = blockIdx.y;
    const int c = cnt[b];
    const int n = st[b].len;
    const int w = which0[b];
    u64* __restrict__ k = (w ? key1 : key0) + (i64)b * capN;
    const u32* __restrict__ v = (w ? val1 : val0) + (i64)b * capN;
    const u32* __restrict__ g = grp + (i64)b * capN;
    const u32* __restrict__ r = isa + (i64)b * capN;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < c; j += gridDim.x * blockDim.x) {
        const u32 i = v[j];
        const i64 nb = (i64)i + h;
        const u32 k2 = (nb < n) ? r[nb] + (u32)n : (u32)(n - 1) - i;
        k[j] = ((u64)g[j] << 32) | k2;
    }
}

__global__ void __launch_bounds__(256)
bwt_emit_kernel(BufTable bt, const BlkState* __restrict__ st, const int* __restrict__ bwtOk,
                const u32* __restrict__ isa, int capN)
{
    const int b = blockIdx.y;
    if (!bwtOk[b])
        return;
    const BlkState bs = st[b];
    const int n = bs.len;
    const u8* __restrict__ src = blk_src(bt, bs, b);
    u8* __restrict__ dst = blk_dst(bt, bs, b);
    const u32* __restrict__ r = isa + (i64)b * capN;
    const int chunks = bwt_chunks(n), pisz = bwt_pisz(n);
    const int hdr = 1 + chunks * pisz;
    const u32 r0 = r[0];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        // header: mode byte, then chunks x pisz big-endian bytes of (primary index - 1)
        dst[0] = (u8)((ilog2_u32((u32)chunks) << 2) | (pisz - 1));
        const int stp = n / chunks;
        const int step = (chunks * stp == n) ? stp : stp + 1;
        int q = 1;
        for (int c = 0; c < chunks; c++) {
            const i64 p = (i64)c * step;
            const u32 v = (p < n) ? r[p] : 0xFFFFFFFFu; // primary index - 1 = rank; unset index -> 0 - 1
            for (int sh = 8 * (pisz - 1); sh >= 0; sh -= 8)
                dst[q++] = (u8)(v >> sh);
        }
        dst[hdr] = src[n - 1];
    }
    for (int p = blockIdx.x * blockDim.x + threadIdx.x + 1; p < n; p += gridDim.x * blockDim.x) {
        const u32 rp = r[p];
        dst[hdr + rp + (rp < r0 ? 1 : 0)] = src[p - 1];
    }
}

// ---- rounds: tile-local sort + global sort of the groups that straddle tiles ----
// After the gather, survivors sit in group order with key (group rank << 32 | key2).
// Sorting every aligned 2048-element tile by key in shared memory (bitonic network;
// ties never matter: equal keys stay one group) finishes every group that lies inside
// one tile -- almost all of them.  Elements of groups that cross a tile boundary form
// the prefix / suffix of their tiles; they are copied out, radix sorted globally and
// written back to the same positions.
struct XCtx {
    u64* key[2];
    u32* val[2];
    const u32* grp;    // group rank by position (survivor order)
    const int* which0; // buffer holding the survivors
    int* which8;       // row read by the group kernels
    const int* cnt;
    u32* xcnt;         // [nBlocks][maxTiles][2] : npre, nsuf  -> exclusive offsets after the scan
    int* cntX;         // [nBlocks] extracted elements
    u64* xkey;         // extraction arrays (buffer 0 of the side sort)
    u32* xval;
    u32* xpos;
    u64* xkeyAlt;      // buffer 1 of the side sort
    u32* xvalAlt;
    const int* whichX8;
    int capN, maxTiles, maxBlocks;
};

__global__ void __launch_bounds__(RS_THREADS)
bwt_tile_sort_kernel(XCtx X)
{
    __shared__ u64 sk[RS_TILE];
    __shared__ u32 sv[RS_TILE];
    __shared__ u32 sg[RS_TILE];
    __shared__ u32 s_red[2][8];
    __shared__ u32 s_big;
    const int b = blockIdx.y;
    const int cnt = X.cnt[b];
    const int tbase = blockIdx.x * RS_TILE;
    if (tbase >= cnt)
        return;
    const int w = X.which0[b];
    u64* __restrict__ K = X.key[w] + (i64)b * X.capN;
    u32* __restrict__ V = X.val[w] + (i64)b * X.capN;
    const u32* __restrict__ G = X.grp + (i64)b * X.capN;
    const int tend = min(tbase + RS_TILE, cnt);
    const int tsz = tend - tbase;
    // groups that cross the tile edges are finished by the side sort: leave them alone here
    const u32 g0 = G[tbase], gl = G[tend - 1];
    const bool pre = (tbase > 0) && (G[tbase - 1] == g0);
    const bool suf = (tend < cnt) && (G[tend] == gl);
    for (int i = threadIdx.x; i < RS_TILE; i += RS_THREADS) {
        const int j = tbase + i;
        sk[i] = (j < cnt) ? K[j] : ~0ull;
        sv[i] = (j < cnt) ? V[j] : 0u;
        u32 g = (j < cnt) ? G[j] : 0xFFFFFFFFu;
        if ((pre && g == g0) || (suf && g == gl))
            g = 0xFFFFFFFEu - (u32)i; // unique id: never compared with a neighbour
        sg[i] = g;
    }
    if (threadIdx.x == 0)
        s_big = 0;
    __syncthreads();
    // largest group in the tile, as a power of two bound: a group has more than P
    // members iff some position i has the same group id at i + P
    u32 big = 0;
    for (int i = threadIdx.x; i < tsz; i += RS_THREADS) {
        const u32 g = sg[i];
#pragma unroll
        for (int lg = 1; lg <= 5; lg++)
            if (i + (1 << lg) < tsz && sg[i + (1 << lg)] == g)
                big |= 1u << lg;
    }
    if (big)
        atomicOr(&s_big, big);
    __syncthreads();
    const u32 bigAll = s_big;
    if (!(bigAll & 32u)) {
        // every group has <= 32 members: odd-even transposition inside groups
        const int phases = (bigAll & 16u) ? 32 : (bigAll & 8u) ? 16 : (bigAll & 4u) ? 8 : (bigAll & 2u) ? 4 : 2;
        for (int ph = 0; ph < phases; ph++) {
            for (int t = threadIdx.x; t < RS_TILE / 2; t += RS_THREADS) {
                const int i = 2 * t + (ph & 1);
                if (i + 1 < tsz && sg[i] == sg[i + 1]) {
                    const u64 a = sk[i], c = sk[i + 1];
                    if (a > c) {
                        sk[i] = c;
                        sk[i + 1] = a;
                        const u32 va = sv[i];
                        sv[i] = sv[i + 1];
                        sv[i + 1] = va;
                    }
                }
            }
            __syncthreads();
        }
    } else {
        for (int k = 2; k <= RS_TILE; k <<= 1) {
            for (int jj = k >> 1; jj > 0; jj >>= 1) {
                for (int t = threadIdx.x; t < RS_TILE / 2; t += RS_THREADS) {
                    const int i = ((t & ~(jj - 1)) << 1) | (t & (jj - 1));
                    const int p2 = i | jj;
                    const bool up = (i & k) == 0;
                    const u64 a = sk[i], c = sk[p2];
                    if ((a > c) == up) {
                        sk[i] = c;
                        sk[p2] = a;
                        const u32 va = sv[i];
                        sv[i] = sv[p2];
                        sv[p2] = va;
                    }
                }
                __syncthreads();
            }
        }
    }
    for (int i = threadIdx.x; i < RS_TILE; i += RS_THREADS) {
        const int j = tbase + i;
        if (j < cnt) {
            K[j] = sk[i];
            V[j] = sv[i];
        }
    }
    // straddling groups: leading run of the first group / trailing run of the last group
    u32 c0 = 0, c1 = 0;
    for (int j = tbase + threadIdx.x; j < tend; j += RS_THREADS) {
        const u32 g = G[j];
        c0 += (g == g0) ? 1u : 0u;
        c1 += (g == gl) ? 1u : 0u;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c0 += __shfl_xor_sync(FULL_MASK, c0, o);
        c1 += __shfl_xor_sync(FULL_MASK, c1, o);
    }
    if ((threadIdx.x & 31) == 0) {
        s_red[0][threadIdx.x >> 5] = c0;
        s_red[1][threadIdx.x >> 5] = c1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        u32 a = 0, c = 0;
        for (int i = 0; i < 8; i++) {
            a += s_red[0][i];
            c += s_red[1][i];
        }
        u32 npre = pre ? a : 0u;
        u32 nsuf = suf ? c : 0u;
        if (npre + nsuf > (u32)(tend - tbase)) // one group fills the tile and crosses both edges
            nsuf = (u32)(tend - tbase) - npre;
        u32* xc = X.xcnt + ((i64)b * X.maxTiles + blockIdx.x) * 2;
        xc[0] = npre;
        xc[1] = nsuf;
    }
}

// One warp per block: exclusive offsets of the tiles' extracted prefix / suffix runs.
__global__ void __launch_bounds__(32)
bwt_xscan_kernel(XCtx X, int nBlocks)
{
    const int b = blockIdx.x;
    const int lane = threadIdx.x;
    if (b >= nBlocks)
        return;
    if (lane == 0)
        X.which8[b] = X.which0[b];
    const int cnt = X.cnt[b];
    const int tiles = (cnt + RS_TILE - 1) / RS_TILE;
    u32 run = 0;
    u32* xc = X.xcnt + (i64)b * X.maxTiles * 2;
    for (int t0 = 0; t0 < tiles; t0 += 32) {
        const int t = t0 + lane;
        u32 a = 0, c = 0;
        if (t < tiles) {
            a = xc[2 * t];
            c = xc[2 * t + 1];
        }
        const u32 inc = warp_incl_sum(a + c, lane);
        const u32 ex = run + inc - (a + c);
        if (t < tiles) {
            xc[2 * t] = ex; // offset of the tile's prefix elements
            xc[2 * t + 1] = (ex + a) | ((c != 0) ? 0x80000000u : 0u); // offset of its suffix elements (+flag)
        }
        run += __shfl_sync(FULL_MASK, inc, 31);
    }
    if (lane == 0)
        X.cntX[b] = (int)run;
}

__global__ void __launch_bounds__(RS_THREADS)
bwt_extract_kernel(XCtx X, int writeBack)
{
    const int b = blockIdx.y;
    const int cnt = X.cnt[b];
    const int tbase = blockIdx.x * RS_TILE;
    if (tbase >= cnt || X.cntX[b] == 0)
        return;
    const int w = X.which0[b];
    u64* __restrict__ K = X.key[w] + (i64)b * X.capN;
    u32* __restrict__ V = X.val[w] + (i64)b * X.capN;
    const int tend = min(tbase + RS_TILE, cnt);
    const int tiles = (cnt + RS_TILE - 1) / RS_TILE;
    const u32* xc = X.xcnt + ((i64)b * X.maxTiles + blockIdx.x) * 2;
    const u32 offPre = xc[0];
    const u32 offSuf = xc[1] & 0x7FFFFFFFu;
    const u32 npre = offSuf - offPre;
    const u32 offNext = (blockIdx.x + 1 < (u32)tiles) ? xc[2] : (u32)X.cntX[b];
    const u32 nsuf = offNext - offSuf;
    const int wx = writeBack ? X.whichX8[b] : 0;
    u64* __restrict__ xk = (wx ? X.xkeyAlt : X.xkey) + (i64)b * X.capN;
    u32* __restrict__ xv = (wx ? X.xvalAlt : X.xval) + (i64)b * X.capN;
    for (u32 i = threadIdx.x; i < npre + nsuf; i += RS_THREADS) {
        const int j = (i < npre) ? (tbase + (int)i) : (tend - (int)nsuf + (int)(i - npre));
        const u32 x = (i < npre) ? (offPre + i) : (offSuf + (i - npre));
        if (writeBack) {
            K[j] = xk[x];
            V[j] = xv[x];
        } else {
            xk[x] = K[j];
            xv[x] = V[j];
        }
    }
}

void launch_bwt_forward(const StageLaunch& L, Workspace& ws, cudaStream_t s, u64* launches)
{
    const int nB = L.nBlocks;
    const int maxTiles = (ws.capN + RS_TILE - 1) / RS_TILE;
    int* whichRows = ws.which;
    KLAUNCH(bwt_decide_kernel, (nB + 63) / 64, 64, s, L, ws.cnt, whichRows, ws.bwtOk);
    // blocks of <= 4 MiB: index-only initial sort (digits gathered from the text); else keys travel
    static int textMode = -1;
    if (textMode < 0) {
        const char* e = getenv("KNZ_BWT_TEXTSORT");
        textMode = (e && atoi(e) == 0) ? 0 : 1;
    }
    // L.maxLen bounds the stage OUTPUT (input + 33 per BWT stage so far): indexes stay below 2^22
    const bool useText = textMode && (L.maxLen - 33 * (L.stageIdx + 1) <= (1 << TX_IDX_BITS));
    if (getenv("KNZ_VERBOSE"))
        fprintf(stderr, "bwt_forward: nB=%d maxLen=%d stage=%d text-mode=%d\n", nB, L.maxLen, L.stageIdx, (int)useText);
    *launches += 1;
    if (!useText) {
        const int initBlocks = min((L.maxLen + 255) / 256, 1024);
        KLAUNCH(bwt_init_keys_kernel, dim3(initBlocks, nB), 256, s, L.bt, L.stIn, ws.bwtOk, ws.capN, ws.keyA, ws.valA);
        *launches += 1;
    }

    SortArrays A;
    A.key[0] = ws.keyA;
    A.key[1] = ws.keyB;
    A.val[0] = ws.valA;
    A.val[1] = ws.valB;
    A.hist = ws.hist;
    A.ticket = reinterpret_cast<int*>(ws.digitBase);
    A.totals = ws.totals;
    A.which = ws.which;
    A.trivial = ws.trivial;
    A.cnt = ws.cnt;
    A.capN = ws.capN;
    A.maxTiles = maxTiles;
    A.maxBlocks = ws.maxBlocks;

    GrpCtx G;
    G.key[0] = ws.keyA;
    G.key[1] = ws.keyB;
    G.val[0] = ws.valA;
    G.val[1] = ws.valB;
    G.valOut[0] = ws.valA;
    G.valOut[1] = ws.valB;
    G.grpOut = ws.grpA;
    G.isa = ws.isa;
    G.which = ws.which;
    G.whichNext = ws.which; // row 0
    G.cnt = ws.cnt;
    G.cntNext = ws.cntNext;
    G.part = ws.scanA;
    G.capN = ws.capN;
    G.maxTiles = maxTiles;
    G.maxBlocks = ws.maxBlocks;

    SortArrays AX = A; // side sort of the extracted (tile-crossing) elements
    AX.key[0] = ws.xkeyA;
    AX.key[1] = ws.xkeyB;
    AX.val[0] = ws.xvalA;
    AX.val[1] = ws.xvalB;
    AX.which = ws.whichX;
    AX.cnt = ws.cntX;

    XCtx X;
    X.key[0] = ws.keyA;
    X.key[1] = ws.keyB;
    X.val[0] = ws.valA;
    X.val[1] = ws.valB;
    X.grp = ws.grpA;
    X.which0 = ws.which;
    X.which8 = ws.which + 8 * ws.maxBlocks;
    X.cnt = ws.cnt;
    X.xcnt = ws.scanB;
    X.cntX = ws.cntX;
    X.xkey = ws.xkeyA;
    X.xval = ws.xvalA;
    X.xpos = NULL;
    X.xkeyAlt = ws.xkeyB;
    X.xvalAlt = ws.xvalB;
    X.whichX8 = ws.whichX + 8 * ws.maxBlocks;
    X.capN = ws.capN;
    X.maxTiles = maxTiles;
    X.maxBlocks = ws.maxBlocks;

    int maxCnt = L.maxLen;
    // key2 < 2n, group rank < n  ->  which byte positions can be non-zero
    const int lowBits = ilog2_u32((u32)(2 * (i64)L.maxLen > 1 ? 2 * (i64)L.maxLen : 2)) + 1;
    const int highBits = ilog2_u32((u32)(L.maxLen > 1 ? L.maxLen : 2)) + 1;
    u32 roundMask = 0;
    for (int p = 0; p < 4; p++)
        if (8 * p < lowBits)
            roundMask |= 1u << p;
    for (int p = 0; p < 4; p++)
        if (8 * p < highBits)
            roundMask |= 1u << (4 + p);

    for (int round = 0, h = 8;; round++) {
        G.initial = (round == 0) ? 1 : 0;
        const int tiles = (maxCnt + RS_TILE - 1) / RS_TILE;
        if (round == 0) {
            if (useText) {
                TextSort T;
                T.bt = L.bt;
                T.st = L.stIn;
                radix_sort_text(A, T, nB, maxCnt, s, launches);
            } else {
                radix_sort(A, nB, maxCnt, 0xFFu, s, launches);
            }
        } else {
            // groups inside one tile: shared-memory sort; groups crossing tiles: side radix sort
            KLAUNCH(bwt_tile_sort_kernel, dim3(tiles, nB), RS_THREADS, s, X);
            KLAUNCH(bwt_xscan_kernel, nB, 32, s, X, nB);
            *launches += 2;
            cudaMemcpyAsync(ws.h_cnt, ws.cntX, sizeof(int) * nB, cudaMemcpyDeviceToHost, s);
            cudaStreamSynchronize(s);
            int maxX = 0;
            for (int b = 0; b < nB; b++)
                maxX = max(maxX, ws.h_cnt[b]);
            if (maxX > 0) {
                KLAUNCH(bwt_extract_kernel, dim3(tiles, nB), RS_THREADS, s, X, 0);
                cudaMemsetAsync(ws.whichX, 0, sizeof(int) * ws.maxBlocks, s); // extracted data starts in buffer 0
                radix_sort(AX, nB, maxX, roundMask, s, launches);
                KLAUNCH(bwt_extract_kernel, dim3(tiles, nB), RS_THREADS, s, X, 1);
                *launches += 2;
            }
        }
        KLAUNCH(bwt_grp_partials_kernel, dim3(tiles, nB), RS_THREADS, s, G);
        KLAUNCH(bwt_grp_scan_kernel, nB, 32, s, G, nB);
        KLAUNCH(bwt_grp_apply_kernel, dim3(tiles, nB), RS_THREADS, s, G);
        *launches += 3;
        // survivors per block -> host (sizes the next round's grids, detects the end)
        cudaMemcpyAsync(ws.h_cnt, ws.cntNext, sizeof(int) * nB, cudaMemcpyDeviceToHost, s);
        cudaMemcpyAsync(ws.cnt, ws.cntNext, sizeof(int) * nB, cudaMemcpyDeviceToDevice, s);
        cudaStreamSynchronize(s);
        maxCnt = 0;
        for (int b = 0; b < nB; b++)
            maxCnt = max(maxCnt, ws.h_cnt[b]);
        if (maxCnt == 0)
            break;
        if (round > 40) { // cannot happen: h exceeds any block length after 30 doublings
            int e = KERR_INTERNAL;
            cudaMemcpyAsync(L.errFlag, &e, sizeof(int), cudaMemcpyHostToDevice, s);
            break;
        }
        const int gblocks = min((maxCnt + 255) / 256, 2048);
        KLAUNCH(bwt_gather_kernel, dim3(gblocks, nB), 256, s, L.stIn, ws.cnt, ws.which, ws.keyA, ws.keyB, ws.valA, ws.valB,
                ws.grpA, ws.isa, ws.capN, h);
        *launches += 1;
        h = (h < (1 << 29)) ? h * 2 : h;
    }
    const int eblocks = min((L.maxLen + 255) / 256, 2048);
    KLAUNCH(bwt_emit_kernel, dim3(eblocks, nB), 256, s, L.bt, L.stIn, ws.bwtOk, ws.isa, ws.capN);
    *launches += 1;
}

// ------------------------------------------------------------------ inverse
#define BWT_END 0xFFFFFFFFu // psi of the end-of-string row: the text ends here
#define SPL_LOG 8           // one splitter every 256 sorted positions
// Header parse + decision (BWTBlockCodec.cpp:89-168, bsVersion 6 branch).
__global__ void bwt_inv_decide_kernel(StageLaunch L, int* __restrict__ cnt, int* __restrict__ which0,
                                      int* __restrict__ pidx, int* __restrict__ bwtOk)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= L.nBlocks)
        return;
    const BlkState bs = L.stIn[b];
    BlkState ns = bs;
    cnt[b] = 0;
    which0[b] = 0;
    bwtOk[b] = 0;
    if (bs.flags & (1 << (7 - L.stageIdx))) {
        L.stOut[b] = ns;
        return;
    }
    const u8* __restrict__ src = blk_src(L.bt, bs, b);
    const int n = bs.len;
    const int cap = (bs.swaps & 1) ? L.capOdd[b] : L.capEven[b];
    bool ok = n >= 2;
    int m = 0;
    if (ok) {
        const int mode = src[0];
        const int chunks = 1 << ((mode >> 2) & 7);
        const int pisz = (mode & 3) + 1;
        const int hdr = 1 + chunks * pisz;
        m = n - hdr;
        ok = (n >= hdr) && (chunks <= 8) && (m >= 1) && (chunks == bwt_chunks(m)) && (m <= cap);
        if (ok) {
            int q = 1;
            for (int c = 0; c < chunks; c++) {
                u32 v = 0;
                for (int x = 0; x < pisz; x++)
                    v = (v << 8) | src[q++];
                const i64 p = (i64)v + 1;
                // chunk starts that do not exist carry index 0 (encoded as -1): ignore them
                const int stp = m / chunks;
                const int step = (chunks * stp == m) ? stp : stp + 1;
                const bool used = (i64)c * step < m;
                if (used && (p < 1 || p > m))
                    ok = false;
                pidx[b * 8 + c] = used ? (int)p : 0;
            }
        }
    }
    if (!ok) {
        atomicExch(L.errFlag, KERR_BAD_STREAM);
        L.stOut[b] = ns;
        return;
    }
    ns.len = m;
    ns.cur = next_cur(bs.cur);
    ns.swaps = bs.swaps + 1;
    L.stOut[b] = ns;
    cnt[b] = m;
    bwtOk[b] = 1;
}

// psi walk, parallelised by list ranking with splitters (Helman-JaJa style):
//   nodes   = every 256th sorted position + the (<= 8) primary-index positions
//   walk    = every node follows psi to the next node and keeps the symbols of its segment
//             in a private slot: (next node, segment length)
//   rank    = one thread per primary index walks the ~n/256 nodes of its chunk and
//             hands every node its text offset
//   copy    = one warp per node moves the slot to its text offset.
// packed[t] = (psi[t] << 8) | F[t]  (one 4- or 8-byte gather per step).
struct InvCtx {
    BufTable bt;
    const BlkState* stIn;
    const int* bwtOk;
    const int* pidx;
    u64* packed[2]; // [0]: symbol slots of the walk, [1]: packed psi
    u32* symTotals; // [nBlocks][256]
    u32* status;    // one-sweep status words [nBlocks][statusStride][256]
    int* ticket;    // [nBlocks]
    int statusStride;
    u32* node; // [nBlocks][nodeStride][4]: next, len, base, pad
    int capN, nodeStride, nBlocks;
    int b0;     // first block of the group a walk launch covers (L2-sized groups)
    int narrow; // 1: packed entries are 32-bit ((psi << 8) | F, blocks < 16 MiB), else 64-bit
    int slot;   // bytes of symbol storage per node (multiple of 4)
    int* errFlag;
};

struct InvBlk {
    int m, hdr, chunks, step, S;
    u32 anchor[8];
};

__device__ __forceinline__ InvBlk inv_blk(const InvCtx& C, int b, const u8* src, int len)
{
    InvBlk B;
    const int mode = src[0];
    B.chunks = 1 << ((mode >> 2) & 7);
    B.hdr = 1 + B.chunks * ((mode & 3) + 1);
    B.m = len - B.hdr;
    const int stp = B.m / B.chunks;
    B.step = (B.chunks * stp == B.m) ? stp : stp + 1;
    B.S = (B.m + (1 << SPL_LOG) - 1) >> SPL_LOG;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const bool used = k < B.chunks && (i64)k * B.step < B.m;
        B.anchor[k] = used ? (u32)(C.pidx[b * 8 + k] - 1) : 0xFFFFFFFEu;
    }
    return B;
}

// node id of sorted position t, or -1.  Anchors take precedence over regular splitters.
__device__ __forceinline__ int inv_node_of(const InvBlk& B, u32 t)
{
#pragma unroll
    for (int k = 0; k < 8; k++)
        if (t == B.anchor[k])
            return B.S + k;
    if ((t & ((1u << SPL_LOG) - 1)) == 0)
        return (int)(t >> SPL_LOG);
    return -1;
}

__device__ __forceinline__ u32 inv_node_pos(const InvBlk& B, int id, bool* valid)
{
    if (id >= B.S) {
        const u32 a = B.anchor[id - B.S];
        *valid = a != 0xFFFFFFFEu;
        return a;
    }
    const u32 t = (u32)id << SPL_LOG;
    *valid = inv_node_of(B, t) == id; // a regular splitter that coincides with an anchor is represented by the anchor
    return t;
}

// Symbol totals of the L column, per block (the digit bases of the counting sort below).
__global__ void __launch_bounds__(256)
bwt_inv_hist_kernel(InvCtx C)
{
    __shared__ u32 s_h[256];
    const int b = blockIdx.y;
    if (!C.bwtOk[b])
        return;
    const BlkState bs = C.stIn[b];
    const u8* __restrict__ src = blk_src(C.bt, bs, b);
    const int mode = src[0];
    const int hdr = 1 + (1 << ((mode >> 2) & 7)) * ((mode & 3) + 1);
    const int m = bs.len - hdr;
    const int base = blockIdx.x * 8192;
    if (base >= m)
        return;
    s_h[threadIdx.x] = 0;
    __syncthreads();
    const int end = min(base + 8192, m);
    for (int i = base + threadIdx.x; i < end; i += 256)
        atomicAdd(&s_h[src[hdr + i]], 1u);
    __syncthreads();
    if (s_h[threadIdx.x])
        atomicAdd(&C.symTotals[(i64)b * 256 + threadIdx.x], s_h[threadIdx.x]);
}

// psi by ONE stable counting-sort pass over the L bytes (BWT.cpp:203-219 builds the same
// mapping with a serial bucket scan): sorted position of row i = C[L[i]] + #{i' < i : L[i'] = L[i]},
// and that position receives (row -> suffix-rank value of i) << 8 | L[i].  Same one-sweep
// structure as the radix pass (tickets, look-back status words, shared-memory staging),
// but the element is a byte on the way in and the packed psi entry on the way out:
// 1 + 4 bytes of traffic per symbol instead of an init / sort / pack chain over 12-byte pairs.
#define PSI_ITEMS 8
#define PSI_TILE (RS_THREADS * PSI_ITEMS)
template <bool NARROW>
__global__ void __launch_bounds__(RS_THREADS, 4)
bwt_inv_psi_kernel(InvCtx C)
{
    __shared__ u32 s_cnt[RS_THREADS / 32][256];
    __shared__ u32 s_delta[256], s_toff[256], s_w[9];
    __shared__ u64 s_out64[NARROW ? 1 : PSI_TILE];
    __shared__ u32 s_out32[NARROW ? PSI_TILE : 1];
    const int b = blockIdx.y;
    if (!C.bwtOk[b])
        return;
    const BlkState bs = C.stIn[b];
    const u8* __restrict__ src = blk_src(C.bt, bs, b);
    const int mode = src[0];
    const int hdr = 1 + (1 << ((mode >> 2) & 7)) * ((mode & 3) + 1);
    const int m = bs.len - hdr;
    const int tiles = (m + PSI_TILE - 1) / PSI_TILE;
    if ((int)blockIdx.x >= tiles)
        return;
    if (threadIdx.x == 0)
        s_w[8] = (u32)atomicAdd(&C.ticket[b], 1);
    for (int i = threadIdx.x; i < (RS_THREADS / 32) * 256; i += RS_THREADS)
        (&s_cnt[0][0])[i] = 0;
    __syncthreads();
    const int tile = (int)s_w[8];
    const int tbase = tile * PSI_TILE;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const u32 p0 = (u32)C.pidx[b * 8];
    u32 sym[PSI_ITEMS];
    u16 rnk[PSI_ITEMS];
#pragma unroll
    for (int it = 0; it < PSI_ITEMS; it++) {
        const int j = tbase + w * (32 * PSI_ITEMS) + it * 32 + lane;
        sym[it] = (j < m) ? (u32)src[hdr + j] : 256u; // 256 = padding class
    }
#pragma unroll
    for (int it = 0; it < PSI_ITEMS; it++) {
        const u32 d = sym[it];
        const u32 peers = __match_any_sync(FULL_MASK, d);
        const u32 prior = (d < 256) ? s_cnt[w][d] : 0;
        __syncwarp();
        if (d < 256 && (peers & lanemask_lt()) == 0)
            s_cnt[w][d] = prior + __popc(peers);
        __syncwarp();
        rnk[it] = (u16)(prior + __popc(peers & lanemask_lt()));
    }
    __syncthreads();
    const int d = threadIdx.x;
    u32 tcount = 0;
#pragma unroll
    for (int x = 0; x < RS_THREADS / 32; x++) {
        const u32 v = s_cnt[x][d];
        s_cnt[x][d] = tcount;
        tcount += v;
    }
    u32* st = C.status + ((i64)b * C.statusStride) * 256 + d;
    os_publish(st + (i64)tile * 256, (tile == 0) ? (OS_INC | tcount) : (OS_AGG | tcount));
    u32 before = 0;
    for (int t = tile - 1; t >= 0;) {
        u32 v[OS_LOOK];
#pragma unroll
        for (int k = 0; k < OS_LOOK; k++)
            v[k] = (t - k >= 0) ? os_peek(st + (i64)(t - k) * 256) : OS_INC;
        bool done = false;
#pragma unroll
        for (int k = 0; k < OS_LOOK; k++) {
            if (done || !(v[k] & (OS_INC | OS_AGG)))
                break;
            before += v[k] & OS_VAL;
            t--;
            done = (v[k] & OS_INC) != 0;
        }
        if (done)
            break;
    }
    if (tile > 0)
        os_publish(st + (i64)tile * 256, OS_INC | (before + tcount));
    u32 tot;
    const u32 dbase = block_excl_sum_256(C.symTotals[(i64)b * 256 + d], s_w, &tot);
    const u32 toff = block_excl_sum_256(tcount, s_w, &tot);
    s_delta[d] = dbase + before - toff;
    s_toff[d] = toff;
    __syncthreads();
#pragma unroll
    for (int it = 0; it < PSI_ITEMS; it++) {
        const int j = tbase + w * (32 * PSI_ITEMS) + it * 32 + lane;
        if (j < m) {
            const u32 q = s_toff[sym[it]] + s_cnt[w][sym[it]] + rnk[it];
            // row -> suffix rank: row 0 ends the text, rows < p0 hold rank row-1, rows >= p0 their index
            const u32 nxv = (j == 0) ? BWT_END : (((u32)j < p0) ? (u32)(j - 1) : (u32)j);
            if (NARROW)
                s_out32[q] = (nxv << 8) | sym[it]; // the end marker keeps its all-ones pattern in 24 bits
            else
                s_out64[q] = ((u64)nxv << 8) | (u64)sym[it];
        }
    }
    __syncthreads();
    const int valid = min(PSI_TILE, m - tbase);
    u64* __restrict__ P = C.packed[1] + (i64)b * C.capN;
#pragma unroll
    for (int it = 0; it < PSI_ITEMS; it++) {
        const int q = it * RS_THREADS + threadIdx.x;
        if (q < valid) {
            if (NARROW) {
                const u32 e = s_out32[q];
                reinterpret_cast<u32*>(P)[s_delta[e & 0xFF] + (u32)q] = e;
            } else {
                const u64 e = s_out64[q];
                P[s_delta[(u32)(e & 0xFF)] + (u32)q] = e;
            }
        }
    }
}

// entry t of the packed psi array -> (symbol, next position); BWT_END when the text ends here
__device__ __forceinline__ u32 inv_entry(const u64* __restrict__ P, bool narrow, u32 t, u32& nxt)
{
    if (narrow) {
        const u32 e = reinterpret_cast<const u32*>(P)[t];
        nxt = e >> 8;
        if (nxt == 0x00FFFFFFu)
            nxt = BWT_END;
        return e & 0xFF;
    }
    const u64 e = P[t];
    nxt = (u32)(e >> 8);
    return (u32)(e & 0xFF);
}

// ONE walk: every node follows psi to the next node and keeps the symbols it meets in its
// own INV_SLOT-byte slot (segments average 256 symbols; the 0.25% that overflow the slot
// are re-walked by the copy kernel from the position saved in nd[3]).
#define INV_SLOT 1536 // upper bound; small blocks get a smaller slot (InvCtx::slot)
__global__ void __launch_bounds__(128)
bwt_inv_walk_kernel(InvCtx C)
{
    const int b = C.b0 + blockIdx.y;
    if (b >= C.nBlocks || !C.bwtOk[b])
        return;
    const BlkState bs = C.stIn[b];
    const u8* __restrict__ src = blk_src(C.bt, bs, b);
    const InvBlk B = inv_blk(C, b, src, bs.len);
    const int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= B.S + 8)
        return;
    bool valid;
    u32 t = inv_node_pos(B, id, &valid);
    u32* nd = C.node + ((i64)b * C.nodeStride + id) * 4;
    if (!valid || t >= (u32)B.m) {
        nd[0] = 0xFFFFFFFFu;
        nd[1] = 0;
        nd[2] = 0xFFFFFFFFu;
        nd[3] = 0;
        return;
    }
    const u64* __restrict__ P = C.packed[1] + (i64)b * C.capN;
    u32* __restrict__ slot = reinterpret_cast<u32*>(C.packed[0] + (i64)b * C.capN) + (i64)id * (C.slot / 4);
    const u32 cap = (u32)C.slot;
    const bool narrow = C.narrow != 0;
    u32 len = 0, acc = 0, tcap = 0;
    int nxt = -1;
    for (;;) {
        u32 nx;
        const u32 sym = inv_entry(P, narrow, t, nx);
        if (len < cap) {
            acc |= sym << (8 * (len & 3));
            if ((len & 3) == 3) {
                slot[len >> 2] = acc;
                acc = 0;
            }
        }
        len++;
        t = nx;
        if (len == cap)
            tcap = t;
        if (t == BWT_END)
            break;
        if (t >= (u32)B.m || len > (u32)B.m) { // not a permutation: malformed input
            atomicExch(C.errFlag, KERR_BAD_STREAM);
            break;
        }
        nxt = inv_node_of(B, t);
        if (nxt >= 0)
            break;
    }
    if (len < cap && (len & 3))
        slot[len >> 2] = acc;
    nd[0] = (u32)nxt;
    nd[1] = len;
    nd[2] = 0xFFFFFFFFu;
    nd[3] = tcap;
}

// Segments to their text offsets: one warp per node copies the slot; the rare overflow
// tail is re-walked by lane 0.
__global__ void __launch_bounds__(256)
bwt_inv_copy_kernel(InvCtx C)
{
    const int b = C.b0 + blockIdx.y;
    if (b >= C.nBlocks || !C.bwtOk[b])
        return;
    const BlkState bs = C.stIn[b];
    const u8* __restrict__ src = blk_src(C.bt, bs, b);
    const InvBlk B = inv_blk(C, b, src, bs.len);
    const int lane = threadIdx.x & 31;
    const u64* __restrict__ P = C.packed[1] + (i64)b * C.capN;
    const u8* __restrict__ slots = reinterpret_cast<const u8*>(C.packed[0] + (i64)b * C.capN);
    const bool narrow = C.narrow != 0;
    u8* __restrict__ dst = blk_dst(C.bt, bs, b);
    const int warpsPerGrid = gridDim.x * (blockDim.x >> 5);
    for (int id = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); id < B.S + 8; id += warpsPerGrid) {
        const u32* nd = C.node + ((i64)b * C.nodeStride + id) * 4;
        const u32 len = nd[1], base = nd[2];
        if (len == 0 || base == 0xFFFFFFFFu || (u64)base + len > (u64)B.m)
            continue;
        const u32 cap = (u32)C.slot;
        const u32 n = len < cap ? len : cap;
        const u8* __restrict__ sp = slots + (i64)id * cap;
        u8* __restrict__ d = dst + base;
        for (u32 k = lane; k < n; k += 32)
            d[k] = sp[k];
        if (len > cap && lane == 0) {
            u32 t = nd[3];
            for (u32 k = cap; k < len; k++) {
                u32 nx;
                d[k] = (u8)inv_entry(P, narrow, t, nx);
                t = nx;
            }
        }
    }
}

// one thread per (block, chunk): hand out text offsets along the chunk's node chain
__global__ void bwt_inv_rank_kernel(InvCtx C)
{
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = gid >> 3, k = gid & 7;
    if (b >= C.nBlocks || !C.bwtOk[b])
        return;
    const BlkState bs = C.stIn[b];
    const u8* __restrict__ src = blk_src(C.bt, bs, b);
    const InvBlk B = inv_blk(C, b, src, bs.len);
    if (B.anchor[k] == 0xFFFFFFFEu)
        return;
    u32* nodes = C.node + (i64)b * C.nodeStride * 4;
    u32 off = (u32)((i64)k * B.step);
    const u32 endOff = (u32)min((i64)(k + 1) * B.step, (i64)B.m);
    int id = B.S + k;
    int guard = B.S + 16;
    while (guard-- > 0) {
        u32* nd = nodes + (i64)id * 4;
        nd[2] = off;
        off += nd[1];
        const u32 nx = nd[0];
        if (nx == 0xFFFFFFFFu || (int)nx >= B.S) // text end or the next chunk's anchor
            break;
        id = (int)nx;
    }
    if (off != endOff)
        atomicExch(C.errFlag, KERR_BAD_STREAM);
}

// Same ranking with the chunk chains walked in shared memory: one CTA per block stages (next, length)
// of its ~n/256 nodes (8 B each, 128 KiB for a 4 MiB block), then eight threads follow their chains
// at shared-memory latency (the global version pays a DRAM round trip per node: 4 ms per GiB).
#define INV_RANK_SMEM_MAX (200 * 1024)
__global__ void __launch_bounds__(256)
bwt_inv_rank_smem_kernel(InvCtx C)
{
    KNZ_DYN_SMEM(ir_smem);
    uint2* s_nd = reinterpret_cast<uint2*>(ir_smem);
    const int b = blockIdx.x;
    if (b >= C.nBlocks || !C.bwtOk[b])
        return;
    const BlkState bs = C.stIn[b];
    const u8* __restrict__ src = blk_src(C.bt, bs, b);
    const InvBlk B = inv_blk(C, b, src, bs.len);
    u32* nodes = C.node + (i64)b * C.nodeStride * 4;
    const int total = B.S + 8;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const uint2 v = *reinterpret_cast<const uint2*>(nodes + (i64)i * 4);
        s_nd[i] = v; // x = next node, y = symbols in this node's segment
    }
    __syncthreads();
    const int k = threadIdx.x;
    if (k >= 8 || B.anchor[k] == 0xFFFFFFFEu)
        return;
    u32 off = (u32)((i64)k * B.step);
    const u32 endOff = (u32)min((i64)(k + 1) * B.step, (i64)B.m);
    int id = B.S + k;
    int guard = B.S + 16;
    while (guard-- > 0) {
        const uint2 nd = s_nd[id];
        nodes[(i64)id * 4 + 2] = off;
        off += nd.y;
        if (nd.x == 0xFFFFFFFFu || (int)nd.x >= B.S) // text end or the next chunk's anchor
            break;
        id = (int)nd.x;
    }
    if (off != endOff)
        atomicExch(C.errFlag, KERR_BAD_STREAM);
}

void launch_bwt_inverse(const StageLaunch& L, Workspace& wsAll, cudaStream_t s, u64* launches)
{
    const int nB = L.nBlocks;
    const int maxTiles = (wsAll.capN + RS_TILE - 1) / RS_TILE;
    // this launch's slice of the per-block scratch (L.wsBlock0 .. + nB)
    Workspace ws = wsAll;
    {
        const i64 g0 = L.wsBlock0;
        ws.cnt += g0;
        ws.which += g0; // rows stay maxBlocks apart
        ws.pidx += g0 * 8;
        ws.bwtOk += g0;
        ws.keyA += g0 * wsAll.capN;
        ws.keyB += g0 * wsAll.capN;
        ws.totals += g0 * 256; // 256 symbol totals per block in this stage
        ws.hist += g0 * (i64)maxTiles * 256;
        ws.digitBase += g0; // one ticket per block
    }
    KLAUNCH(bwt_inv_decide_kernel, (nB + 63) / 64, 64, s, L, ws.cnt, ws.which, ws.pidx, ws.bwtOk);
    *launches += 1;
    InvCtx C;
    C.bt = L.bt;
    C.stIn = L.stIn;
    C.bwtOk = ws.bwtOk;
    C.pidx = ws.pidx;
    C.packed[0] = ws.keyA;
    C.packed[1] = ws.keyB;
    C.symTotals = ws.totals;
    C.status = ws.hist;
    C.statusStride = maxTiles;
    C.ticket = reinterpret_cast<int*>(ws.digitBase);
    C.node = ws.hist; // [maxBlocks][sTiles*256] u32 >= 4 * (capN/256 + 8) per block (after the psi pass)
    C.capN = ws.capN;
    C.nodeStride = maxTiles * 64;
    C.nBlocks = nB;
    C.b0 = 0;
    C.narrow = (L.maxLen < (1 << 24) - 1) ? 1 : 0;
    C.errFlag = L.errFlag;
    {
        const int ptiles = (L.maxLen + PSI_TILE - 1) / PSI_TILE;
        cudaMemsetAsync(ws.totals, 0, sizeof(u32) * 256 * (size_t)nB, s);
        cudaMemsetAsync(C.ticket, 0, sizeof(int) * (size_t)nB, s);
        cudaMemset2DAsync(C.status, sizeof(u32) * 256 * (size_t)maxTiles, 0, sizeof(u32) * 256 * (size_t)ptiles, (size_t)nB, s);
        KLAUNCH(bwt_inv_hist_kernel, dim3((L.maxLen + 8191) / 8192, nB), 256, s, C);
        if (C.narrow)
            KLAUNCH(bwt_inv_psi_kernel<true>, dim3(ptiles, nB), RS_THREADS, s, C);
        else
            KLAUNCH(bwt_inv_psi_kernel<false>, dim3(ptiles, nB), RS_THREADS, s, C);
        *launches += 2;
    }
    const int nodes = ((L.maxLen + 255) >> SPL_LOG) + 8;
    {
        // the slots live in the sort's input buffer (8 * capN bytes per block, free after the pack)
        const i64 maxNodes = ((i64)ws.capN >> SPL_LOG) + 9;
        const i64 fit = (8 * (i64)ws.capN / maxNodes) & ~(i64)3;
        C.slot = (int)(fit < INV_SLOT ? fit : INV_SLOT);
    }
    KLAUNCH(bwt_inv_walk_kernel, dim3((nodes + 127) / 128, nB), 128, s, C);
    {
        static bool attr = false;
        const size_t need = (size_t)(nodes + 8) * sizeof(uint2);
        if (need <= INV_RANK_SMEM_MAX) {
#ifndef KNZ_SIM
            if (!attr) {
                cudaFuncSetAttribute(bwt_inv_rank_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, INV_RANK_SMEM_MAX);
                attr = true;
            }
#endif
            KLAUNCH_DYN(bwt_inv_rank_smem_kernel, nB, 256, need, s, C);
        } else {
            KLAUNCH(bwt_inv_rank_kernel, (nB * 8 + 63) / 64, 64, s, C);
        }
        (void)attr;
    }
    KLAUNCH(bwt_inv_copy_kernel, dim3(min((nodes + 7) / 8, 1024), nB), 256, s, C);
    *launches += 3;
}

// ------------------------------------------------------------------ misc stages
__global__ void none_stage_kernel(StageLaunch L)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= L.nBlocks)
        return;
    // NullTransform copies (transform/NullTransform.hpp:45-65); a copy is
    // observable only through the swap parity and the cleared skip bit, so the
    // data stays where it is.
    BlkState ns = L.stIn[b];
    const int cap = (ns.swaps & 1) ? L.capOdd[b] : L.capEven[b];
    if (ns.len <= cap) {
        ns.swaps += 1;
        ns.flags &= ~(1 << (7 - L.stageIdx));
    }
    L.stOut[b] = ns;
}

void launch_none_forward(const StageLaunch& L, cudaStream_t s, u64* launches)
{
    KLAUNCH(none_stage_kernel, (L.nBlocks + 63) / 64, 64, s, L);
    *launches += 1;
}

__global__ void __launch_bounds__(256)
copy_out_kernel(BufTable bt, const BlkState* __restrict__ st, u8* __restrict__ out, i64 outStride, int outCap,
                int* __restrict__ errFlag)
{
    const int b = blockIdx.y;
    const BlkState bs = st[b];
    const u8* __restrict__ src = blk_src(bt, bs, b);
    u8* __restrict__ dst = out + (i64)b * outStride;
    const int n = bs.len;
    if (n > outCap) { // never write past the block's destination slot
        if (blockIdx.x == 0 && threadIdx.x == 0)
            atomicExch(errFlag, KERR_OUT_OVERFLOW);
        return;
    }
    if ((((uintptr_t)dst) & 15) == 0) {
        const int n16 = n >> 4;
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        uint4* d4 = reinterpret_cast<uint4*>(dst);
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x)
            d4[i] = s4[i];
        for (int i = (n16 << 4) + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
            dst[i] = src[i];
    } else {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
            dst[i] = src[i];
    }
}

void launch_copy_out(const BufTable& bt, const BlkState* st, int nBlocks, u8* out, i64 outStride, int outCap,
                     int* errFlag, cudaStream_t s, u64* launches)
{
    KLAUNCH(copy_out_kernel, dim3(64, nBlocks), 256, s, bt, st, out, outStride, outCap, errFlag);
    *launches += 1;
}

// ------------------------------------------------------------------ workspace
template <class T>
static bool wsalloc(T** p, i64 count)
{
    return cudaMalloc((void**)p, (size_t)(count > 0 ? count : 1) * sizeof(T)) == cudaSuccess;
}

bool workspace_alloc(Workspace& ws, int maxBlocks, int capN)
{
    memset(&ws, 0, sizeof(ws));
    ws.maxBlocks = maxBlocks;
    ws.capN = capN;
    const i64 nb = maxBlocks;
    const i64 zTiles = (capN + 4095) / 4096 + 1;
    bool ok = true;
    ws.tileWords = nb * zTiles * 8;
    ok = ok && wsalloc(&ws.tileA, ws.tileWords);
    ok = ok && wsalloc(&ws.tileB, ws.tileWords);
    ws.occWords = nb * zTiles * 512;
    ok = ok && wsalloc(&ws.occ, ws.occWords);
    ok = ok && wsalloc(&ws.zlen, nb);
    return ok;
}

// The suffix sorter's arrays (56 bytes per input byte of a batch) are only allocated by contexts
// that run a BWT stage: entropy-only pipelines on small blocks keep large batches cheap.
bool workspace_alloc_bwt(Workspace& ws)
{
    if (ws.bwtReady)
        return true;
    const i64 nb = ws.maxBlocks;
    const i64 elems = nb * ws.capN;
    const i64 sTiles = (ws.capN + RS_TILE - 1) / RS_TILE + 1;
    bool ok = true;
    ok = ok && wsalloc(&ws.keyA, elems);
    ok = ok && wsalloc(&ws.keyB, elems);
    ok = ok && wsalloc(&ws.valA, elems);
    ok = ok && wsalloc(&ws.valB, elems);
    ok = ok && wsalloc(&ws.grpA, elems);
    ok = ok && wsalloc(&ws.isa, elems);
    ok = ok && wsalloc(&ws.hist, nb * sTiles * 256);
    ok = ok && wsalloc(&ws.digitBase, nb * 256);
    ok = ok && wsalloc(&ws.totals, nb * 2048);
    ok = ok && wsalloc(&ws.which, 9 * nb);
    ok = ok && wsalloc(&ws.trivial, 8 * nb);
    ok = ok && wsalloc(&ws.cnt, nb);
    ok = ok && wsalloc(&ws.cntNext, nb);
    ok = ok && wsalloc(&ws.scanA, nb * sTiles * 4 + nb);
    ok = ok && wsalloc(&ws.scanB, nb * sTiles * 2 + nb);
    ok = ok && wsalloc(&ws.xkeyA, elems);
    ok = ok && wsalloc(&ws.xkeyB, elems);
    ok = ok && wsalloc(&ws.xvalA, elems);
    ok = ok && wsalloc(&ws.xvalB, elems);
    ok = ok && wsalloc(&ws.whichX, 9 * nb);
    ok = ok && wsalloc(&ws.cntX, nb);
    ok = ok && wsalloc(&ws.pidx, nb * 8);
    ok = ok && wsalloc(&ws.bwtOk, nb);
    ok = ok && (cudaMallocHost((void**)&ws.h_cnt, sizeof(int) * (size_t)nb) == cudaSuccess);
    ws.bwtReady = ok;
    return ok;
}

void workspace_free(Workspace& ws)
{
    void* d[] = { ws.tileA, ws.tileB, ws.occ, ws.zlen, ws.keyA, ws.keyB, ws.valA, ws.valB, ws.grpA, ws.isa, ws.hist,
                  ws.digitBase, ws.totals, ws.which, ws.trivial, ws.cnt, ws.cntNext, ws.scanA, ws.pidx, ws.bwtOk,
                  ws.scanB, ws.xkeyA, ws.xkeyB, ws.xvalA, ws.xvalB, ws.whichX, ws.cntX };
    for (size_t i = 0; i < sizeof(d) / sizeof(d[0]); i++)
        if (d[i])
            cudaFree(d[i]);
    if (ws.h_cnt)
        cudaFreeHost(ws.h_cnt);
    memset(&ws, 0, sizeof(ws));
}
