// srt.cu -- Sorted Rank Transform (kanzi SRT) on sm_100a.
//
// Reference: transform/SRT.cpp:22-109 (forward), :111-204 (inverse), preprocess :206-244,
// header :246-305.  The forward direction is a classic move-to-front whose list starts in the
// order of FIRST APPEARANCE of the symbols, every rank being written into the bucket of its own
// symbol; buckets are laid out by (frequency descending, symbol ascending) behind a header of 256
// varint frequencies.  Parallel formulation:
//   1. per 4 KiB tile: symbol counts and first positions (shared-memory atomics)
//   2. per block: totals, first-appearance order -> relabelling table, bucket starts (sorted by
//      frequency), per-tile bucket offsets (exclusive scan over tiles), header bytes
//   3. relabel the block by first-appearance index: SRT's rank stream is then exactly the
//      identity-initialised move-to-front of the relabelled bytes, i.e. kanzi's MTFT, which the
//      tile-parallel rank kernels of sbrt.cu already evaluate (last-two-occurrence tables + replay)
//   4. stable scatter of the ranks into their symbol's bucket (one warp per tile, rows of 32
//      positions ranked with match_any)
// The inverse consumes 256 interleaved per-symbol rank streams with a sequential automaton
// (SRT.cpp:164-198): one dependency chain per block, run by one lane with the list in shared memory.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

#define SRT_TILE 4096
#define SRT_HDR_MAX 1024 // getMaxEncodedLength = n + 1024 (transform/SRT.hpp:38)

__device__ __forceinline__ bool srt_applies(const StageLaunch& L, int b, const BlkState& bs)
{
    const int cap = (bs.swaps & 1) ? L.capOdd[b] : L.capEven[b];
    return bs.len > 0 && cap >= bs.len + SRT_HDR_MAX; // SRT.cpp:33-34
}

// 1. counts and first positions of every symbol inside a tile
__global__ void __launch_bounds__(256)
srt_tile_kernel(StageLaunch L, int maxTiles, u32* __restrict__ cntT, u32* __restrict__ firstT)
{
    __shared__ u32 s_cnt[256], s_first[256];
    const int b = blockIdx.y, t = blockIdx.x;
    const BlkState bs = L.stIn[b];
    if (!srt_applies(L, b, bs))
        return;
    const int n = bs.len;
    const int base = t * SRT_TILE;
    if (base >= n)
        return;
    const u8* __restrict__ src = blk_src(L.bt, bs, b);
    s_cnt[threadIdx.x] = 0;
    s_first[threadIdx.x] = 0xFFFFFFFFu;
    __syncthreads();
    const int end = min(base + SRT_TILE, n);
    for (int i = base + threadIdx.x; i < end; i += 256) {
        const u32 c = src[i];
        atomicAdd(&s_cnt[c], 1u);
        atomicMin(&s_first[c], (u32)i);
    }
    __syncthreads();
    const i64 o = ((i64)b * maxTiles + t) * 256 + threadIdx.x;
    cntT[o] = s_cnt[threadIdx.x];
    firstT[o] = s_first[threadIdx.x];
}

// Bucket start of symbol `sym` for frequencies f[256] in shared memory: buckets are ordered by
// (frequency descending, symbol ascending) -- the order SRT::preprocess sorts into (:206-244).
__device__ __forceinline__ u32 srt_bucket_start(const u32* f, int sym)
{
    const u32 fs = f[sym];
    u32 start = 0;
    for (int u = 0; u < 256; u++) {
        const u32 fu = f[u];
        if (fu != 0 && (fu > fs || (fu == fs && u < sym)))
            start += fu;
    }
    return start;
}

// 2. one CTA per block, thread = symbol
__global__ void __launch_bounds__(256)
srt_plan_kernel(StageLaunch L, int maxTiles, u32* __restrict__ cntT, const u32* __restrict__ firstT,
                u8* __restrict__ relabel, u32* __restrict__ bucketBase, BlkState* __restrict__ fakeIn,
                BlkState* __restrict__ fakeOut)
{
    __shared__ u32 s_f[256], s_first[256];
    __shared__ int s_hdr;
    const int b = blockIdx.x, sym = threadIdx.x;
    const BlkState bs = L.stIn[b];
    BlkState ns = bs;
    const bool ok = srt_applies(L, b, bs);
    BlkState fi, fo;
    fi.len = bs.len, fi.cur = 0, fi.swaps = 0, fi.flags = 0;
    fo = fi;
    fo.swaps = ok ? 1 : 0;
    if (sym == 0) {
        fakeIn[b] = fi;
        fakeOut[b] = fo;
    }
    if (!ok) {
        if (sym == 0)
            L.stOut[b] = ns; // stage refused: skip flag stays set
        return;
    }
    const int n = bs.len;
    const int tiles = (n + SRT_TILE - 1) / SRT_TILE;
    u32 tot = 0, first = 0xFFFFFFFFu;
    u32* ct = cntT + (i64)b * maxTiles * 256 + sym;
    const u32* ft = firstT + (i64)b * maxTiles * 256 + sym;
    for (int t = 0; t < tiles; t++) {
        const u32 c = ct[(i64)t * 256];
        ct[(i64)t * 256] = tot; // exclusive prefix over tiles: offset of the tile inside the bucket
        tot += c;
        first = min(first, ft[(i64)t * 256]);
    }
    s_f[sym] = tot;
    s_first[sym] = first;
    __syncthreads();
    // first-appearance index of this symbol (initial list position, SRT.cpp:41-60)
    u32 idx = 0;
    for (int u = 0; u < 256; u++)
        idx += (s_first[u] < first) ? 1u : 0u;
    relabel[b * 256 + sym] = (u8)idx;
    const u32 start = srt_bucket_start(s_f, sym);
    // header: 256 varints (SRT.cpp:246-277)
    u8* __restrict__ dst = blk_dst(L.bt, bs, b);
    if (sym == 0) {
        int p = 0;
        for (int i = 0; i < 256; i++) {
            u32 f = s_f[i];
            for (int k = 0; k < 4 && f >= 128; k++) {
                dst[p++] = (u8)(0x80 | f);
                f >>= 7;
            }
            dst[p++] = (u8)f;
        }
        s_hdr = p;
        ns.len = n + p;
        ns.cur = next_cur(bs.cur);
        ns.swaps = bs.swaps + 1;
        ns.flags = bs.flags & ~(1 << (7 - L.stageIdx));
        L.stOut[b] = ns;
    }
    __syncthreads();
    bucketBase[b * 256 + sym] = (u32)s_hdr + start;
}

// 3. tmp[i] = first-appearance index of src[i]
__global__ void __launch_bounds__(256)
srt_relabel_kernel(StageLaunch L, const BlkState* __restrict__ fakeOut, const u8* __restrict__ relabel,
                   u8* __restrict__ tmp, i64 tmpStride)
{
    __shared__ u8 s_map[256];
    const int b = blockIdx.y;
    if (fakeOut[b].swaps == 0)
        return;
    const BlkState bs = L.stIn[b];
    const int n = bs.len;
    s_map[threadIdx.x] = relabel[b * 256 + threadIdx.x];
    __syncthreads();
    const u8* __restrict__ src = blk_src(L.bt, bs, b);
    u8* __restrict__ o = tmp + (i64)b * tmpStride;
    const int n4 = ((((size_t)src) & 3) == 0) ? (n >> 2) : 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
        const u32 w = reinterpret_cast<const u32*>(src)[i];
        reinterpret_cast<u32*>(o)[i] = (u32)s_map[w & 0xFF] | ((u32)s_map[(w >> 8) & 0xFF] << 8) |
                                       ((u32)s_map[(w >> 16) & 0xFF] << 16) | ((u32)s_map[w >> 24] << 24);
    }
    for (int i = (n4 << 2) + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        o[i] = s_map[src[i]];
}

// 4. one warp per tile: ranks go to their symbol's bucket in text order
#define SRT_SC_WARPS 4
__global__ void __launch_bounds__(SRT_SC_WARPS * 32)
srt_scatter_kernel(StageLaunch L, const BlkState* __restrict__ fakeOut, int maxTiles, const u32* __restrict__ cntT,
                   const u32* __restrict__ bucketBase, const u8* __restrict__ ranks, i64 rankStride)
{
    __shared__ u32 s_pos[SRT_SC_WARPS][256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int t = blockIdx.x * SRT_SC_WARPS + w;
    if (fakeOut[b].swaps == 0)
        return;
    const BlkState bs = L.stIn[b];
    const int n = bs.len;
    const int base = t * SRT_TILE;
    if (base >= n)
        return;
    const u8* __restrict__ src = blk_src(L.bt, bs, b);
    u8* __restrict__ dst = blk_dst(L.bt, bs, b);
    const u8* __restrict__ rk = ranks + (i64)b * rankStride;
    u32* pos = s_pos[w];
    const u32* ct = cntT + ((i64)b * maxTiles + t) * 256;
    for (int i = lane; i < 256; i += 32)
        pos[i] = bucketBase[b * 256 + i] + ct[i];
    __syncwarp();
    const int end = min(base + SRT_TILE, n);
    for (int g = base; g < end; g += 32) {
        const int i = g + lane;
        const bool live = i < end;
        const u32 c = live ? (u32)src[i] : 256u + (u32)lane; // dead lanes match nobody
        const u32 m = __match_any_sync(FULL_MASK, c);
        if (live) {
            const u32 p = pos[c] + (u32)__popc(m & lanemask_lt());
            dst[p] = rk[i];
        }
        __syncwarp();
        if (live && (m & lanemask_lt()) == 0)
            pos[c] += (u32)__popc(m);
        __syncwarp();
    }
}

void launch_srt_forward(const StageLaunch& L, Workspace& ws, SrtWork& W, cudaStream_t s, u64* launches)
{
    const int maxTiles = (ws.capN + SRT_TILE - 1) / SRT_TILE;
    const int tiles = (L.maxLen + SRT_TILE - 1) / SRT_TILE;
    const int nB = L.nBlocks;
    KLAUNCH(srt_tile_kernel, dim3(tiles, nB), 256, s, L, maxTiles, W.cntT, W.firstT);
    KLAUNCH(srt_plan_kernel, nB, 256, s, L, maxTiles, W.cntT, W.firstT, W.relabel, W.bucketBase, W.fakeIn, W.fakeOut);
    const int rb = min((L.maxLen / 4 + 255) / 256 + 1, 256);
    KLAUNCH(srt_relabel_kernel, dim3(rb, nB), 256, s, L, W.fakeOut, W.relabel, W.tmp1, W.tmpStride);
    BufTable bt2;
    bt2.base[0] = W.tmp1;
    bt2.base[1] = W.tmp2;
    bt2.base[2] = W.tmp1;
    bt2.stride[0] = bt2.stride[1] = bt2.stride[2] = W.tmpStride;
    launch_sbrt_rank_only(bt2, W.fakeIn, W.fakeOut, nB, L.maxLen, 1, ws, s, launches);
    KLAUNCH(srt_scatter_kernel, dim3((tiles + SRT_SC_WARPS - 1) / SRT_SC_WARPS, nB), SRT_SC_WARPS * 32, s, L, W.fakeOut,
            maxTiles, W.cntT, W.bucketBase, W.tmp2, W.tmpStride);
    *launches += 4;
}

// ------------------------------------------------------------------ inverse
// One warp per block; lane 0 runs the automaton (SRT.cpp:111-204), the warp parses the header
// and sorts the buckets.
__global__ void __launch_bounds__(32)
srt_inverse_kernel(StageLaunch L)
{
    __shared__ u32 s_f[256];
    __shared__ u32 s_pos[256], s_end[256];
    __shared__ u8 s_r2s[256];
    __shared__ int s_hdr, s_ok;
    const int b = blockIdx.x, lane = threadIdx.x;
    const BlkState bs = L.stIn[b];
    BlkState ns = bs;
    const int bit = 1 << (7 - L.stageIdx);
    if (bs.flags & bit) { // stage was skipped by the encoder
        if (lane == 0)
            L.stOut[b] = ns;
        return;
    }
    const int cap = (bs.swaps & 1) ? L.capOdd[b] : L.capEven[b];
    const u8* __restrict__ src = blk_src(L.bt, bs, b);
    u8* __restrict__ dst = blk_dst(L.bt, bs, b);
    int length = bs.len;
    if (lane == 0) {
        int ok = (length >= 256) ? 1 : 0; // SRT.cpp:122
        int p = 0;
        for (int i = 0; i < 256 && ok; i++) { // decodeHeader :279-305
            u32 res = 0;
            int shift = 0;
            for (int j = 0; j < 5; j++) {
                if (p >= length) {
                    ok = 0;
                    break;
                }
                const u32 val = src[p++];
                res |= (val & 0x7F) << shift;
                if ((val & 0x80) == 0)
                    break;
                if (j == 4) {
                    ok = 0;
                    break;
                }
                shift += 7;
            }
            s_f[i] = res;
        }
        s_hdr = p;
        s_ok = ok;
    }
    __syncwarp();
    bool ok = s_ok != 0;
    const int hdr = s_hdr;
    length -= hdr;
    if (ok && (length < 0 || length > cap))
        ok = false;
    if (ok) {
        // bucket starts (every lane: 8 symbols), then the initial list: the first rank stored in a
        // bucket is its symbol's first-appearance index
        u64 total = 0;
        for (int k = 0; k < 8; k++) {
            const int sym = lane * 8 + k;
            const u32 st = srt_bucket_start(s_f, sym);
            s_pos[sym] = st;
            s_end[sym] = st + s_f[sym];
            total += s_f[sym];
        }
        for (int o = 16; o > 0; o >>= 1)
            total += __shfl_xor_sync(FULL_MASK, total, o);
        // (the reference does not compare the sum with the length; buckets past the end fail below)
        (void)total;
        __syncwarp();
        for (int i = lane; i < 256; i += 32)
            s_r2s[i] = 0;
        __syncwarp();
        if (lane == 0) {
            // symbols in bucket order: r2s[src[bucketPos]] = c (SRT.cpp:150-160)
            int nbSymbols = 0;
            for (int i = 0; i < 256; i++)
                nbSymbols += (s_f[i] != 0) ? 1 : 0;
            // walk the symbols in bucket order = increasing start among present symbols
            // (ties impossible: present symbols have distinct starts)
            int done = 0;
            u32 lastStart = 0;
            bool firstIter = true;
            while (done < nbSymbols && ok) {
                int best = -1;
                u32 bestStart = 0xFFFFFFFFu;
                for (int i = 0; i < 256; i++)
                    if (s_f[i] != 0) {
                        const u32 st = s_pos[i] & 0x7FFFFFFFu;
                        if ((firstIter || st > lastStart) && st < bestStart && !(s_pos[i] & 0x80000000u)) {
                            bestStart = st;
                            best = i;
                        }
                    }
                if (best < 0) {
                    ok = false;
                    break;
                }
                if ((i64)bestStart >= (i64)length) {
                    ok = false;
                    break;
                }
                s_r2s[src[hdr + bestStart]] = (u8)best;
                s_pos[best] = (bestStart + 1) | 0x80000000u; // consumed the first rank; mark visited
                lastStart = bestStart;
                firstIter = false;
                done++;
            }
            if (ok) {
                for (int i = 0; i < 256; i++)
                    s_pos[i] &= 0x7FFFFFFFu;
                const u8* __restrict__ rk = src + hdr;
                u32 c = s_r2s[0];
                for (int i = 0; i < length; i++) {
                    dst[i] = (u8)c;
                    const u32 p = s_pos[c];
                    if (p < s_end[c]) {
                        // ranks beyond the data are a malformed stream in the reference too (it would
                        // read past `length`); refuse instead
                        if ((i64)p >= (i64)length) {
                            ok = false;
                            break;
                        }
                        const u32 r = rk[p];
                        s_pos[c] = p + 1;
                        if (r == 0)
                            continue;
                        for (u32 k = 0; k < r; k++)
                            s_r2s[k] = s_r2s[k + 1];
                        s_r2s[r] = (u8)c;
                        c = s_r2s[0];
                    } else {
                        if (nbSymbols == 1)
                            continue;
                        nbSymbols--;
                        for (int k = 0; k < nbSymbols; k++)
                            s_r2s[k] = s_r2s[k + 1];
                        c = s_r2s[0];
                    }
                }
            }
            s_ok = ok ? 1 : 0;
        }
        __syncwarp();
        ok = s_ok != 0;
    }
    if (lane == 0) {
        if (ok) {
            ns.len = length;
            ns.cur = next_cur(bs.cur);
            ns.swaps = bs.swaps + 1;
        } else {
            atomicExch(L.errFlag, KERR_BAD_STREAM);
        }
        L.stOut[b] = ns;
    }
}

void launch_srt_inverse(const StageLaunch& L, cudaStream_t s, u64* launches)
{
    KLAUNCH(srt_inverse_kernel, L.nBlocks, 32, s, L);
    *launches += 1;
}

// ------------------------------------------------------------------ workspace
bool srt_work_alloc(SrtWork& W, int maxBlocks, int capN)
{
    memset(&W, 0, sizeof(W));
    const i64 nb = maxBlocks;
    const i64 maxTiles = (capN + SRT_TILE - 1) / SRT_TILE;
    W.tmpStride = ((i64)capN + 255) / 256 * 256;
    bool ok = true;
#define SRTALLOC(p, bytes) ok = ok && (cudaMalloc((void**)&(p), (size_t)(bytes)) == cudaSuccess)
    SRTALLOC(W.cntT, nb * maxTiles * 256 * sizeof(u32));
    SRTALLOC(W.firstT, nb * maxTiles * 256 * sizeof(u32));
    SRTALLOC(W.relabel, nb * 256);
    SRTALLOC(W.bucketBase, nb * 256 * sizeof(u32));
    SRTALLOC(W.fakeIn, nb * sizeof(BlkState));
    SRTALLOC(W.fakeOut, nb * sizeof(BlkState));
    SRTALLOC(W.tmp1, nb * W.tmpStride + 256);
    SRTALLOC(W.tmp2, nb * W.tmpStride + 256);
#undef SRTALLOC
    if (!ok)
        srt_work_free(W);
    return ok;
}

void srt_work_free(SrtWork& W)
{
    void* d[] = { W.cntT, W.firstT, W.relabel, W.bucketBase, W.fakeIn, W.fakeOut, W.tmp1, W.tmp2 };
    for (size_t i = 0; i < sizeof(d) / sizeof(d[0]); i++)
        if (d[i])
            cudaFree(d[i]);
    memset(&W, 0, sizeof(W));
}
