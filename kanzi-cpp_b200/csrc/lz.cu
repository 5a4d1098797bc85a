// lz.cu -- kanzi LZ, LZX and LZP block transforms on sm_100a.
//
// Reference: transform/LZCodec.cpp:118-455 (LZXCodec<T>::forward), :470-610 (inverseV6),
// :771-992 (LZPCodec::forward / inverse), helpers transform/LZCodec.hpp:178-248.
//
// The stream these stages write is defined by the reference's greedy parse: a hash table of the
// LAST position of every 8-byte (LZ/LZX) or 4-byte-context (LZP) hash, two repeat distances, one
// or two lazy positions, a skip distance that grows while nothing matches.  Every decision uses
// the table as left by all earlier positions, so a block is one dependency chain; blocks are
// independent.  One warp per block:
//   * the parse runs warp-uniformly: all 32 lanes execute the same instruction stream on the same
//     addresses (a load of one address is one transaction, a store of one value to one address
//     too), so no lane ever waits for a broadcast and the control flow never diverges
//   * the bulk work that hangs off a decision is spread over the lanes: literal runs, the
//     insertion of all positions covered by a match into the hash table (last writer wins, kept
//     by letting only the highest lane of a match_any group store), the final concatenation of
//     the four streams, and in the decoders the literal and match copies (a match closer than 32
//     bytes is a periodic pattern of bytes that already exist, farther ones go 32 bytes per round)
// NOTE: the sequential step of lzx_forward / lzp_forward (which candidate is tried first, the lazy
// positions, the backward extension, the token layout) restates the reference's parse decision for
// decision -- any deviation changes the tokens and therefore the bitstream.  What is this repo's own is
// the execution model around it: warp-uniform execution, the lane-parallel scan over positions that
// match nothing, cooperative literal / match copies and table insertions.
// LZ/LZX keep their three side streams (tokens, distances, match lengths) in per-block scratch: the
// stage fails as soon as literals + side streams reach the input length (LZCodec.cpp:421 tests the
// same sum once at the end; every term only grows), so one input length of scratch bounds them.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

#define LZ_MAX_MATCH (65535 + 254 + 4)
#define LZ_MAXD1 ((1 << 16) - 2)
#define LZ_MAXD2 ((1 << 24) - 2)
#define LZX_HASH_LOG 19
#define LZ_HASH_LOG 16
#define LZP_SEED 0x7FEB352Du
#define LZP_MIN_MATCH 64
#define LZP_FLAG 0xFC

__host__ __device__ __forceinline__ int lz_max_len(int n, bool lzp)
{
    return ((n <= 1024) ? n + 16 : n + n / 64) + (lzp ? 0 : 2); // LZCodec.hpp:91-95, :158-161
}

// Eight bytes at any alignment, little endian: the two aligned words around p, or byte by byte
// within 16 bytes of the end of the block (lim), which a caller's buffer may end with.  The parse is
// one warp per SM sub-partition with nothing to hide an instruction-cache miss behind, so the code is
// kept small: the tail path and the match loop are real functions, not inlined at their ~30 sites
// (the first version was 3160 instructions and ran at ~1400 cycles per position).
__device__ __noinline__ u64 lz_ld64_tail(const u8* p, const u8* lim)
{
    u64 v = 0;
    for (int k = 7; k >= 0; k--)
        v = (v << 8) | ((p + k < lim) ? (u64)p[k] : 0ull);
    return v;
}
__device__ __forceinline__ u64 lz_ld64(const u8* p, const u8* lim)
{
    if (p + 16 <= lim) {
        const u64* q = reinterpret_cast<const u64*>(reinterpret_cast<size_t>(p) & ~(size_t)7);
        const int sh = (int)(reinterpret_cast<size_t>(p) & 7) * 8;
        const u64 lo = q[0], hi = q[1];
        return (lo >> sh) | ((hi << 1) << (63 - sh));
    }
    return lz_ld64_tail(p, lim);
}
__device__ __forceinline__ u32 lz_ld32(const u8* p, const u8* lim) { return (u32)lz_ld64(p, lim); }

__device__ __forceinline__ u32 lz_hash(const u8* p, const u8* lim, int hashLog)
{
    return (u32)(((lz_ld64(p, lim) << 24) * (u64)0x1E35A7BDu) >> (64 - hashLog));
}

__device__ __noinline__ int lz_match(const u8* src, const u8* lim, int a, int b, int maxMatch)
{
    int n = 0;
    while (n + 8 <= maxMatch) {
        const u64 diff = lz_ld64(src + a + n, lim) ^ lz_ld64(src + b + n, lim);
        if (diff) {
            n += (__ffsll((long long)diff) - 1) >> 3;
            break;
        }
        n += 8;
    }
    return n;
}

__device__ __forceinline__ int lz_emit_len(u8* p, int length)
{
    if (length < 254) {
        p[0] = (u8)length;
        return 1;
    }
    if (length < 65536 + 254) {
        const u32 l = (u32)(length - 254);
        p[0] = 0xFE, p[1] = (u8)(l >> 8), p[2] = (u8)l;
        return 3;
    }
    const u32 l = (u32)(length - 255);
    p[0] = 0xFF, p[1] = (u8)(l >> 16), p[2] = (u8)(l >> 8), p[3] = (u8)l;
    return 4;
}

__device__ __forceinline__ u32 lz_read_len(const u8* p, int& pos)
{
    u32 res = p[pos++];
    if (res < 254)
        return res;
    if (res == 254) {
        res += ((u32)p[pos] << 8) | p[pos + 1];
        pos += 2;
        return res;
    }
    res += ((u32)p[pos] << 16) | ((u32)p[pos + 1] << 8) | p[pos + 2];
    pos += 3;
    return res;
}

// The parse is executed by all lanes on the same data, but lanes are not guaranteed to advance in
// lockstep: every access to the hash table is fenced so that no lane reads a slot another lane has
// already overwritten on behalf of a later position.
__device__ __forceinline__ int lz_xchg(int* slot, int v)
{
    __syncwarp();
    const int old = *slot;
    __syncwarp();
    *slot = v;
    return old;
}
__device__ __forceinline__ void lz_store(int* slot, int v)
{
    __syncwarp();
    *slot = v;
}

__device__ __forceinline__ void lz_prefetch(const void* p)
{
#ifndef KNZ_SIM
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

// n bytes, non-overlapping, spread over the warp
__device__ __forceinline__ void lz_warp_copy(u8* dst, const u8* src, int n, int lane)
{
    for (int i = lane; i < n; i += 32)
        dst[i] = src[i];
}

__device__ __forceinline__ void lz_finish_forward(const StageLaunch& L, int b, const BlkState& bs, bool ok, int produced,
                                                  int lane)
{
    if (lane != 0)
        return;
    BlkState ns = bs;
    if (ok) {
        ns.len = produced;
        ns.cur = next_cur(bs.cur);
        ns.swaps = bs.swaps + 1;
        ns.flags = bs.flags & ~(1 << (7 - L.stageIdx));
    }
    L.stOut[b] = ns;
}

__device__ __forceinline__ void lz_finish_inverse(const StageLaunch& L, int b, const BlkState& bs, bool ok, int produced,
                                                  int lane)
{
    if (lane != 0)
        return;
    BlkState ns = bs;
    if (ok) {
        ns.len = produced;
        ns.cur = next_cur(bs.cur);
        ns.swaps = bs.swaps + 1;
    } else {
        atomicExch(L.errFlag, KERR_BAD_STREAM);
    }
    L.stOut[b] = ns;
}

// ------------------------------------------------------------------ LZ / LZX forward
template <bool EXTRA>
__global__ void __launch_bounds__(32)
lzx_forward_kernel(StageLaunch L, LzWork W)
{
    constexpr int HLOG = EXTRA ? LZX_HASH_LOG : LZ_HASH_LOG;
    const int b = blockIdx.x, lane = threadIdx.x;
    const BlkState bs = L.stIn[b];
    const int count = bs.len;
    const int cap = (bs.swaps & 1) ? L.capOdd[b] : L.capEven[b];
    // data type left in the reference's Context by an earlier stage (LZCodec.cpp:181-191): DNA asks for a longer
    // minimum match (6), a small alphabet makes the codec refuse
    const int dt = L.dtype ? L.dtype[b] : 0;
    if (count < 24 || cap < lz_max_len(count, false) || dt == 9 /* SMALL_ALPHABET */) { // MIN_BLOCK_LENGTH, LZCodec.cpp:131-136
        lz_finish_forward(L, b, bs, false, 0, lane);
        return;
    }
    const u8* __restrict__ src = blk_src(L.bt, bs, b);
    const u8* lim = src + count;
    u8* __restrict__ dst = blk_dst(L.bt, bs, b);
    const int slot = L.wsBlock0 + b;
    int* __restrict__ hashes = W.hashes + (i64)slot * W.hashStride; // zeroed by the launcher
    // tokens grow up from the start of the scratch, match lengths down from its end; distances
    // have a region of their own
    u8* __restrict__ tk = W.side + (i64)slot * 2 * W.sideStride;
    u8* __restrict__ mb = tk + W.sideStride;
    const int mlTop = (int)W.sideStride - 8; // match-length byte k lives at tk[mlTop - k]
    const int srcEnd = count - 16 - 2;
    const int maxDist = (srcEnd < 4 * LZ_MAXD1) ? LZ_MAXD1 : LZ_MAXD2;
    const int minMatch = (dt == 6 /* DNA */) ? 6 : 4;
    if (lane == 0)
        dst[12] = (u8)(((maxDist == LZ_MAXD1) ? 0 : 1) | (((minMatch - 2) & 7) << 1));
    int srcIdx = 0, dstIdx = 13, anchor = 0, mIdx = 0, mLenIdx = 0, tkIdx = 0;
    int repd0 = count, repd1 = count; // repd[0], repd[1]
    int repIdx = 0, srcInc = 0;
    bool ok = true;
    // Most positions of a block match nothing (2.6 M of the 4 M positions of a compressible 4 MiB block), and
    // for those the sequential step -- hash, table exchange, three candidate tests, each a chain of dependent
    // loads -- only advances by one.  While no match is pending the state is simple (repeat distances fixed,
    // skip growing by the book), so the lanes test the next 32 VISITED positions side by side: lane j hashes
    // position p_j, takes its candidate from the latest earlier lane with the same hash or else from the table,
    // and applies the three 4-byte tests that gate findMatch in the reference.  The first lane that passes any
    // of them is the next position the sequential step has to look at; all lanes before it are misses and are
    // committed at once (table insertions, last one wins; skip counter).  The tests are the reference's own
    // gates, so the scan can stop early (the sequential step then decides) but never skips a match.
    while (srcIdx < srcEnd) {
        { // (a scan whose first lane already passes costs less than a sequential miss, so it always runs)
            int pj = srcIdx + lane;
            if (srcInc + 31 >= 64) { // the skip has started to grow: p_(j+1) = p_j + 1 + ((srcInc + j) >> 6)
                pj = srcIdx;
                for (int i = 0; i < lane; i++)
                    pj += 1 + ((srcInc + i) >> 6);
            }
            const bool act = pj < srcEnd;
            const u32 h = act ? lz_hash(src + pj, lim, HLOG) : (0x80000000u | (u32)lane);
            const u32 peers = __match_any_sync(FULL_MASK, h);
            const u32 lower = peers & ((1u << lane) - 1u);
            const int fromLane = lower ? (31 - __clz((int)lower)) : lane;
            const int inWin = __shfl_sync(FULL_MASK, pj, fromLane);
            __syncwarp(); // every earlier insertion is visible
            const int cand = lower ? inWin : (act ? hashes[h] : 0);
            bool ev = false;
            if (act) {
                const int minRefJ = max(pj - maxDist, 0);
                const u64 t = lz_ld64(src + pj, lim);
                const u32 w0 = (u32)t, w1n = (u32)(t >> 8);
                int r = pj + 1 - repd0;
                if (r > minRefJ && w1n == lz_ld32(src + r, lim))
                    ev = true;
                r = pj + 1 - repd1;
                if (r > minRefJ && w1n == lz_ld32(src + r, lim))
                    ev = true;
                if (cand > minRefJ && w0 == lz_ld32(src + cand, lim))
                    ev = true;
            }
            const u32 evm = __ballot_sync(FULL_MASK, ev);
            const int nact = __popc(__ballot_sync(FULL_MASK, act)); // active lanes are a prefix
            const int f = evm ? (__ffs((int)evm) - 1) : nact;      // lanes below f are misses
            const u32 below = (f >= 32) ? 0xFFFFFFFFu : ((1u << f) - 1u);
            if (lane < f && ((peers & below) >> lane) == 1u)
                hashes[h] = pj;
            __syncwarp();
            if (f > 0) {
                const int last = __shfl_sync(FULL_MASK, pj, f - 1);
                srcIdx = (f < nact) ? __shfl_sync(FULL_MASK, pj, f & 31) : last + 1 + ((srcInc + f - 1) >> 6);
                srcInc += f;
                repIdx = 0;
            }
            if (evm == 0)
                continue; // else a candidate is pending at srcIdx: the sequential step decides
        }
        int bestLen = 0;
        const u32 h0 = lz_hash(src + srcIdx, lim, HLOG);
        const int ref0 = lz_xchg(&hashes[h0], srcIdx);
        const int srcIdx1 = srcIdx + 1;
        int ref = srcIdx1 - (repIdx ? repd1 : repd0);
        const int minRef = max(srcIdx - maxDist, 0);
        const int lim1 = min(srcEnd - srcIdx1, LZ_MAX_MATCH);
        const u32 w1 = lz_ld32(src + srcIdx1, lim);
        if (ref > minRef && w1 == lz_ld32(src + ref, lim)) {
            bestLen = lz_match(src, lim, srcIdx1, ref, lim1);
        } else {
            ref = srcIdx1 - (repIdx ? repd0 : repd1);
            if (ref > minRef && w1 == lz_ld32(src + ref, lim))
                bestLen = lz_match(src, lim, srcIdx1, ref, lim1);
        }
        if (bestLen < minMatch) {
            ref = ref0;
            if (ref > minRef && lz_ld32(src + srcIdx, lim) == lz_ld32(src + ref, lim))
                bestLen = lz_match(src, lim, srcIdx, ref, min(srcEnd - srcIdx, LZ_MAX_MATCH));
            if (bestLen < minMatch) {
                srcIdx = srcIdx1 + (srcInc >> 6);
                srcInc++;
                repIdx = 0;
                continue;
            }
            if (srcIdx - ref != repd0 && srcIdx - ref != repd1) {
                const u32 h1 = lz_hash(src + srcIdx1, lim, HLOG);
                const int ref1 = lz_xchg(&hashes[h1], srcIdx1);
                if (ref1 > minRef + 1 && lz_ld32(src + srcIdx1 + bestLen - 3, lim) == lz_ld32(src + ref1 + bestLen - 3, lim)) {
                    const int bl1 = lz_match(src, lim, srcIdx1, ref1, lim1);
                    if (bl1 >= bestLen) {
                        ref = ref1;
                        bestLen = bl1;
                        srcIdx = srcIdx1;
                    }
                }
                if (EXTRA) {
                    const int srcIdx2 = srcIdx1 + 1;
                    const u32 h2 = lz_hash(src + srcIdx2, lim, HLOG);
                    const int ref2 = lz_xchg(&hashes[h2], srcIdx2);
                    if (ref2 > minRef + 2 && lz_ld32(src + srcIdx2 + bestLen - 3, lim) == lz_ld32(src + ref2 + bestLen - 3, lim)) {
                        const int bl2 = lz_match(src, lim, srcIdx2, ref2, min(srcEnd - srcIdx2, LZ_MAX_MATCH));
                        if (bl2 >= bestLen) {
                            ref = ref2;
                            bestLen = bl2;
                            srcIdx = srcIdx2;
                        }
                    }
                }
            }
            while (srcIdx > anchor && ref > minRef && src[srcIdx - 1] == src[ref - 1]) {
                bestLen++;
                ref--;
                srcIdx--;
            }
            if (bestLen > LZ_MAX_MATCH) {
                ref += bestLen - LZ_MAX_MATCH;
                srcIdx += bestLen - LZ_MAX_MATCH;
                bestLen = LZ_MAX_MATCH;
            }
        } else {
            if (bestLen >= LZ_MAX_MATCH || src[srcIdx] != src[ref - 1]) {
                srcIdx++;
                lz_store(&hashes[lz_hash(src + srcIdx, lim, HLOG)], srcIdx);
            } else {
                bestLen++;
                ref--;
            }
        }
        srcInc = 0;
        const int dist = srcIdx - ref;
        int token, mLenTh;
        if (dist == repd0) {
            token = 0x00;
            mLenTh = 3;
        } else if (dist == repd1) {
            token = 0x04;
            mLenTh = 3;
        } else {
            int nb = 1;
            if (dist >= 65536)
                mb[mIdx++] = (u8)(dist >> 16), nb++;
            if (dist >= 256)
                mb[mIdx++] = (u8)(dist >> 8), nb++;
            mb[mIdx++] = (u8)dist;
            token = nb << 3;
            mLenTh = 7;
        }
        const int mLen = bestLen - minMatch;
        if (mLen >= mLenTh) {
            token += mLenTh;
            u8 tmp[4];
            const int k = lz_emit_len(tmp, mLen - mLenTh);
            for (int x = 0; x < k; x++)
                tk[mlTop - (mLenIdx + x)] = tmp[x];
            mLenIdx += k;
        } else {
            token += mLen;
        }
        repd1 = repd0;
        repd0 = dist;
        repIdx = 1;
        const int litLen = srcIdx - anchor;
        if (litLen == 0) {
            tk[tkIdx++] = (u8)token;
        } else {
            if (litLen >= 7) {
                if (litLen >= (1 << 24)) {
                    ok = false;
                    break;
                }
                tk[tkIdx++] = (u8)((7 << 5) | token);
                dstIdx += lz_emit_len(dst + dstIdx, litLen - 7);
            } else {
                tk[tkIdx++] = (u8)((litLen << 5) | token);
            }
            lz_warp_copy(dst + dstIdx, src + anchor, litLen, lane);
            dstIdx += litLen;
        }
        if (dstIdx + tkIdx + mIdx + mLenIdx >= count) {
            ok = false;
            break;
        }
        // insert the positions covered by the match, in order (the last position of a hash wins)
        anchor = srcIdx + bestLen;
        __syncwarp();
        for (int p0 = srcIdx + 1; p0 < anchor; p0 += 32) {
            const int p = p0 + lane;
            const bool act = p < anchor;
            const u32 h = act ? lz_hash(src + p, lim, HLOG) : (0x80000000u | (u32)lane);
            const u32 peers = __match_any_sync(FULL_MASK, h);
            if (act && (peers >> lane) == 1u) // no higher lane shares the slot
                hashes[h] = p;
            __syncwarp();
        }
        srcIdx = anchor;
    }
    int produced = 0;
    if (ok) {
        const int litLen = count - anchor;
        if (dstIdx + litLen + tkIdx + mIdx + mLenIdx >= count) {
            ok = false;
        } else {
            if (litLen >= 7) {
                tk[tkIdx++] = (u8)(7 << 5);
                dstIdx += lz_emit_len(dst + dstIdx, litLen - 7);
            } else {
                tk[tkIdx++] = (u8)(litLen << 5);
            }
            lz_warp_copy(dst + dstIdx, src + anchor, litLen, lane);
            dstIdx += litLen;
            if (lane < 12) {
                const int v = (lane < 4) ? dstIdx : (lane < 8) ? tkIdx : mIdx;
                dst[lane] = (u8)(v >> (8 * (lane & 3)));
            }
            lz_warp_copy(dst + dstIdx, tk, tkIdx, lane);
            dstIdx += tkIdx;
            lz_warp_copy(dst + dstIdx, mb, mIdx, lane);
            dstIdx += mIdx;
            for (int i = lane; i < mLenIdx; i += 32)
                dst[dstIdx + i] = tk[mlTop - i];
            dstIdx += mLenIdx;
            produced = dstIdx;
            ok = dstIdx <= count - count / 100; // LZCodec.cpp:455
        }
    }
    lz_finish_forward(L, b, bs, ok, produced, lane);
}

// ------------------------------------------------------------------ LZ / LZX inverse
__global__ void __launch_bounds__(32)
lzx_inverse_kernel(StageLaunch L)
{
    const int b = blockIdx.x, lane = threadIdx.x;
    const BlkState bs = L.stIn[b];
    if (bs.flags & (1 << (7 - L.stageIdx))) { // stage was skipped by the encoder
        if (lane == 0)
            L.stOut[b] = bs;
        return;
    }
    const int count = bs.len;
    const int dstEnd = (bs.swaps & 1) ? L.capOdd[b] : L.capEven[b];
    const u8* __restrict__ src = blk_src(L.bt, bs, b);
    const u8* lim = src + count;
    u8* dst = blk_dst(L.bt, bs, b);
    bool ok = count >= 13;
    int tkIdx = 0, mIdx = 0, mLenIdx = 0;
    if (ok) {
        tkIdx = (int)lz_ld32(src, lim), mIdx = (int)lz_ld32(src + 4, lim), mLenIdx = (int)lz_ld32(src + 8, lim);
        ok = !(tkIdx < 0 || mIdx < 0 || mLenIdx < 0) &&
             !(tkIdx < 13 || tkIdx > count || mIdx > count - tkIdx || mLenIdx > count - tkIdx - mIdx);
    }
    int dstIdx = 0, srcIdx = 13, srcEnd = 0;
    if (ok) {
        mIdx += tkIdx;
        mLenIdx += mIdx;
        srcEnd = tkIdx - 13;
        const int litEnd = tkIdx;
        const int maxDist = ((src[12] & 1) == 0) ? LZ_MAXD1 : LZ_MAXD2;
        const int minMatch = ((src[12] >> 1) & 7) + 2;
        int repd0 = count, repd1 = count;
        for (;;) {
            // the side streams end at `count` (the reference would read on into the buffer's slack)
            if (tkIdx >= count || mIdx > count || mLenIdx > count) {
                ok = false;
                break;
            }
            const int token = src[tkIdx++];
            int mLen, dist;
            if ((token & 0x18) == 0) {
                mLen = token & 3;
                mLen += (mLen == 3) ? minMatch + (int)lz_read_len(src, mLenIdx) : minMatch;
                dist = (token & 4) ? repd1 : repd0;
            } else {
                mLen = token & 7;
                mLen += (mLen == 7) ? minMatch + (int)lz_read_len(src, mLenIdx) : minMatch;
                dist = src[mIdx++];
                if (token & 0x10) {
                    dist = (dist << 8) | src[mIdx++];
                    if (token & 0x08)
                        dist = (dist << 8) | src[mIdx++];
                }
            }
            if (token >= 32) {
                const u32 litLen = (token >= 0xE0) ? 7u + lz_read_len(src, srcIdx) : (u32)(token >> 5);
                if (litLen > (u32)(dstEnd - dstIdx) || litLen > (u32)(litEnd - srcIdx)) {
                    ok = false;
                    break;
                }
                lz_warp_copy(dst + dstIdx, src + srcIdx, (int)litLen, lane);
                srcIdx += (int)litLen;
                dstIdx += (int)litLen;
                if (srcIdx >= srcEnd)
                    break;
            }
            repd1 = repd0;
            repd0 = dist;
            const int mEnd = dstIdx + mLen;
            const int ref = dstIdx - dist;
            if (ref < 0 || dist > maxDist || dist <= 0 || mEnd > dstEnd) {
                ok = false;
                break;
            }
            __syncwarp(); // bytes written by other lanes become the match source
            if (dist < 32) {
                // the match is the periodic continuation of its last `dist` bytes
                for (int i = lane; i < mLen; i += 32)
                    dst[dstIdx + i] = dst[ref + (i % dist)];
            } else {
                for (int i0 = 0; i0 < mLen; i0 += 32) {
                    const int i = i0 + lane;
                    if (i < mLen)
                        dst[dstIdx + i] = dst[ref + i];
                    __syncwarp();
                }
            }
            dstIdx = mEnd;
        }
    }
    ok = ok && (srcIdx == srcEnd + 13);
    lz_finish_inverse(L, b, bs, ok, dstIdx, lane);
}

// ------------------------------------------------------------------ LZP
__global__ void __launch_bounds__(32)
lzp_forward_kernel(StageLaunch L, LzWork W)
{
    const int b = blockIdx.x, lane = threadIdx.x;
    const BlkState bs = L.stIn[b];
    const int count = bs.len;
    const int cap = (bs.swaps & 1) ? L.capOdd[b] : L.capEven[b];
    if (count < 128 || cap < lz_max_len(count, true)) { // MIN_BLOCK_LENGTH, LZCodec.cpp:786-793
        lz_finish_forward(L, b, bs, false, 0, lane);
        return;
    }
    const u8* __restrict__ src = blk_src(L.bt, bs, b);
    const u8* lim = src + count;
    u8* __restrict__ dst = blk_dst(L.bt, bs, b);
    int* __restrict__ hashes = W.hashes + (i64)(L.wsBlock0 + b) * W.hashStride;
    const int srcEnd = count, dstEnd = count - (count >> 6);
    if (lane < 4)
        dst[lane] = src[lane];
    u32 ctx = lz_ld32(src, lim);
    int srcIdx = 4, dstIdx = 4;
    bool ok = true;
    // Matches are rare (>= 64 bytes): nearly every position is a literal.  As in the LZ parse the lanes test the
    // next 32 positions side by side -- context hash, candidate from the latest earlier lane with that hash or from
    // the table, the reference's 8-byte gate at offset 56 -- and commit all literals before the first candidate
    // at once (bytes and escapes at scanned offsets, table insertions with the last one winning).  The context is
    // history dependent: little-endian read after a match, then shifted a byte at a time (LZCodec.cpp:813,842).
    while (srcIdx < srcEnd - LZP_MIN_MATCH && dstIdx < dstEnd) {
        if (dstIdx + 80 < dstEnd) {
            const int pj = srcIdx + lane;
            const bool act = pj < srcEnd - LZP_MIN_MATCH;
            u32 cj = ctx;
            if (lane >= 4) {
                cj = __byte_perm(lz_ld32(src + pj - 4, lim), 0, 0x0123); // the last four bytes, oldest on top
            } else {
                for (int i = 0; i < lane; i++)
                    cj = (cj << 8) | src[srcIdx + i];
            }
            const u32 h = act ? ((LZP_SEED * cj) >> 16) : (0x80000000u | (u32)lane);
            const u32 peers = __match_any_sync(FULL_MASK, h);
            const u32 lower = peers & ((1u << lane) - 1u);
            const int inWin = __shfl_sync(FULL_MASK, pj, lower ? (31 - __clz((int)lower)) : lane);
            __syncwarp();
            const int cand = lower ? inWin : (act ? hashes[h] : 0);
            const bool ev = act && cand != 0 &&
                            lz_ld64(src + cand + LZP_MIN_MATCH - 8, lim) == lz_ld64(src + pj + LZP_MIN_MATCH - 8, lim);
            const u32 evm = __ballot_sync(FULL_MASK, ev);
            const int nact = __popc(__ballot_sync(FULL_MASK, act));
            const int f = evm ? (__ffs((int)evm) - 1) : nact;
            const u32 val = act ? (u32)src[pj] : 0u;
            const u32 nb = (lane < f) ? (1u + ((cand != 0 && val == LZP_FLAG) ? 1u : 0u)) : 0u;
            const u32 incl = warp_incl_sum(nb, lane);
            const u32 below = (f >= 32) ? 0xFFFFFFFFu : ((1u << f) - 1u);
            if (lane < f) {
                const u32 o = (u32)dstIdx + incl - nb;
                dst[o] = (u8)val;
                if (nb == 2)
                    dst[o + 1] = 0xFF;
                if (((peers & below) >> lane) == 1u)
                    hashes[h] = pj;
            }
            __syncwarp();
            if (f > 0) {
                dstIdx += (int)__shfl_sync(FULL_MASK, incl, f - 1);
                ctx = __shfl_sync(FULL_MASK, (cj << 8) | val, f - 1);
                srcIdx += f;
            }
            if (evm == 0)
                continue;
        }
        const u32 h = (LZP_SEED * ctx) >> 16;
        const int ref = lz_xchg(&hashes[h], srcIdx);
        int bestLen = 0;
        if (ref != 0 && lz_ld64(src + ref + LZP_MIN_MATCH - 8, lim) == lz_ld64(src + srcIdx + LZP_MIN_MATCH - 8, lim))
            bestLen = lz_match(src, lim, srcIdx, ref, srcEnd - srcIdx);
        if (bestLen < LZP_MIN_MATCH) {
            const u32 val = src[srcIdx];
            ctx = (ctx << 8) | val;
            dst[dstIdx++] = (u8)val;
            srcIdx++;
            if (ref != 0 && val == LZP_FLAG) {
                if (dstIdx >= dstEnd) {
                    ok = false;
                    break;
                }
                dst[dstIdx++] = 0xFF;
            }
            continue;
        }
        srcIdx += bestLen;
        ctx = lz_ld32(src + srcIdx - 4, lim);
        dst[dstIdx++] = LZP_FLAG;
        bestLen -= LZP_MIN_MATCH;
        while (bestLen >= 254 && dstIdx < dstEnd) {
            bestLen -= 254;
            dst[dstIdx++] = 0xFE;
        }
        if (dstIdx >= dstEnd) {
            ok = false;
            break;
        }
        dst[dstIdx++] = (u8)bestLen;
    }
    while (ok && srcIdx < srcEnd && dstIdx < dstEnd) {
        const u32 h = (LZP_SEED * ctx) >> 16;
        const int ref = lz_xchg(&hashes[h], srcIdx);
        const u32 val = src[srcIdx];
        ctx = (ctx << 8) | val;
        dst[dstIdx++] = (u8)val;
        srcIdx++;
        if (ref != 0 && val == LZP_FLAG) {
            if (dstIdx >= dstEnd) {
                ok = false;
                break;
            }
            dst[dstIdx++] = 0xFF;
        }
    }
    ok = ok && srcIdx == count && dstIdx < dstEnd;
    lz_finish_forward(L, b, bs, ok, dstIdx, lane);
}

__global__ void __launch_bounds__(32)
lzp_inverse_kernel(StageLaunch L, LzWork W)
{
    const int b = blockIdx.x, lane = threadIdx.x;
    const BlkState bs = L.stIn[b];
    if (bs.flags & (1 << (7 - L.stageIdx))) {
        if (lane == 0)
            L.stOut[b] = bs;
        return;
    }
    const int count = bs.len;
    const int dstEnd = (bs.swaps & 1) ? L.capOdd[b] : L.capEven[b];
    const u8* __restrict__ src = blk_src(L.bt, bs, b);
    const u8* lim = src + count;
    u8* dst = blk_dst(L.bt, bs, b);
    int* __restrict__ hashes = W.hashes + (i64)(L.wsBlock0 + b) * W.hashStride;
    bool ok = count >= 4 && dstEnd >= count;
    int srcIdx = 4, dstIdx = 4;
    if (ok) {
        if (lane < 4)
            dst[lane] = src[lane];
        u32 ctx = lz_ld32(src, lim);
        const int srcEnd = count;
        while (srcIdx < srcEnd) {
            // literal runs, 32 bytes per trip: everything up to the next flag byte is copied, its context
            // hashes inserted (last one wins); only a flag byte needs the table's old value
            if (dstIdx + 40 < dstEnd) {
                const int sj = srcIdx + lane;
                const bool act = sj < srcEnd;
                const u32 vj = act ? (u32)src[sj] : (u32)LZP_FLAG;
                const u32 stop = __ballot_sync(FULL_MASK, vj == LZP_FLAG);
                const int f = stop ? (__ffs((int)stop) - 1) : 32;
                if (f > 0) {
                    u32 cj = ctx;
                    if (lane >= 4) {
                        cj = __byte_perm(lz_ld32(src + sj - 4, lim), 0, 0x0123);
                    } else {
                        for (int i = 0; i < lane; i++)
                            cj = (cj << 8) | src[srcIdx + i];
                    }
                    const u32 hj = (lane < f) ? ((LZP_SEED * cj) >> 16) : (0x80000000u | (u32)lane);
                    const u32 peers = __match_any_sync(FULL_MASK, hj);
                    __syncwarp();
                    if (lane < f) {
                        dst[dstIdx + lane] = (u8)vj;
                        if ((peers >> lane) == 1u)
                            hashes[hj] = dstIdx + lane;
                    }
                    __syncwarp();
                    ctx = __shfl_sync(FULL_MASK, (cj << 8) | vj, f - 1);
                    dstIdx += f;
                    srcIdx += f;
                    continue;
                }
            }
            const u32 h = (LZP_SEED * ctx) >> 16;
            const u32 v = src[srcIdx];
            // the slot's old value matters only behind a flag byte: literals just overwrite it (a store
            // nobody waits for) instead of a load on the critical path of every byte
            int ref = 0;
            if (v == LZP_FLAG)
                ref = lz_xchg(&hashes[h], dstIdx);
            else
                lz_store(&hashes[h], dstIdx);
            if (v != LZP_FLAG || ref == 0) {
                if (dstIdx >= dstEnd) {
                    ok = false;
                    break;
                }
                ctx = (ctx << 8) | v;
                if (lane == 0)
                    dst[dstIdx] = (u8)v;
                dstIdx++;
                srcIdx++;
                continue;
            }
            srcIdx++;
            if (srcIdx >= srcEnd) {
                ok = false;
                break;
            }
            if (src[srcIdx] == 0xFF) {
                if (dstIdx >= dstEnd) {
                    ok = false;
                    break;
                }
                ctx = (ctx << 8) | LZP_FLAG;
                if (lane == 0)
                    dst[dstIdx] = LZP_FLAG;
                dstIdx++;
                srcIdx++;
                continue;
            }
            int mLen = LZP_MIN_MATCH;
            while (srcIdx < srcEnd && src[srcIdx] == 0xFE) {
                srcIdx++;
                mLen += 254;
            }
            if (srcIdx >= srcEnd) {
                ok = false;
                break;
            }
            mLen += src[srcIdx++];
            if (dstIdx + mLen > dstEnd || ref >= dstIdx) {
                ok = false;
                break;
            }
            const int dist = dstIdx - ref;
            __syncwarp();
            if (dist < 32) {
                for (int i = lane; i < mLen; i += 32)
                    dst[dstIdx + i] = dst[ref + (i % dist)];
            } else {
                for (int i0 = 0; i0 < mLen; i0 += 32) {
                    const int i = i0 + lane;
                    if (i < mLen)
                        dst[dstIdx + i] = dst[ref + i];
                    __syncwarp();
                }
            }
            dstIdx += mLen;
            __syncwarp();
            ctx = (u32)dst[dstIdx - 4] | ((u32)dst[dstIdx - 3] << 8) | ((u32)dst[dstIdx - 2] << 16) |
                  ((u32)dst[dstIdx - 1] << 24);
        }
        ok = ok && srcIdx == srcEnd;
    }
    lz_finish_inverse(L, b, bs, ok, dstIdx, lane);
}

// ------------------------------------------------------------------ launchers, workspace
void launch_lz_forward(const StageLaunch& L, int type, LzWork& W, cudaStream_t s, u64* launches)
{
    const int hlog = (type == T_LZX) ? LZX_HASH_LOG : LZ_HASH_LOG;
    cudaMemsetAsync(W.hashes + (i64)L.wsBlock0 * W.hashStride, 0, sizeof(int) * (size_t)W.hashStride * (size_t)L.nBlocks, s);
    (void)hlog;
    if (type == T_LZX)
        KLAUNCH(lzx_forward_kernel<true>, L.nBlocks, 32, s, L, W);
    else if (type == T_LZ)
        KLAUNCH(lzx_forward_kernel<false>, L.nBlocks, 32, s, L, W);
    else
        KLAUNCH(lzp_forward_kernel, L.nBlocks, 32, s, L, W);
    *launches += 1;
}

void launch_lz_inverse(const StageLaunch& L, int type, LzWork& W, cudaStream_t s, u64* launches)
{
    if (type == T_LZP) {
        cudaMemsetAsync(W.hashes + (i64)L.wsBlock0 * W.hashStride, 0, sizeof(int) * (size_t)W.hashStride * (size_t)L.nBlocks,
                        s);
        KLAUNCH(lzp_inverse_kernel, L.nBlocks, 32, s, L, W);
    } else {
        KLAUNCH(lzx_inverse_kernel, L.nBlocks, 32, s, L);
    }
    *launches += 1;
}

bool lz_work_alloc(LzWork& W, int maxBlocks, i64 stageStride)
{
    memset(&W, 0, sizeof(W));
    W.hashStride = (i64)1 << LZX_HASH_LOG;
    W.sideStride = stageStride;
    bool ok = cudaMalloc((void**)&W.hashes, sizeof(int) * (size_t)W.hashStride * (size_t)maxBlocks) == cudaSuccess;
    ok = ok && cudaMalloc((void**)&W.side, (size_t)(2 * W.sideStride) * (size_t)maxBlocks) == cudaSuccess;
    if (!ok)
        lz_work_free(W);
    return ok;
}

void lz_work_free(LzWork& W)
{
    if (W.hashes)
        cudaFree(W.hashes);
    if (W.side)
        cudaFree(W.side);
    memset(&W, 0, sizeof(W));
}
