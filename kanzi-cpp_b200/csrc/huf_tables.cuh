// huf_tables.cuh -- per-chunk canonical Huffman code construction (sequential part,
// one lane per 16 KiB chunk) and chunk-header emission.  Bit-exact with
//   entropy/HuffmanEncoder.cpp:58-126  updateFrequencies
//   entropy/HuffmanEncoder.cpp:129-215 limitCodeLengths
//   entropy/HuffmanEncoder.cpp:219-300 computeCodeLengths (Moffat-Katajainen, in place)
//   entropy/HuffmanCommon.cpp:29-63    generateCanonicalCodes
//   entropy/ExpGolombEncoder.hpp:51-62 signed exp-Golomb of the length deltas
// The caller provides the (freq << 8 | symbol) keys already SORTED increasingly
// (the warp sorts them in parallel); everything else is small serial integer work.
// NOTE: huf_inplace_lengths and huf_limit_fast necessarily restate the reference's serial heuristics
// (computeInPlaceSizesPhase1/2, limitCodeLengths) decision for decision: every tie-break and every unit of
// "debt" moved changes code lengths and therefore the bitstream, so there is no freedom of design in these
// ~150 lines; the parallel work around them (histogram, sort, emit) is this repo's own.
#pragma once
#include "ans_tables.cuh"

namespace knz {

#define HUF_MAX_LEN 12
#define HUF_CHUNK 16384

// In-place minimum-redundancy code lengths on freqs d[0..n) sorted increasingly.
// Returns the maximum length; d[i] becomes the length of the i-th smallest symbol.
KNZ_HD int huf_inplace_lengths(u32* d, int n)
{
    if (n < 2)
        return 0;
    for (int s = 0, r = 0, t = 0; t < n - 1; t++) {
        u32 sum = 0;
        for (int i = 0; i < 2; i++) {
            if (s >= n || (r < t && d[r] < d[s])) {
                sum += d[r];
                d[r] = (u32)t;
                r++;
                continue;
            }
            sum += d[s];
            if (s > t)
                d[s] = 0;
            s++;
        }
        d[t] = sum;
    }
    u32 topLevel = (u32)n - 2, depth = 1, totalNodes = 2;
    int m = n;
    while (m > 0) {
        u32 k = topLevel;
        while (k != 0 && d[k - 1] >= topLevel)
            k--;
        const int internal = (int)(topLevel - k);
        const int leaves = (int)totalNodes - internal;
        for (int j = 0; j < leaves; j++)
            d[--m] = depth;
        totalNodes = (u32)internal << 1;
        topLevel = k;
        depth++;
    }
    return (int)depth - 1;
}

// Fold lengths above 12 bits back under the limit ("bit debt" repayment).
// order[] = symbols by increasing frequency, sizes[] by symbol.  Returns the new
// maximum length, or -1 when the fast repayment fails (caller takes the slow path).
KNZ_HD int huf_limit_fast(u8* sizes, const u8* order, int count, u8* scratch /* >= 6*256 */)
{
    int n = 0, debt = 0;
    while (n < count && sizes[order[n]] >= HUF_MAX_LEN) {
        debt += sizes[order[n]] - HUF_MAX_LEN;
        sizes[order[n]] = HUF_MAX_LEN;
        n++;
    }
    if (debt == 0)
        return HUF_MAX_LEN;
    int vn[6] = { 0, 0, 0, 0, 0, 0 }, vh[6] = { 0, 0, 0, 0, 0, 0 };
    while (n < count) {
        const int idx = HUF_MAX_LEN - 1 - sizes[order[n]];
        if (idx > 5 || debt < (1 << idx))
            break;
        scratch[idx * 256 + vn[idx]++] = (u8)n;
        n++;
    }
    int idx = 5;
    while (debt > 0 && idx >= 0) {
        if (vh[idx] >= vn[idx] || debt < (1 << idx)) {
            idx--;
            continue;
        }
        sizes[order[scratch[idx * 256 + vh[idx]]]]++;
        debt -= 1 << idx;
        vh[idx]++;
    }
    idx = 0;
    while (debt > 0 && idx < 6) {
        if (vh[idx] >= vn[idx]) {
            idx++;
            continue;
        }
        sizes[order[scratch[idx * 256 + vh[idx]]]]++;
        debt -= 1 << idx;
        vh[idx]++;
    }
    return (debt > 0) ? -1 : HUF_MAX_LEN;
}

// Canonical codes: symbols ordered by (length, symbol) receive consecutive codes.
// codes[s] = (length << 12) | code.
KNZ_HD void huf_canonical_codes(const u8* sizes, u16* codes)
{
    int code = 0, curLen = 0;
    bool first = true;
    for (int len = 1; len <= HUF_MAX_LEN; len++)
        for (int s = 0; s < 256; s++) {
            if (sizes[s] != len)
                continue;
            if (first) {
                curLen = len;
                first = false;
            }
            code <<= (len - curLen);
            curLen = len;
            codes[s] = (u16)((len << 12) | code);
            code++;
        }
}

KNZ_HD void put_expgolomb_signed(BitSink& w, int v)
{
    if (v == 0) {
        w.put(1, 1);
        return;
    }
    const u32 x = (u32)(v < 0 ? -v : v) + 1;
    const int lg = log2_floor(x);
    w.put(0, lg);
    w.put(x, lg + 1);
    w.put(v < 0 ? 1u : 0u, 1);
}

// Alphabet + length deltas (HuffmanEncoder.cpp:69, :112-122).  present[s] != 0 marks the alphabet.
KNZ_HD void huf_put_header(BitSink& w, const u8* sizes, const u32* freqs, int count)
{
    if (count == 256) {
        w.put(0, 2);
    } else if (count == 0) {
        w.put(1, 2);
    } else {
        int last = 0;
        for (int i = 255; i >= 0; i--)
            if (freqs[i] != 0) {
                last = i >> 3;
                break;
            }
        w.put(1, 1);
        w.put((u32)last, 5);
        for (int b = 0; b <= last; b++) {
            u32 m = 0;
            for (int j = 0; j < 8; j++)
                if (freqs[8 * b + j] != 0)
                    m |= 1u << j;
            w.put(m, 8);
        }
    }
    int prev = 2;
    for (int s = 0; s < 256; s++) {
        if (freqs[s] == 0)
            continue;
        put_expgolomb_signed(w, (int)(int8_t)(sizes[s] - prev));
        prev = sizes[s];
    }
}

} // namespace knz
