// ans_tables.cuh -- per-chunk rANS statistics: frequency normalisation, chunk
// header emission and encoder/decoder table construction.  Sequential code run
// by ONE lane per 16 KiB chunk (8 chunks per warp in flight); also compiled for
// the host (KNZ_HD) so the emulator build (tests/sim) exercises the same small-integer code; the
// CPU suite compares it with the oracle.
//
// Format and arithmetic follow the reference bit for bit:
//   normalisation  entropy/EntropyUtils.cpp:131-245
//   alphabet       entropy/EntropyUtils.cpp:57-89
//   chunk header   entropy/ANSRangeEncoder.cpp:83-155
//   symbol reset   entropy/ANSRangeEncoder.hpp:92-116
// NOTE: the slow error-spreading path of the normaliser (normalize_spread / normalize_counts) restates
// EntropyUtils::normalizeFrequencies decision for decision -- which symbol gives up or receives a unit of
// frequency decides the header bytes and every later state -- so these lines have no freedom of design.
// The fast path is warp-parallel (ans.cu) and original.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define KNZ_HD __host__ __device__ __forceinline__
#else
#define KNZ_HD inline
#endif

namespace knz {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

// MSB-first bit sink over private bytes (no concurrent writers).
struct BitSink {
    u8* p;
    u64 acc;
    int n;     // pending bits in acc (< 8 after put)
    u32 total; // bits written so far

    KNZ_HD void init(u8* dst)
    {
        p = dst;
        acc = 0;
        n = 0;
        total = 0;
    }

    KNZ_HD void put(u32 v, int bits) // bits in [0, 32]
    {
        if (bits == 0)
            return;
        acc = (acc << bits) | (u64)(v & (bits == 32 ? 0xFFFFFFFFu : ((1u << bits) - 1u)));
        n += bits;
        total += (u32)bits;
        while (n >= 8) {
            *p++ = (u8)(acc >> (n - 8));
            n -= 8;
        }
    }

    KNZ_HD void finish() // zero-pad the last partial byte (total is unchanged)
    {
        if (n > 0)
            *p = (u8)((acc << (8 - n)) & 0xFF);
    }
};

KNZ_HD int log2_floor(u32 x) // x >= 1
{
    int r = 0;
    while (x > 1) {
        x >>= 1;
        r++;
    }
    return r;
}

// Scale the 256 counts in f[] (sum = total > 0) to sum `scale`; returns the
// alphabet size.  f[] is updated in place.  `Arr` is anything indexable by symbol
// (a plain u32* or the lane-rotated view the encoder keeps in shared memory).
template <class Arr>
KNZ_HD int normalize_counts(Arr f, u32 total, u32 scale)
{
    int asz = 0;
    if (total == scale) {
        for (int i = 0; i < 256; i++)
            asz += (f[i] != 0) ? 1 : 0;
        return asz;
    }
    u32 sumScaled = 0, sumFreq = 0;
    int idxMax = 0, first = -1;
    for (int i = 0; i < 256; i++) {
        const u32 c = f[i];
        if (c == 0)
            continue;
        if (first < 0)
            first = i;
        asz++;
        const u64 sf = (u64)c * scale;
        const u32 sc = (sf <= (u64)total) ? 1u : (u32)((sf + (total >> 1)) / total);
        sumScaled += sc;
        f[i] = sc;
        sumFreq += c;
        if (sc > f[idxMax])
            idxMax = i;
        if (sumFreq >= total)
            break;
    }
    if (asz == 0)
        return 0;
    if (asz == 1) {
        f[first] = scale;
        return 1;
    }
    if (sumScaled == scale)
        return asz;
    int delta = (int)(sumScaled - scale);
    const int errThr = (int)f[idxMax] >> 4;
    const int ad = (delta < 0) ? -delta : delta;
    if (ad <= errThr) {
        f[idxMax] -= (u32)delta;
        return asz;
    }
    if (delta < 0) {
        delta += errThr;
        f[idxMax] += (u32)errThr;
    } else {
        delta -= errThr;
        f[idxMax] -= (u32)errThr;
    }
    const int inc = (delta < 0) ? 1 : -1;
    delta = (delta < 0) ? -delta : delta;
    int round = 0;
    while ((++round < 6) && (delta > 0)) {
        int adjustments = 0;
        for (int i = 0; i < 256; i++) {
            if (f[i] <= 2) // absent symbols (0) are skipped by the same test
                continue;
            f[i] += (u32)inc;
            adjustments++;
            delta--;
            if (delta == 0)
                break;
        }
        if (adjustments == 0)
            break;
    }
    const u32 v = f[idxMax] - (u32)delta;
    f[idxMax] = (v > 1u) ? v : 1u;
    return asz;
}

// Emit the order-0 chunk header for normalised freqs f[] (alphabet size asz):
// logRange-8 (3 bits), alphabet, then freq groups.
template <class Arr>
KNZ_HD void put_chunk_header(BitSink& w, Arr f, int asz, int lr)
{
    w.put((u32)(lr - 8), 3);
    if (asz == 256) {
        w.put(0, 2);
    } else if (asz == 0) {
        w.put(1, 2);
    } else {
        int last = 0;
        for (int i = 255; i >= 0; i--)
            if (f[i] != 0) {
                last = i >> 3;
                break;
            }
        w.put(1, 1);
        w.put((u32)last, 5);
        for (int b = 0; b <= last; b++) {
            u32 m = 0;
            for (int j = 0; j < 8; j++)
                if (f[8 * b + j] != 0)
                    m |= 1u << j;
            w.put(m, 8);
        }
    }
    if (asz <= 1)
        return;
    const int chk = (asz >= 64) ? 8 : 6;
    const int llr = log2_floor((u32)lr) + 1;
    // walk present symbols in increasing order; the first one is implicit
    int sym = 0;
    while (f[sym] == 0)
        sym++;
    sym++; // skip alphabet[0]
    int left = asz - 1;
    while (left > 0) {
        const int cnt = (left < chk) ? left : chk;
        // pass 1: max over the next cnt present symbols
        u32 mx = 0;
        int s = sym, k = 0;
        while (k < cnt) {
            if (f[s] != 0) {
                if (f[s] - 1 > mx)
                    mx = f[s] - 1;
                k++;
            }
            s++;
        }
        const int logMax = (mx == 0) ? 0 : log2_floor(mx) + 1;
        w.put((u32)logMax, llr);
        // pass 2: values
        k = 0;
        while (k < cnt) {
            if (f[sym] != 0) {
                if (logMax != 0)
                    w.put(f[sym] - 1, logMax);
                k++;
            }
            sym++;
        }
        left -= cnt;
    }
}

// Packed encoder entry (8 bytes): lo = invFreq (32 bits);
// hi = freq[0:12] | bias[12:25] | (invShift-32)[25:29].   freq is clamped to
// 2^lr - 1 as in ANSEncSymbol::reset; cmplFreq = 2^lr - freq is recomputed.
KNZ_HD u64 make_enc_entry(int cum, int freq, int lr)
{
    if (freq >= (1 << lr))
        freq = (1 << lr) - 1;
    u32 invFreq, sh, bias;
    if (freq < 2) {
        invFreq = 0xFFFFFFFFu;
        sh = 0;
        bias = (u32)(cum + (1 << lr) - 1);
    } else {
        int shift = 0;
        while (freq > (1 << shift))
            shift++;
        invFreq = (u32)(((((u64)1 << (shift + 31)) + (u64)freq - 1) / (u64)freq) & 0xFFFFFFFFull);
        sh = (u32)(shift - 1);
        bias = (u32)cum;
    }
    const u32 hi = (u32)freq | (bias << 12) | (sh << 25);
    return ((u64)hi << 32) | invFreq;
}

// One rANS encoder step on a packed entry: returns the new state; *emit gets
// the 16-bit renormalisation word when the function returns true in *did.
KNZ_HD u32 enc_step(u32 st, u64 e, int lr, bool* did, u32* word)
{
    const u32 hi = (u32)(e >> 32), invFreq = (u32)e;
    const u32 freq = hi & 0xFFF, bias = (hi >> 12) & 0x1FFF, sh = hi >> 25;
    const u32 xMax = freq << (31 - lr); // ((ANS_TOP >> lr) << 16) * freq, ANS_TOP = 2^15
    const bool em = st >= xMax;
    *did = em;
    *word = st & 0xFFFF;
    if (em)
        st >>= 16;
#ifdef __CUDA_ARCH__
    const u32 q = __umulhi(st, invFreq) >> sh;
#else
    const u32 q = (u32)(((u64)st * invFreq) >> 32) >> sh;
#endif
    return st + bias + q * ((1u << lr) - freq);
}

} // namespace knz
