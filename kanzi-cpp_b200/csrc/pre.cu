// pre.cu -- HOST stages of the block pipeline: the data-type sniffing pre-transforms that the
// reference's default levels put in front of the device stages.  They are serial per block, mostly refuse
// (`false` = stage skipped) and hand a `dataType` on to the stages behind them through the reference's Context
// (io/CompressedOutputStream.cpp:722-731); the stream-level entry points run them on host threads, one block
// per task, before a batch goes to the device (encode) and after it came back (decode).
//
//   PACK / DNA   AliasCodec   transform/AliasCodec.cpp:37-229 (forward), :232-425 (inverse)
//   MM           FSDCodec     transform/FSDCodec.cpp:103-277 (forward), :279-386 (inverse)
//   UTF          UTFCodec     transform/UTFCodec.cpp:49-226 (forward), :228-326 (inverse), validate :331-422
//   TEXT         TextCodec    pretext.cu
//
// Byte-exact with the reference (tests/test_pre_stages.py pins every stage and whole streams against
// oracle/_ref).  This file holds host code only; it is compiled with the rest of the library.
#include <algorithm>
#include <string.h>
#include <vector>

#include "pre.h"

namespace {

// ---- shared statistics ---------------------------------------------------------------------
void histogram0(const u8* p, int n, u32 f[256])
{
    memset(f, 0, 256 * sizeof(u32));
    for (int i = 0; i < n; i++)
        f[p[i]]++;
}

// Pairs (previous byte, byte) with the first byte of the block counted behind a zero byte: what
// Global::computeHistogram(order 1, no totals) yields once its four quarter streams are summed
// (Global.cpp:276-308: the quarters start from their true predecessors, the block from 0).
void histogram1(const u8* p, int n, u32* f /*65536*/)
{
    memset(f, 0, 65536 * sizeof(u32));
    u32 prev = 0;
    for (int i = 0; i < n; i++) {
        f[(prev << 8) | p[i]]++;
        prev = p[i];
    }
}

int log2_1024(u32 x, const int* tab)
{
    if (x < 256)
        return (tab[x] + 2) >> 2;
    const int lg = 31 - __builtin_clz(x);
    if ((x & (x - 1)) == 0)
        return lg << 10;
    return ((lg - 7) << 10) + ((tab[x >> (lg - 7)] + 2) >> 2);
}

// Global::computeFirstOrderEntropy1024 (Global.cpp:313-329)
int entropy1024(int n, const u32 f[256], const int* tab)
{
    if (n == 0)
        return 0;
    u64 sum = 0;
    const int ln = log2_1024((u32)n, tab);
    for (int i = 0; i < 256; i++)
        if (f[i])
            sum += ((u64)f[i] * (u64)(i64)(ln - log2_1024(f[i], tab))) >> 3;
    return (int)(sum / (u64)n);
}

const int* log2_table()
{
    static int tab[257];
    static bool ready = false;
    if (!ready) {
        knz_log2_table(tab);
        ready = true;
    }
    return tab;
}

// Magic.hpp:69-166 (getType, isCompressed, isMultimedia, isExecutable) as one classification
enum { MG_NONE = 0, MG_BMP, MG_RIFF, MG_PNM, MG_EXE, MG_COMPRESSED, MG_OTHER };
int magic_class(const u8* p)
{
    const u32 key = ((u32)p[0] << 24) | ((u32)p[1] << 16) | ((u32)p[2] << 8) | p[3];
    if ((key & ~0x0Fu) == 0xFFD8FFE0u) // JPG
        return (key == 0xFFD8FFE0u) ? MG_COMPRESSED : MG_OTHER; // isCompressed compares the full key
    const u32 k24 = key >> 8;
    if (k24 == 0x425A68u || k24 == 0x494433u) // BZIP2, MP3 / ID3
        return MG_COMPRESSED;
    switch (key) {
    case 0x47494638u: case 0x504B0304u: case 0x377ABCAFu: case 0x89504E47u: case 0x28B52FFDu: case 0x81CFB2CEu:
    case 0x4D534346u: case 0x664C6143u: case 0xFD377A58u: case 0x4B414E5Au: case 0x52617221u:
        return MG_COMPRESSED; // GIF ZIP 7z PNG ZSTD BROTLI CAB FLAC XZ KNZ RAR
    case 0x7F454C46u: case 0xFEEDFACEu: case 0xCEFAEDFEu: case 0xFEEDFACFu: case 0xCFFAEDFEu:
        return MG_EXE; // ELF, Mach-O
    case 0x25504446u:
        return MG_OTHER; // PDF: known, in no class
    case 0x52494646u:
        return MG_RIFF;
    default:
        break;
    }
    const u32 k16 = key >> 16;
    if (k16 == 0x1F8Bu)
        return MG_COMPRESSED; // GZIP
    if (k16 == 0x424Du)
        return MG_BMP;
    if (k16 == 0x4D5Au)
        return MG_EXE; // MZ
    if (k16 == 0x5034u || k16 == 0x5035u || k16 == 0x5036u) {
        const u32 sub = (key >> 8) & 0xFF;
        if (sub == 0x07 || sub == 0x0A || sub == 0x0D || sub == 0x20)
            return MG_PNM;
    }
    return MG_NONE;
}

struct Ranked { // aliases are handed out by decreasing frequency, ties by decreasing value
    u32 val, freq;
    bool operator<(const Ranked& o) const { return (freq != o.freq) ? freq > o.freq : val > o.val; }
};

// ---- PACK / DNA ----------------------------------------------------------------------------
bool alias_forward(const u8* src, int n, u8* dst, int cap, int* outLen, KnzPreCtx* pc, bool onlyDna)
{
    if (n < 1024 || cap < n + 1024)
        return false;
    int dt = pc->dataType;
    if (dt == KDT_MULTIMEDIA || dt == KDT_UTF8 || dt == KDT_EXE || dt == KDT_BIN)
        return false;
    if (onlyDna && dt != KDT_UNDEFINED && dt != KDT_DNA)
        return false;
    u32 f0[256];
    histogram0(src, n, f0);
    int freeSym[256], nFree = 0;
    for (int i = 0; i < 256; i++)
        if (f0[i] == 0)
            freeSym[nFree++] = i;
    if (nFree < 16)
        return false;
    if (dt == KDT_UNDEFINED) {
        dt = knz_detect_simple_type(n, f0);
        if (dt != KDT_UNDEFINED)
            pc->dataType = dt;
        if (dt != KDT_DNA && onlyDna)
            return false;
    }
    int s = 0, d = 0;
    if (nFree >= 240) { // at most 16 symbols: 2 or 4 bits each
        dst[d++] = (u8)nFree;
        if (nFree == 255) { // a single symbol: value + length
            dst[1] = src[0];
            dst[2] = (u8)n, dst[3] = (u8)(n >> 8), dst[4] = (u8)(n >> 16), dst[5] = (u8)(n >> 24);
            d = 6;
            s = n;
        } else {
            u8 code[256];
            memset(code, 0, sizeof(code));
            for (int i = 0, j = 0; i < 256; i++)
                if (f0[i]) {
                    dst[d++] = (u8)i;
                    code[i] = (u8)j++;
                }
            if (nFree >= 252) {
                const int head = n & 3;
                dst[d++] = (u8)head;
                for (; s < head; s++)
                    dst[d++] = src[s];
                for (; s < n; s += 4)
                    dst[d++] = (u8)((code[src[s]] << 6) | (code[src[s + 1]] << 4) | (code[src[s + 2]] << 2) | code[src[s + 3]]);
            } else {
                dst[d++] = (u8)(n & 1);
                if (n & 1)
                    dst[d++] = src[s++];
                for (; s < n; s += 2)
                    dst[d++] = (u8)((code[src[s]] << 4) | code[src[s + 1]]);
            }
        }
    } else { // unused byte values stand for the most frequent byte pairs
        std::vector<Ranked> pairs;
        {
            std::vector<u32> f1(65536);
            histogram1(src, n, f1.data());
            for (u32 i = 0; i < 65536; i++)
                if (f1[i])
                    pairs.push_back(Ranked{ i, f1[i] });
        }
        if ((int)pairs.size() < nFree) {
            nFree = (int)pairs.size();
            if (nFree < 16)
                return false;
        }
        std::sort(pairs.begin(), pairs.end());
        std::vector<u16> emit(65536); // low byte: what to write, high byte: source bytes consumed
        for (u32 i = 0; i < 65536; i++)
            emit[i] = (u16)(0x100 | (i >> 8));
        dst[0] = (u8)nFree;
        dst[1] = 0;
        d = 2;
        i64 gain = 0;
        for (int i = 0; i < nFree; i++) {
            gain += pairs[i].freq;
            const u32 v = pairs[i].val;
            emit[v] = (u16)(0x200 | freeSym[i]);
            dst[d++] = (u8)(v >> 8);
            dst[d++] = (u8)v;
            dst[d++] = (u8)freeSym[i];
        }
        if (gain < n / 20)
            return false;
        while (s < n - 1) {
            const u16 a = emit[((u32)src[s] << 8) | src[s + 1]];
            dst[d++] = (u8)a;
            s += a >> 8;
        }
        if (s != n) {
            dst[1] = 1;
            dst[d++] = src[s++];
        }
    }
    *outLen = d;
    return d < n;
}

bool alias_inverse(const u8* src, int n, u8* dst, int cap, int* outLen)
{
    int k = src[0];
    if (k < 16)
        return false;
    int s, d = 0;
    if (k >= 240) {
        k = 256 - k;
        s = 1;
        if (k == 1) {
            if (n < 6)
                return false;
            const int len = (int)((u32)src[2] | ((u32)src[3] << 8) | ((u32)src[4] << 16) | ((u32)src[5] << 24));
            if (len < 0 || len > cap)
                return false;
            memset(dst, src[1], (size_t)len);
            s = n;
            d = len;
        } else {
            u8 sym[16];
            memset(sym, 0, sizeof(sym));
            if (s + k + 1 > n)
                return false;
            for (int i = 0; i < k; i++)
                sym[i] = src[s++];
            const int head = src[s++];
            if (head >= 4)
                return false;
            if (k <= 4) {
                if (s + head > n || d + head > cap)
                    return false;
                for (int i = 0; i < head; i++)
                    dst[d++] = src[s++];
                if (n - s > ((cap - d) >> 2))
                    return false;
                for (; s < n; s++, d += 4) {
                    const int b = src[s];
                    dst[d] = sym[(b >> 6) & 3], dst[d + 1] = sym[(b >> 4) & 3], dst[d + 2] = sym[(b >> 2) & 3], dst[d + 3] = sym[b & 3];
                }
            } else {
                if (head != 0) {
                    if (s >= n || d >= cap)
                        return false;
                    dst[d++] = src[s++];
                }
                if (n - s > ((cap - d) >> 1))
                    return false;
                for (; s < n; s++, d += 2) {
                    const int b = src[s];
                    dst[d] = sym[b >> 4], dst[d + 1] = sym[b & 15];
                }
            }
        }
    } else {
        if (n < 2)
            return false;
        const int tail = src[1];
        if (tail > 1)
            return false;
        const int end = n - tail;
        s = 2;
        u32 expand[256]; // bits 0-15: one or two bytes (first in the low byte), bits 16-17: how many
        for (int i = 0; i < 256; i++)
            expand[i] = 0x10000u | (u32)i;
        if (s + 3 * k > end)
            return false;
        for (int i = 0; i < k; i++, s += 3)
            expand[src[s + 2]] = 0x20000u | src[s] | ((u32)src[s + 1] << 8);
        // 256 input bytes at a time while even 256 digrams fit the destination: two bytes stored per input
        // byte without a check (the second one is overwritten by the next store when the entry is a single byte)
        while (end - s >= 256 && d + 512 <= cap) {
            for (const int stop = s + 256; s < stop; s++) {
                const u32 e = expand[src[s]];
                dst[d] = (u8)e;
                dst[d + 1] = (u8)(e >> 8);
                d += (int)(e >> 16);
            }
        }
        for (; s < end; s++) {
            const u32 e = expand[src[s]];
            const int len = (int)(e >> 16);
            if (d + len > cap)
                return false;
            dst[d] = (u8)e;
            if (len == 2)
                dst[d + 1] = (u8)(e >> 8);
            d += len;
        }
        if (tail) {
            if (s >= n || d >= cap)
                return false;
            dst[d++] = src[s++];
        }
    }
    *outLen = d;
    return s == n;
}

// ---- MM (fixed step delta) -----------------------------------------------------------------
bool fsd_forward(const u8* src, int n, u8* dst, int cap, int* outLen, KnzPreCtx* pc)
{
    const int maxOut = knz_pre_max_len(KNZ_T_MM, n);
    if (cap < maxOut || n < 1024)
        return false;
    if (pc->dataType != KDT_UNDEFINED && pc->dataType != KDT_MULTIMEDIA && pc->dataType != KDT_BIN)
        return false;
    {
        const int mg = magic_class(src); // detection only runs on image / audio containers and on unknown data
        if (mg != MG_NONE && mg != MG_BMP && mg != MG_RIFF && mg != MG_PNM)
            return false;
    }
    const int* tab = log2_table();
    const int tenth = n / 10, fifth = 2 * tenth;
    static const int steps[7] = { 0, 1, 2, 3, 4, 8, 16 };
    // three sample windows: [tenth, fifth) of the sections that start at 0, 2/5 and 4/5 of the block
    std::vector<u32> hist(7 * 256, 0u);
    for (int w = 0; w < 3; w++) {
        const u8* q = src + fifth * 2 * w;
        for (int i = tenth; i < fifth; i++) {
            const u8 b = q[i];
            hist[b]++;
            for (int k = 1; k < 7; k++)
                hist[256 * k + (b ^ q[i - steps[k]])]++;
        }
    }
    int ent[7], best = 0;
    for (int k = 0; k < 7; k++) {
        ent[k] = entropy1024(3 * tenth, &hist[256 * k], tab);
        if (ent[k] < ent[best])
            best = k;
    }
    if (ent[best] >= ent[0]) {
        pc->dataType = knz_detect_simple_type(3 * tenth, &hist[0]);
        return false;
    }
    pc->dataType = KDT_MULTIMEDIA;
    const int step = steps[best];
    int big = 0;
    for (int i = 2 * fifth; i < 3 * fifth; i++) {
        const int delta = (int)src[i] - (int)src[i - step];
        big += (delta < -127 || delta > 127) ? 1 : 0;
    }
    const bool useXor = big > (fifth >> 5); // deltas for pictures, xor for audio
    dst[0] = useXor ? 1 : 0;
    dst[1] = (u8)step;
    int s = 0, d = 2;
    for (; s < step; s++)
        dst[d++] = src[s];
    if (!useXor) {
        while (s < n && d < maxOut - 1) {
            const int delta = (int)src[s] - (int)src[s - step];
            if (delta >= -127 && delta <= 127) // zigzag: 0, -1, 1, -2, 2, ... on 0..254; 255 escapes
                dst[d++] = (u8)((delta < 0) ? (-2 * delta - 1) : 2 * delta);
            else {
                dst[d++] = 0xFF;
                dst[d++] = (u8)(src[s] ^ src[s - step]);
            }
            s++;
        }
    } else {
        for (; s < n; s++)
            dst[d++] = (u8)(src[s] ^ src[s - step]);
    }
    if (s != n)
        return false;
    // the output must look better than the input did
    u32 h[256];
    memset(h, 0, sizeof(h));
    for (int i = 0; i < tenth; i++) {
        h[dst[fifth + i]]++;
        h[dst[3 * fifth + i]]++;
    }
    if (entropy1024(fifth, h, tab) >= ent[0])
        return false;
    *outLen = d;
    return true;
}

bool fsd_inverse(const u8* src, int n, u8* dst, int cap, int* outLen)
{
    if (n < 4)
        return false;
    const int mode = src[0], step = src[1];
    if (step < 1 || (step > 4 && step != 8 && step != 16))
        return false;
    if (n < step + 2 || step > cap || mode > 1)
        return false;
    memcpy(dst, src + 2, (size_t)step);
    int s = step + 2, d = step;
    if (mode == 0) {
        while (s < n && d < cap) {
            const int v = src[s++];
            if (v != 0xFF) {
                dst[d] = (u8)(dst[d - step] + ((v >> 1) ^ -(v & 1)));
            } else {
                if (s == n)
                    return false;
                dst[d] = (u8)(src[s++] ^ dst[d - step]);
            }
            d++;
        }
    } else {
        for (; s < n && d < cap; s++, d++)
            dst[d] = (u8)(src[s] ^ dst[d - step]);
    }
    *outLen = d;
    return s == n;
}

// ---- UTF -----------------------------------------------------------------------------------
// sequence length by lead byte (0: continuation byte or not a lead byte)
inline int utf_len(u8 b)
{
    if (b < 0x80)
        return 1;
    if (b >= 0xC2 && b <= 0xDF)
        return 2;
    if (b >= 0xE0 && b <= 0xEF)
        return 3;
    if (b >= 0xF0 && b <= 0xF4)
        return 4;
    return 0;
}

// code point packed on 22 bits: 3-bit length class + payload (UTFCodec.hpp:71-110)
inline int utf_pack(const u8* p, u32* out)
{
    switch (p[0] >> 4) {
    case 0: case 1: case 2: case 3: case 4: case 5: case 6: case 7:
        *out = p[0];
        return 1;
    case 12: case 13:
        *out = (1u << 19) | ((u32)p[0] << 8) | p[1];
        return 2;
    case 14:
        *out = (2u << 19) | (((u32)p[0] & 0x0F) << 12) | (((u32)p[1] & 0x3F) << 6) | (p[2] & 0x3Fu);
        return 3;
    case 15:
        *out = (4u << 19) | (((u32)p[0] & 0x07) << 18) | (((u32)p[1] & 0x3F) << 12) | (((u32)p[2] & 0x3F) << 6) | (p[3] & 0x3Fu);
        return 4;
    default:
        *out = 0;
        return 0;
    }
}

inline int utf_unpack(u32 v, u8 out[4])
{
    switch (v >> 19) {
    case 0:
        out[0] = (u8)v;
        return 1;
    case 1:
        out[0] = (u8)(v >> 8), out[1] = (u8)v;
        return 2;
    case 2:
        out[0] = (u8)(((v >> 12) & 0x0F) | 0xE0), out[1] = (u8)(((v >> 6) & 0x3F) | 0x80), out[2] = (u8)((v & 0x3F) | 0x80);
        return 3;
    case 4: case 5: case 6: case 7:
        out[0] = (u8)(((v >> 18) & 0x07) | 0xF0), out[1] = (u8)(((v >> 12) & 0x3F) | 0x80);
        out[2] = (u8)(((v >> 6) & 0x3F) | 0x80), out[3] = (u8)((v & 0x3F) | 0x80);
        return 4;
    default:
        return 0;
    }
}

} // namespace

// Well-formedness by byte and byte-pair statistics (Unicode table 3-7): no lead byte that cannot occur, every
// lead byte followed by a byte of ITS range, and continuation bytes making up at least an eighth of the data.
bool knz_utf8_plausible(const u32 f0[256], const u32* f1, int n)
{
    u32 bad = f0[0xC0] + f0[0xC1];
    for (int i = 0xF5; i <= 0xFF; i++)
        bad += f0[i];
    if (bad)
        return false;
    for (int lead = 0xC2; lead <= 0xF4; lead++) {
        int lo = 0x80, hi = 0xBF;
        if (lead == 0xE0)
            lo = 0xA0;
        else if (lead == 0xED)
            hi = 0x9F;
        else if (lead == 0xF0)
            lo = 0x90;
        else if (lead == 0xF4)
            hi = 0x8F;
        const u32* row = f1 + 256 * lead;
        for (int i = 0; i < 256; i++)
            if ((i < lo || i > hi) && row[i])
                return false;
    }
    u32 cont = 0;
    for (int i = 0x80; i <= 0xBF; i++)
        cont += f0[i];
    return cont >= (u32)(n / 8);
}

namespace {

bool utf_forward(const u8* src, int n, u8* dst, int cap, int* outLen, KnzPreCtx* pc)
{
    if (n < 1024 || cap < n + 8192)
        return false;
    if (pc->dataType != KDT_UNDEFINED && pc->dataType != KDT_UTF8)
        return false;
    const bool check = pc->dataType != KDT_UTF8;
    int start = 0;
    if (src[0] == 0xEF && src[1] == 0xBB && src[2] == 0xBF)
        start = 3; // byte order mark
    else
        while (start < 4 && utf_len(src[start]) == 0)
            start++; // a sequence cut by the block boundary
    if (check) {
        const u8* p = src + start;
        const int m = n - start - 4;
        u32 f0[256];
        std::vector<u32> f1(65536);
        histogram0(p, m, f0);
        histogram1(p, m, f1.data());
        if (!knz_utf8_plausible(f0, f1.data(), m))
            return false;
    }
    pc->dataType = KDT_UTF8;
    std::vector<u32> seen(1u << 22, 0u); // occurrences per packed code point, then its alias
    std::vector<Ranked> syms;
    bool ok = true;
    for (int i = start; i < n - 4;) {
        u32 v;
        const int len = utf_pack(src + i, &v);
        ok = len != 0;
        if (len == 3)
            ok = ok && (src[i + 2] & 0xC0) == 0x80;
        if (len == 4)
            ok = ok && (src[i + 2] & 0xC0) == 0x80 && (src[i + 3] & 0xC0) == 0x80;
        if (seen[v] == 0) {
            ok = ok && (int)syms.size() + 1 < 32768;
            syms.push_back(Ranked{ v, 0 });
        }
        if (!ok)
            break;
        seen[v]++;
        i += len;
    }
    const int cnt = (int)syms.size();
    const int target = n - n / 10;
    if (!ok || cnt == 0 || 3 * cnt + 6 >= target)
        return false;
    for (Ranked& r : syms)
        r.freq = seen[r.val];
    std::sort(syms.begin(), syms.end());
    int d = 2;
    dst[d++] = (u8)(cnt >> 8);
    dst[d++] = (u8)cnt;
    i64 estimate = d + 6;
    for (int i = 0; i < cnt; i++) {
        estimate += (i < 128) ? syms[i].freq : 2 * (i64)syms[i].freq;
        const u32 v = syms[i].val;
        // one byte below 128; else low 7 bits + flag, then the upper bits; bits 16.. hold the alias length
        seen[v] = (i < 128) ? (0x10000u | (u32)i) : (0x20000u | 0x80u | (((u32)i << 1) & 0xFF00u) | ((u32)i & 0x7Fu));
        dst[d++] = (u8)(v >> 16), dst[d++] = (u8)(v >> 8), dst[d++] = (u8)v;
    }
    if (estimate >= target)
        return false;
    for (int i = 0; i < start; i++)
        dst[d++] = src[i];
    int s = start;
    while (s < n - 4) {
        u32 v;
        s += utf_pack(src + s, &v);
        const u32 a = seen[v];
        dst[d] = (u8)a;
        dst[d + 1] = (u8)(a >> 8);
        d += (int)(a >> 16);
    }
    dst[0] = (u8)start;
    dst[1] = (u8)(s - (n - 4));
    while (s < n)
        dst[d++] = src[s++];
    *outLen = d;
    return d < target;
}

bool utf_inverse(const u8* src, int n, u8* dst, int cap, int* outLen)
{
    if (n < 4)
        return false;
    const int start = src[0] & 3, adjust = src[1] & 3;
    const int cnt = (src[2] << 8) + src[3];
    if (cnt == 0 || cnt >= 32768 || 3 * cnt > n - 4)
        return false;
    struct Sym {
        u8 b[4];
        u8 len;
    };
    std::vector<Sym> map((size_t)cnt);
    int s = 4;
    for (int i = 0; i < cnt; i++, s += 3) {
        if (s + 3 > n)
            return false;
        memset(map[i].b, 0, 4);
        const int len = utf_unpack(((u32)src[s] << 16) | ((u32)src[s + 1] << 8) | src[s + 2], map[i].b);
        if (len == 0)
            return false;
        map[i].len = (u8)len;
    }
    int d = 0;
    const int end = n - 4 + adjust;
    if (cap - 4 < 0 || end > n || s + start > end || start > cap)
        return false;
    for (int i = 0; i < start; i++)
        dst[d++] = src[s++];
    while (s < end) {
        u32 a = src[s++];
        if (a >= 128) {
            if (s >= n)
                return false;
            a = ((u32)src[s++] << 7) + (a & 0x7F);
        }
        if (a >= (u32)cnt)
            return false;
        const Sym& m = map[a];
        if (d + m.len > cap)
            return false;
        for (int i = 0; i < m.len; i++)
            dst[d + i] = m.b[i];
        d += m.len;
    }
    if (s == end && d < cap - 4 + adjust) {
        if (s + 4 - adjust > n || d + 4 - adjust > cap)
            return false;
        for (int i = 0; i < 4 - adjust; i++)
            dst[d++] = src[s++];
    }
    *outLen = d;
    return s == n;
}

} // namespace

// ---- shared entry points -------------------------------------------------------------------
void knz_log2_table(int tab[257])
{
    // Global::LOG2_4096 (Global.cpp:47-74) is round(4096 * log2(i)); generated here, every entry checked
    // against the reference's literal table by tests/test_stream_features.py
    tab[0] = 0;
    for (int i = 1; i <= 256; i++)
        tab[i] = (int)floor(4096.0 * log2((double)i) + 0.5);
}

// Global::detectSimpleType (Global.cpp:354-397)
int knz_detect_simple_type(int n, const u32 f[256])
{
    i64 sum = 0;
    for (const char* p = "acgntuACGNTU"; *p; p++)
        sum += f[(u8)*p];
    if (sum > n - n / 12)
        return KDT_DNA;
    sum = 0;
    for (const char* p = "0123456789+-*/=,.:; "; *p; p++)
        sum += f[(u8)*p];
    if (sum == n)
        return KDT_NUMERIC;
    sum = (f['='] == 1) ? 1 : 0; // a single padding character ends base64 data
    for (const char* p = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/"; *p; p++)
        sum += f[(u8)*p];
    if (sum == n)
        return KDT_BASE64;
    int distinct = 0;
    for (int i = 0; i < 256; i++)
        distinct += f[i] ? 1 : 0;
    if (distinct == 256)
        return KDT_BIN;
    return (distinct <= 4) ? KDT_SMALL_ALPHABET : KDT_UNDEFINED;
}

// What EncodingTask puts into the Context before the transforms run (io/CompressedOutputStream.cpp:722-731)
int knz_magic_data_type(const u8* block, int n)
{
    if (n < 4)
        return KDT_UNDEFINED;
    switch (magic_class(block)) {
    case MG_COMPRESSED:
        return KDT_BIN;
    case MG_BMP:
    case MG_RIFF:
    case MG_PNM:
        return KDT_MULTIMEDIA;
    case MG_EXE:
        return KDT_EXE;
    default:
        return KDT_UNDEFINED;
    }
}

bool knz_magic_known(const u8* block) { return magic_class(block) != MG_NONE; }

bool knz_is_host_stage(int type)
{
    return type == KNZ_T_TEXT || type == KNZ_T_PACK || type == KNZ_T_DNA || type == KNZ_T_MM || type == KNZ_T_UTF;
}

int knz_pre_max_len(int type, int n)
{
    switch (type) {
    case KNZ_T_PACK:
    case KNZ_T_DNA:
        return n + 1024;
    case KNZ_T_MM:
        return n + ((n < 1024) ? 64 : (n >> 4));
    case KNZ_T_UTF:
        return n + 8192;
    default:
        return n;
    }
}

bool knz_pre_forward(int type, const u8* src, int n, u8* dst, int cap, int* outLen, KnzPreCtx* pc)
{
    if (n == 0) {
        *outLen = 0;
        return true;
    }
    switch (type) {
    case KNZ_T_PACK:
        return alias_forward(src, n, dst, cap, outLen, pc, false);
    case KNZ_T_DNA:
        return alias_forward(src, n, dst, cap, outLen, pc, true);
    case KNZ_T_MM:
        return fsd_forward(src, n, dst, cap, outLen, pc);
    case KNZ_T_UTF:
        return utf_forward(src, n, dst, cap, outLen, pc);
    case KNZ_T_TEXT:
        return knz_text_forward(src, n, dst, cap, outLen, pc);
    default:
        return false;
    }
}

bool knz_pre_inverse(int type, const u8* src, int n, u8* dst, int cap, int* outLen, const KnzPreCtx* pc)
{
    if (n == 0) {
        *outLen = 0;
        return true;
    }
    switch (type) {
    case KNZ_T_PACK:
    case KNZ_T_DNA:
        return alias_inverse(src, n, dst, cap, outLen);
    case KNZ_T_MM:
        return fsd_inverse(src, n, dst, cap, outLen);
    case KNZ_T_UTF:
        return utf_inverse(src, n, dst, cap, outLen);
    case KNZ_T_TEXT:
        return knz_text_inverse(src, n, dst, cap, outLen, pc->blockSize, pc->eType);
    default:
        return false;
    }
}
