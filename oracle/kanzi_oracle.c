/* oracle/kanzi_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked or loaded by the product).
 *
 * Plain-C, single-threaded restatement of kanzi's per-block transform -> entropy
 * path (bitstream v6) written from the reference's behaviour, used as the CPU
 * oracle of the parity tests.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load it.
 *
 * PINNING: this restatement is checked byte-for-byte against the unmodified
 * reference compiled into oracle/_ref/libkanzi_ref.so (tests/test_oracle_vs_ref.py,
 * run in the build container) and against the committed fixtures in
 * tests/golden/ that were generated from that same library (the reference ships
 * no golden vectors of its own: SURVEY.md §8(c)).
 *
 * Each function cites the reference file:line it restates (paths under
 * /root/reference/src).  Style is deliberately straight-line C; nothing here is
 * tuned for speed.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int64_t i64;

/* ------------------------------------------------------------------ bit I/O
 * MSB-first packing.  bitstream/DefaultOutputBitStream.hpp:97-132 (writeBits),
 * .cpp:42-128 (byte-array writes: whole bytes, then the TOP `rem` bits of the
 * last byte), written() exact in bits after close (.hpp:73, .cpp:130-150).   */
typedef struct {
    u8* buf;
    i64 cap;   /* bytes */
    i64 bits;  /* bits written */
    int overflow;
} BitW;

static void bw_init(BitW* w, u8* buf, i64 cap)
{
    w->buf = buf;
    w->cap = cap;
    w->bits = 0;
    w->overflow = 0;
    if (cap > 0)
        memset(buf, 0, (size_t)cap);
}

static void bw_put(BitW* w, u64 v, int n)
{
    for (int k = n - 1; k >= 0; k--) {
        const i64 byte = w->bits >> 3;
        if (byte >= w->cap) {
            w->overflow = 1;
            w->bits++;
            continue;
        }
        if ((v >> k) & 1)
            w->buf[byte] |= (u8)(0x80 >> (w->bits & 7));
        w->bits++;
    }
}

static void bw_put_bytes(BitW* w, const u8* p, i64 nbits)
{
    i64 i = 0;
    for (; nbits >= 8; nbits -= 8)
        bw_put(w, p[i++], 8);
    if (nbits > 0)
        bw_put(w, (u64)(p[i] >> (8 - nbits)), (int)nbits);
}

/* bitstream/DefaultInputBitStream.hpp:88-150 */
typedef struct {
    const u8* buf;
    i64 nbits; /* total bits available */
    i64 pos;
    int underflow;
} BitR;

static void br_init(BitR* r, const u8* buf, i64 nbits)
{
    r->buf = buf;
    r->nbits = nbits;
    r->pos = 0;
    r->underflow = 0;
}

static u64 br_get(BitR* r, int n)
{
    u64 v = 0;
    for (int k = 0; k < n; k++) {
        u64 b = 0;
        if (r->pos < r->nbits)
            b = (r->buf[r->pos >> 3] >> (7 - (r->pos & 7))) & 1;
        else
            r->underflow = 1;
        v = (v << 1) | b;
        r->pos++;
    }
    return v;
}

static void br_get_bytes(BitR* r, u8* dst, i64 nbytes)
{
    for (i64 i = 0; i < nbytes; i++)
        dst[i] = (u8)br_get(r, 8);
}

static int ilog2(u32 x) /* Global.hpp:91 _log2 */
{
    int r = 0;
    while (x > 1) {
        x >>= 1;
        r++;
    }
    return r;
}

/* entropy/EntropyUtils.cpp:247-259 writeVarInt / :261-286 readVarInt */
static void put_varint(BitW* w, u32 v)
{
    while (v >= 128) {
        bw_put(w, 0x80 | (v & 0x7F), 8);
        v >>= 7;
    }
    bw_put(w, v, 8);
}

static int get_varint(BitR* r, u32* out)
{
    u32 v = (u32)br_get(r, 8);
    u32 res = v & 0x7F;
    for (int shift = 7; v >= 128; shift += 7) {
        v = (u32)br_get(r, 8);
        if (shift == 28) {
            if (v >= 128 || (v & 0x70))
                return -1;
            res |= (v & 0x0F) << shift;
            break;
        }
        res |= (v & 0x7F) << shift;
    }
    *out = res;
    return 0;
}

/* entropy/EntropyUtils.cpp:57-89 encodeAlphabet: "00" = all 256, "01" = empty,
 * else "1", 5-bit index of the last non-zero mask byte, then the mask bytes
 * (bit j of byte i = symbol 8i+j present).                                  */
static void put_alphabet(BitW* w, const u32* alphabet, int count)
{
    if (count == 0) {
        bw_put(w, 0, 1);
        bw_put(w, 1, 1);
    } else if (count == 256) {
        bw_put(w, 0, 1);
        bw_put(w, 0, 1);
    } else {
        u8 masks[32] = { 0 };
        bw_put(w, 1, 1);
        for (int i = 0; i < count; i++)
            masks[alphabet[i] >> 3] |= (u8)(1 << (alphabet[i] & 7));
        const int last = (int)(alphabet[count - 1] >> 3);
        bw_put(w, (u64)last, 5);
        bw_put_bytes(w, masks, 8 * (last + 1));
    }
}

/* entropy/EntropyUtils.cpp:91-123 decodeAlphabet */
static int get_alphabet(BitR* r, u32* alphabet)
{
    if (br_get(r, 1) == 0) {
        const int n = (br_get(r, 1) == 0) ? 256 : 0;
        for (int i = 0; i < n; i++)
            alphabet[i] = (u32)i;
        return n;
    }
    const int last = (int)br_get(r, 5);
    int count = 0;
    for (int i = 0; i <= last; i++) {
        const u32 m = (u32)br_get(r, 8);
        for (int j = 0; j < 8; j++)
            if ((m >> j) & 1)
                alphabet[count++] = (u32)(8 * i + j);
    }
    return count;
}

/* entropy/EntropyUtils.cpp:131-245 normalizeFrequencies (length fixed to 256).
 * All arithmetic on freqs is uint32 (wrap-around included, :244).            */
static int normalize_freqs(u32* freqs, u32* alphabet, u32 total, u32 scale)
{
    if (total == 0)
        return 0;
    int asz = 0;
    if (total == scale) {
        for (int i = 0; i < 256; i++)
            if (freqs[i] != 0)
                alphabet[asz++] = (u32)i;
        return asz;
    }
    u32 sumScaled = 0, sumFreq = 0;
    int idxMax = 0;
    for (int i = 0; i < 256; i++) {
        alphabet[i] = 0;
        const u32 f = freqs[i];
        if (f == 0)
            continue;
        alphabet[asz++] = (u32)i;
        const i64 sf = (i64)f * (i64)scale;
        const u32 sc = (sf <= (i64)total) ? 1u : (u32)((sf + ((i64)total >> 1)) / (i64)total);
        sumScaled += sc;
        freqs[i] = sc;
        sumFreq += f;
        if (sc > freqs[idxMax])
            idxMax = i;
        if (sumFreq >= total)
            break;
    }
    if (asz == 0)
        return 0;
    if (asz == 1) {
        freqs[alphabet[0]] = scale;
        return 1;
    }
    if (sumScaled == scale)
        return asz;
    int delta = (int)(sumScaled - scale);
    const int errThr = (int)freqs[idxMax] >> 4;
    if (abs(delta) <= errThr) {
        freqs[idxMax] -= (u32)delta;
        return asz;
    }
    if (delta < 0) {
        delta += errThr;
        freqs[idxMax] += (u32)errThr;
    } else {
        delta -= errThr;
        freqs[idxMax] -= (u32)errThr;
    }
    const int inc = (delta < 0) ? 1 : -1;
    delta = abs(delta);
    int round = 0;
    while ((++round < 6) && (delta > 0)) {
        int adjustments = 0;
        for (int i = 0; i < asz; i++) {
            const u32 idx = alphabet[i];
            if (freqs[idx] <= 2)
                continue;
            freqs[idx] += (u32)inc;
            adjustments++;
            delta--;
            if (delta == 0)
                break;
        }
        if (adjustments == 0)
            break;
    }
    {
        const u32 v = freqs[idxMax] - (u32)delta;
        freqs[idxMax] = (v > 1u) ? v : 1u;
    }
    return asz;
}

/* ------------------------------------------------------------------ rANS
 * entropy/ANSRangeEncoder.hpp:92-116 ANSEncSymbol::reset                      */
#define ANS_TOP (1 << 15)
typedef struct {
    int xMax, bias, cmplFreq, invShift;
    u64 invFreq;
} EncSym;

static void encsym_reset(EncSym* s, int cum, int freq, int lr)
{
    if (freq >= (1 << lr))
        freq = (1 << lr) - 1;
    s->xMax = ((ANS_TOP >> lr) << 16) * freq;
    s->cmplFreq = (1 << lr) - freq;
    if (freq < 2) {
        s->invFreq = 0xFFFFFFFFull;
        s->invShift = 32;
        s->bias = cum + (1 << lr) - 1;
    } else {
        int shift = 0;
        while (freq > (1 << shift))
            shift++;
        s->invFreq = ((((u64)1 << (shift + 31)) + (u64)freq - 1) / (u64)freq) & 0xFFFFFFFFull;
        s->invShift = 32 + shift - 1;
        s->bias = cum;
    }
}

/* entropy/ANSRangeEncoder.hpp:119-131 encodeSymbol (buffer grows downwards) */
static int enc_step(u8** pp, int st, const EncSym* s)
{
    u8* p = *pp;
    if (st >= s->xMax) {
        *p-- = (u8)st;
        *p-- = (u8)(st >> 8);
        st >>= 16;
    }
    *pp = p;
    return st + s->bias + (int)(((u64)(i64)st * s->invFreq) >> s->invShift) * s->cmplFreq;
}

/* entropy/ANSRangeEncoder.cpp:83-116 updateFrequencies + :119-155 encodeHeader
 * (order 0: one table; order 1: 256 tables, freqs row stride 257).           */
static int ans_write_tables(BitW* w, u32* freqs, EncSym* syms, int order, int lr)
{
    int res = 0;
    const int endk = 255 * order + 1;
    bw_put(w, (u64)(lr - 8), 3);
    u32 alphabet[256];
    for (int k = 0; k < endk; k++) {
        u32* f = &freqs[k * 257];
        const int asz = normalize_freqs(f, alphabet, f[256], 1u << lr);
        if (asz > 0) {
            int sum = 0, count = 0;
            for (int i = 0; i < 256; i++) {
                if (f[i] == 0)
                    continue;
                encsym_reset(&syms[(k << 8) + i], sum, (int)f[i], lr);
                sum += (int)f[i];
                if (++count >= asz)
                    break;
            }
        }
        put_alphabet(w, alphabet, asz);
        if (asz > 1) {
            const int chk = (asz >= 64) ? 8 : 6;
            const int llr = ilog2((u32)lr) + 1;
            for (int i = 1; i < asz; i += chk) {
                const int endj = (i + chk < asz) ? i + chk : asz;
                u32 mx = f[alphabet[i]] - 1;
                for (int j = i + 1; j < endj; j++)
                    if (f[alphabet[j]] - 1 > mx)
                        mx = f[alphabet[j]] - 1;
                const int logMax = (mx == 0) ? 0 : ilog2(mx) + 1;
                bw_put(w, (u64)logMax, llr);
                if (logMax == 0)
                    continue;
                for (int j = i; j < endj; j++)
                    bw_put(w, f[alphabet[j]] - 1, logMax);
            }
        }
        res += asz;
    }
    return res;
}

/* Global.cpp:170-309 computeHistogram as used by rebuildStatistics
 * (ANSRangeEncoder.cpp:264-287): order 0 with total; order 1 = four quarter
 * streams each starting in context 0, no total -> row totals are added here
 * the way the order-1 "withTotal=false" call leaves them: NOT computed by the
 * histogram; updateFrequencies reads f[256] ... see note in ans_encode.       */
static void histo_o1_quarter(const u8* p, int len, u32* freqs)
{
    /* Global.cpp:271-307 (order 1, no total): length<32 -> single stream from
     * context 0; else 4 sub-streams; sub-stream 0 starts in context 0, the
     * others in the context of the byte preceding them.                       */
    if (len < 32) {
        u32 prv = 0;
        for (int i = 0; i < len; i++) {
            freqs[prv + p[i]]++;
            prv = 256u * p[i];
        }
        return;
    }
    const int q = len >> 2;
    for (int s = 0; s < 4; s++) {
        u32 prv = (s == 0) ? 0 : 256u * p[s * q - 1];
        const int end = (s == 3) ? len : (s + 1) * q;
        for (int i = s * q; i < end; i++) {
            freqs[prv + p[i]]++;
            prv = 256u * p[i];
        }
    }
}

/* ANS block encoder.  entropy/ANSRangeEncoder.cpp:158-192 encode, :194-261
 * encodeChunk, :264-287 rebuildStatistics.  chunkSize: order 0 -> 16384,
 * order 1 -> 16384<<8; logRange 12 / 11 (ANSRangeEncoder.cpp:59-67).        */
static void ans_encode(BitW* w, const u8* block, u32 count, int order)
{
    if (count <= 32) {
        bw_put_bytes(w, block, 8 * (i64)count);
        return;
    }
    const u32 chunkSize = (order == 0) ? 16384u : (16384u << 8);
    const int lr = (order == 0) ? 12 : 11;
    const int dim = 255 * order + 1;
    u32* freqs = (u32*)malloc(sizeof(u32) * 257 * (size_t)dim);
    EncSym* syms = (EncSym*)malloc(sizeof(EncSym) * 256 * (size_t)dim);
    const u32 bufSize = 2 * chunkSize + 65536;
    u8* buf = (u8*)malloc(bufSize);
    u32 start = 0;
    while (start < count) {
        const u32 sz = (chunkSize < count - start) ? chunkSize : count - start;
        const u8* b = block + start;
        memset(freqs, 0, sizeof(u32) * 257 * (size_t)dim);
        if (order == 0) {
            for (u32 i = 0; i < sz; i++)
                freqs[b[i]]++;
            freqs[256] = sz;
        } else {
            /* rebuildStatistics :273-284 calls the order-1 histogram with
             * withTotal=true on each quarter (row stride 257, totals kept).   */
            const int quarter = (int)(sz >> 2);
            u32* tmp = (u32*)calloc(65536, sizeof(u32));
            if (quarter == 0) {
                histo_o1_quarter(b, (int)sz, tmp);
            } else {
                for (int s = 0; s < 4; s++)
                    histo_o1_quarter(b + s * quarter, quarter, tmp);
            }
            for (int c = 0; c < 256; c++) {
                u32 tot = 0;
                for (int s = 0; s < 256; s++) {
                    freqs[c * 257 + s] = tmp[c * 256 + s];
                    tot += tmp[c * 256 + s];
                }
                freqs[c * 257 + 256] = tot;
            }
            free(tmp);
        }
        const int asz = ans_write_tables(w, freqs, syms, order, lr);
        if (asz <= 1 && order == 0) {
            start += sz;
            continue;
        }
        /* encodeChunk */
        int st0 = ANS_TOP, st1 = ANS_TOP, st2 = ANS_TOP, st3 = ANS_TOP;
        u8* p = &buf[bufSize - 1];
        u8* const p0 = p;
        const int end = (int)sz;
        const int end4 = end & -4;
        for (int i = end - 1; i >= end4; i--)
            *p-- = b[i];
        if (order == 0) {
            for (int i = end4 - 1; i > 0; i -= 4) {
                st0 = enc_step(&p, st0, &syms[b[i]]);
                st1 = enc_step(&p, st1, &syms[b[i - 1]]);
                st2 = enc_step(&p, st2, &syms[b[i - 2]]);
                st3 = enc_step(&p, st3, &syms[b[i - 3]]);
            }
        } else {
            const int quarter = end4 >> 2;
            int i0 = quarter - 2, i1 = 2 * quarter - 2, i2 = 3 * quarter - 2, i3 = end4 - 2;
            int prv0 = b[i0 + 1], prv1 = b[i1 + 1], prv2 = b[i2 + 1], prv3 = b[i3 + 1];
            for (; i0 >= 0; i0--, i1--, i2--, i3--) {
                const int c0 = b[i0], c1 = b[i1], c2 = b[i2], c3 = b[i3];
                st0 = enc_step(&p, st0, &syms[(c0 << 8) | prv0]);
                st1 = enc_step(&p, st1, &syms[(c1 << 8) | prv1]);
                st2 = enc_step(&p, st2, &syms[(c2 << 8) | prv2]);
                st3 = enc_step(&p, st3, &syms[(c3 << 8) | prv3]);
                prv0 = c0;
                prv1 = c1;
                prv2 = c2;
                prv3 = c3;
            }
            st0 = enc_step(&p, st0, &syms[prv0]);
            st1 = enc_step(&p, st1, &syms[prv1]);
            st2 = enc_step(&p, st2, &syms[prv2]);
            st3 = enc_step(&p, st3, &syms[prv3]);
        }
        put_varint(w, (u32)(p0 - p));
        bw_put(w, (u32)st0, 32);
        bw_put(w, (u32)st1, 32);
        bw_put(w, (u32)st2, 32);
        bw_put(w, (u32)st3, 32);
        if (p != p0)
            bw_put_bytes(w, p + 1, 8 * (i64)(p0 - p));
        start += sz;
    }
    free(buf);
    free(syms);
    free(freqs);
}

/* entropy/ANSRangeDecoder.cpp:80-175 decodeHeader, :177-216 decode, :218-292
 * decodeChunk, ANSRangeDecoder.hpp:92-103 decodeSymbol.  Returns bytes
 * decoded or -1.                                                             */
static int ans_decode(BitR* r, u8* block, u32 count, int order)
{
    if (count <= 32) {
        br_get_bytes(r, block, count);
        return (int)count;
    }
    const u32 chunkSize = (order == 0) ? 16384u : (16384u << 8);
    const int dim = 255 * order + 1;
    u32* freqs = (u32*)malloc(sizeof(u32) * 256 * (size_t)dim);
    u8* f2s = (u8*)malloc((size_t)dim << 15);
    uint16_t* cumf = (uint16_t*)malloc(sizeof(uint16_t) * 256 * (size_t)dim);
    uint16_t* frq = (uint16_t*)malloc(sizeof(uint16_t) * 256 * (size_t)dim);
    const u32 bufSize = 2 * chunkSize;
    u8* buf = (u8*)malloc(bufSize);
    u32 alphabet[256];
    u32 start = 0;
    int rc = (int)count;
    while (start < count && rc >= 0) {
        const u32 sz = (chunkSize < count - start) ? chunkSize : count - start;
        const int lr = 8 + (int)br_get(r, 3);
        const u32 scale = 1u << lr;
        const int llr = ilog2((u32)lr) + 1;
        int total = 0, lastAsz = 0;
        for (int k = 0; k < dim && rc >= 0; k++) {
            const int asz = get_alphabet(r, alphabet);
            if (asz == 0)
                continue;
            u32* f = &freqs[k << 8];
            memset(f, 0, sizeof(u32) * 256);
            const int chk = (asz >= 64) ? 8 : 6;
            u32 sum = 0;
            for (int i = 1; i < asz; i += chk) {
                const u32 logMax = (u32)br_get(r, llr);
                if (logMax > (u32)lr) {
                    rc = -1;
                    break;
                }
                const int endj = (i + chk < asz) ? i + chk : asz;
                for (int j = i; j < endj; j++) {
                    const u32 fr = (logMax == 0) ? 1u : (u32)br_get(r, (int)logMax) + 1u;
                    if (fr >= scale)
                        rc = -1;
                    f[alphabet[j]] = fr;
                    sum += fr;
                }
            }
            if (rc < 0 || scale <= sum) {
                rc = -1;
                break;
            }
            f[alphabet[0]] = scale - sum;
            sum = 0;
            for (int i = 0; i < 256; i++) {
                if (f[i] == 0)
                    continue;
                memset(&f2s[((size_t)k << lr) + sum], i, f[i]);
                cumf[(k << 8) + i] = (uint16_t)sum;
                frq[(k << 8) + i] = (f[i] >= scale) ? (uint16_t)(scale - 1) : (uint16_t)f[i];
                sum += f[i];
            }
            total += asz;
            lastAsz = asz;
        }
        if (rc < 0)
            break;
        if (total == 0) {
            rc = (int)start;
            break;
        }
        u8* out = block + start;
        if (order == 0 && total == 1) {
            (void)lastAsz;
            memset(out, (int)alphabet[0], sz);
            start += sz;
            continue;
        }
        u32 psz;
        if (get_varint(r, &psz) < 0 || psz >= (1u << 27) || psz > bufSize - 2) {
            rc = -1;
            break;
        }
        u32 st0 = (u32)br_get(r, 32), st1 = (u32)br_get(r, 32), st2 = (u32)br_get(r, 32), st3 = (u32)br_get(r, 32);
        memset(buf, 0, bufSize);
        br_get_bytes(r, buf, psz);
        const u8* p = buf;
        const u32 mask = scale - 1;
        const int count4 = (int)sz & -4;
#define DEC_STEP(st, k, s)                                                              \
    do {                                                                                \
        st = (u32)frq[((k) << 8) + (s)] * (st >> lr) + (st & mask) - cumf[((k) << 8) + (s)]; \
        if (st < ANS_TOP) {                                                             \
            st = (st << 16) | ((u32)p[0] << 8) | p[1];                                  \
            p += 2;                                                                     \
        }                                                                               \
    } while (0)
        if (order == 0) {
            for (int i = 0; i < count4; i += 4) {
                const u8 c3 = f2s[st3 & mask];
                out[i] = c3;
                DEC_STEP(st3, 0, c3);
                const u8 c2 = f2s[st2 & mask];
                out[i + 1] = c2;
                DEC_STEP(st2, 0, c2);
                const u8 c1 = f2s[st1 & mask];
                out[i + 2] = c1;
                DEC_STEP(st1, 0, c1);
                const u8 c0 = f2s[st0 & mask];
                out[i + 3] = c0;
                DEC_STEP(st0, 0, c0);
            }
        } else {
            const int quarter = count4 >> 2;
            int prv0 = 0, prv1 = 0, prv2 = 0, prv3 = 0;
            for (int i = 0; i < quarter; i++) {
                const u8 c3 = f2s[((size_t)prv3 << lr) + (st3 & mask)];
                const u8 c2 = f2s[((size_t)prv2 << lr) + (st2 & mask)];
                const u8 c1 = f2s[((size_t)prv1 << lr) + (st1 & mask)];
                const u8 c0 = f2s[((size_t)prv0 << lr) + (st0 & mask)];
                DEC_STEP(st3, prv3, c3);
                DEC_STEP(st2, prv2, c2);
                DEC_STEP(st1, prv1, c1);
                DEC_STEP(st0, prv0, c0);
                out[3 * quarter + i] = c3;
                out[2 * quarter + i] = c2;
                out[quarter + i] = c1;
                out[i] = c0;
                prv3 = c3;
                prv2 = c2;
                prv1 = c1;
                prv0 = c0;
            }
        }
#undef DEC_STEP
        for (u32 i = (u32)count4; i < sz; i++)
            out[i] = *p++;
        if (p != buf + psz) {
            rc = -1;
            break;
        }
        start += sz;
    }
    free(buf);
    free(frq);
    free(cumf);
    free(f2s);
    free(freqs);
    return rc;
}

/* ------------------------------------------------------------------ Huffman
 * entropy/HuffmanEncoder.cpp:58-126 updateFrequencies, :129-215 limitCodeLengths,
 * :219-300 computeCodeLengths (Moffat-Katajainen in-place), :304-421 encode /
 * encodeChunk; entropy/HuffmanCommon.cpp:29-63 generateCanonicalCodes;
 * entropy/ExpGolombEncoder.hpp:51-62 (signed exp-Golomb of the code-length deltas);
 * decoder entropy/HuffmanDecoder.cpp:65-108 readLengths, :156-201 decodeV6,
 * :204-347 decodeChunk.  Chunk 16 KiB, codes <= 12 bits, 4 fragments per chunk.   */
#define HUF_MAX_LEN 12
#define HUF_CHUNK 16384

/* signed exp-Golomb: '1' for 0; else floor(log2(|v|+1)) zeros, (|v|+1) in binary, sign bit */
static void put_expgolomb_signed(BitW* w, int v)
{
    if (v == 0) {
        bw_put(w, 1, 1);
        return;
    }
    const u32 x = (u32)(v < 0 ? -v : v) + 1;
    const int lg = ilog2(x);
    bw_put(w, 0, lg);
    bw_put(w, x, lg + 1);
    bw_put(w, (u64)(v < 0 ? 1 : 0), 1);
}

static int get_expgolomb_signed(BitR* r) /* ExpGolombDecoder.hpp:52-75 */
{
    if (br_get(r, 1) == 1)
        return 0;
    u32 lg = 1;
    while (br_get(r, 1) == 0 && !r->underflow)
        lg++;
    lg &= 7;
    int res = (int)br_get(r, (int)lg + 1);
    const int sgn = res & 1;
    res = (res >> 1) + (1 << lg) - 1;
    return (int)(int8_t)((res - sgn) ^ -sgn);
}

static int cmp_u32(const void* a, const void* b)
{
    const u32 x = *(const u32*)a, y = *(const u32*)b;
    return (x > y) - (x < y);
}

/* computeCodeLengths: ranks[] = (freq << 8) | symbol, sorted increasing; sizes by symbol */
static int huf_code_lengths(uint16_t* sizes, u32* ranks, int count)
{
    qsort(ranks, (size_t)count, sizeof(u32), cmp_u32);
    u32 d[256] = { 0 };
    for (int i = 0; i < count; i++) {
        d[i] = ranks[i] >> 8;
        ranks[i] &= 0xFF;
        if (d[i] == 0)
            return 0;
    }
    const int n = count;
    /* phase 1 */
    for (int s = 0, rr = 0, t = 0; t < n - 1; t++) {
        u32 sum = 0;
        for (int i = 0; i < 2; i++) {
            if (s >= n || (rr < t && d[rr] < d[s])) {
                sum += d[rr];
                d[rr] = (u32)t;
                rr++;
                continue;
            }
            sum += d[s];
            if (s > t)
                d[s] = 0;
            s++;
        }
        d[t] = sum;
    }
    /* phase 2 */
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
    for (int i = 0; i < count; i++)
        sizes[ranks[i]] = (uint16_t)d[i];
    return (int)depth - 1;
}

/* EntropyUtils::normalizeFrequencies with an explicit array length (used by the
 * length limiter with length = alphabet size, EntropyUtils.cpp:131-245).          */
static int normalize_freqs_n(u32* freqs, u32* alphabet, int length, u32 total, u32 scale)
{
    if (length == 0 || total == 0)
        return 0;
    int asz = 0;
    if (total == scale) {
        for (int i = 0; i < 256; i++)
            if (freqs[i] != 0)
                alphabet[asz++] = (u32)i;
        return asz;
    }
    u32 sumScaled = 0, sumFreq = 0;
    int idxMax = 0;
    for (int i = 0; i < length; i++) {
        alphabet[i] = 0;
        const u32 f = freqs[i];
        if (f == 0)
            continue;
        alphabet[asz++] = (u32)i;
        const i64 sf = (i64)f * (i64)scale;
        const u32 sc = (sf <= (i64)total) ? 1u : (u32)((sf + ((i64)total >> 1)) / (i64)total);
        sumScaled += sc;
        freqs[i] = sc;
        sumFreq += f;
        if (sc > freqs[idxMax])
            idxMax = i;
        if (sumFreq >= total)
            break;
    }
    if (asz == 0)
        return 0;
    if (asz == 1) {
        freqs[alphabet[0]] = scale;
        return 1;
    }
    if (sumScaled == scale)
        return asz;
    int delta = (int)(sumScaled - scale);
    const int errThr = (int)freqs[idxMax] >> 4;
    if (abs(delta) <= errThr) {
        freqs[idxMax] -= (u32)delta;
        return asz;
    }
    if (delta < 0) {
        delta += errThr;
        freqs[idxMax] += (u32)errThr;
    } else {
        delta -= errThr;
        freqs[idxMax] -= (u32)errThr;
    }
    const int inc = (delta < 0) ? 1 : -1;
    delta = abs(delta);
    int round = 0;
    while ((++round < 6) && (delta > 0)) {
        int adjustments = 0;
        for (int i = 0; i < asz; i++) {
            const u32 idx = alphabet[i];
            if (freqs[idx] <= 2)
                continue;
            freqs[idx] += (u32)inc;
            adjustments++;
            delta--;
            if (delta == 0)
                break;
        }
        if (adjustments == 0)
            break;
    }
    {
        const u32 v = freqs[idxMax] - (u32)delta;
        freqs[idxMax] = (v > 1u) ? v : 1u;
    }
    return asz;
}

/* limitCodeLengths: ranks[] now holds symbols sorted by increasing frequency */
static int huf_limit_lengths(const u32* alphabet, u32* freqs, uint16_t* sizes, u32* ranks, int count)
{
    int n = 0, debt = 0;
    while (n < count && sizes[ranks[n]] >= HUF_MAX_LEN) {
        debt += sizes[ranks[n]] - HUF_MAX_LEN;
        sizes[ranks[n]] = HUF_MAX_LEN;
        n++;
    }
    if (debt == 0)
        return HUF_MAX_LEN;
    int v[6][256], vn[6] = { 0 }, vh[6] = { 0 };
    while (n < count) {
        const int idx = HUF_MAX_LEN - 1 - sizes[ranks[n]];
        if (idx > 5 || debt < (1 << idx))
            break;
        v[idx][vn[idx]++] = n;
        n++;
    }
    int idx = 5;
    while (debt > 0 && idx >= 0) {
        if (vh[idx] >= vn[idx] || debt < (1 << idx)) {
            idx--;
            continue;
        }
        sizes[ranks[v[idx][vh[idx]]]]++;
        debt -= 1 << idx;
        vh[idx]++;
    }
    idx = 0;
    while (debt > 0 && idx < 6) {
        if (vh[idx] >= vn[idx]) {
            idx++;
            continue;
        }
        sizes[ranks[v[idx][vh[idx]]]]++;
        debt -= 1 << idx;
        vh[idx]++;
    }
    if (debt > 0) {
        u32 alpha[256] = { 0 }, f[256] = { 0 }, total = 0;
        for (int i = 0; i < count; i++) {
            f[i] = freqs[alphabet[i]];
            total += f[i];
        }
        normalize_freqs_n(f, alpha, count, total, HUF_CHUNK >> 3);
        for (int i = 0; i < count; i++) {
            freqs[alphabet[i]] = f[i];
            ranks[i] = (f[i] << 8) | alphabet[i];
        }
        return huf_code_lengths(sizes, ranks, count);
    }
    return HUF_MAX_LEN;
}

/* generateCanonicalCodes: symbols ordered by (length, symbol) get consecutive codes */
static int huf_canonical(const uint16_t* sizes, uint16_t* codes, u32* symbols, int count)
{
    if (count == 0)
        return 0;
    if (count > 1) {
        u8 present[(HUF_MAX_LEN << 8) + 256];
        memset(present, 0, sizeof(present));
        for (int i = 0; i < count; i++) {
            const u32 sy = symbols[i];
            if (sy > 255 || sizes[sy] > HUF_MAX_LEN || sizes[sy] == 0)
                return -1;
            present[((sizes[sy] - 1) << 8) | sy] = 1;
        }
        for (int i = 0, n = 0; n < count; i++) {
            symbols[n] = (u32)(i & 0xFF);
            n += present[i];
        }
    }
    int curLen = sizes[symbols[0]];
    for (int i = 0, code = 0; i < count; i++) {
        const u32 sy = symbols[i];
        code <<= (sizes[sy] - curLen);
        curLen = sizes[sy];
        codes[sy] = (uint16_t)code;
        code++;
    }
    return count;
}

/* updateFrequencies: writes alphabet + length deltas, fills codes[] = (len << 12) | code */
static int huf_update(BitW* w, u32* freqs, uint16_t* codes)
{
    int count = 0;
    uint16_t sizes[256] = { 0 };
    u32 alphabet[256] = { 0 };
    for (int i = 0; i < 256; i++) {
        codes[i] = 0;
        if (freqs[i] > 0)
            alphabet[count++] = (u32)i;
    }
    put_alphabet(w, alphabet, count);
    if (count == 0)
        return 0;
    if (count == 1) {
        codes[alphabet[0]] = 1 << 12;
        sizes[alphabet[0]] = 1;
    } else {
        u32 ranks[256];
        for (int i = 0; i < count; i++)
            ranks[i] = (freqs[alphabet[i]] << 8) | alphabet[i];
        int maxLen = huf_code_lengths(sizes, ranks, count);
        if (maxLen > HUF_MAX_LEN)
            maxLen = huf_limit_lengths(alphabet, freqs, sizes, ranks, count);
        if (maxLen > HUF_MAX_LEN) {
            for (int i = 0; i < count; i++) {
                codes[alphabet[i]] = (uint16_t)i;
                sizes[alphabet[i]] = 8;
            }
        } else {
            huf_canonical(sizes, codes, ranks, count);
        }
    }
    int prev = 2;
    for (int i = 0; i < count; i++) {
        const u32 sy = alphabet[i];
        codes[sy] |= (uint16_t)(sizes[sy] << 12);
        put_expgolomb_signed(w, (int)(int8_t)(sizes[sy] - prev));
        prev = sizes[sy];
    }
    return count;
}

static void huf_encode(BitW* w, const u8* block, u32 count)
{
    uint16_t codes[256];
    u32 start = 0;
    while (start < count) {
        const u32 sz = (HUF_CHUNK < count - start) ? HUF_CHUNK : count - start;
        const u8* b = block + start;
        if (sz < 32) {
            bw_put_bytes(w, b, 8 * (i64)sz);
        } else {
            u32 freqs[256] = { 0 };
            for (u32 i = 0; i < sz; i++)
                freqs[b[i]]++;
            if (huf_update(w, freqs, codes) > 1) {
                /* encodeChunk: 4 fragments of sz/4 symbols, each its own bit string */
                const u32 frag = sz / 4;
                u32 nbits[4];
                for (int j = 0; j < 4; j++) {
                    u32 bits = 0;
                    for (u32 i = 0; i < frag; i++)
                        bits += codes[b[j * frag + i]] >> 12;
                    nbits[j] = bits;
                }
                for (int j = 0; j < 4; j++)
                    put_varint(w, nbits[j]);
                for (int j = 0; j < 4; j++)
                    for (u32 i = 0; i < frag; i++) {
                        const uint16_t c = codes[b[j * frag + i]];
                        bw_put(w, (u64)(c & 0x0FFF), c >> 12);
                    }
                for (u32 i = 4 * frag; i < sz; i++)
                    bw_put(w, b[i], 8);
            }
        }
        start += sz;
    }
}

static int huf_decode(BitR* r, u8* block, u32 count)
{
    u32 start = 0;
    while (start < count) {
        const u32 sz = (HUF_CHUNK < count - start) ? HUF_CHUNK : count - start;
        u8* out = block + start;
        if (sz < 32) {
            br_get_bytes(r, out, sz);
            start += sz;
            continue;
        }
        u32 alphabet[256];
        uint16_t sizes[256] = { 0 }, codes[256] = { 0 };
        const int asz = get_alphabet(r, alphabet);
        if (asz <= 0)
            return (int)start;
        int cur = 2;
        for (int i = 0; i < asz; i++) {
            cur += get_expgolomb_signed(r);
            cur = (int)(int8_t)cur;
            if (cur <= 0 || cur > HUF_MAX_LEN)
                return -1;
            sizes[alphabet[i]] = (uint16_t)cur;
        }
        if (asz == 1) {
            memset(out, (int)alphabet[0], sz);
            start += sz;
            continue;
        }
        if (huf_canonical(sizes, codes, alphabet, asz) < 0)
            return -1;
        /* 12-bit direct table: (symbol << 8) | length */
        static uint16_t table[1 << HUF_MAX_LEN];
        memset(table, 0, sizeof(table));
        for (int i = 0; i < asz; i++) {
            const u32 sy = alphabet[i];
            const int wdt = 1 << (HUF_MAX_LEN - sizes[sy]);
            const int idx = codes[sy] * wdt;
            if (idx + wdt > (1 << HUF_MAX_LEN))
                return -1;
            for (int k = 0; k < wdt; k++)
                table[idx + k] = (uint16_t)((sy << 8) | sizes[sy]);
        }
        u32 nbits[4];
        for (int j = 0; j < 4; j++)
            if (get_varint(r, &nbits[j]) < 0)
                return -1;
        const u32 frag = sz / 4;
        for (int j = 0; j < 4; j++) {
            const i64 fragEnd = r->pos + nbits[j];
            for (u32 i = 0; i < frag; i++) {
                /* peek 12 bits (zero padded past the fragment) */
                u32 v = 0;
                for (int k = 0; k < HUF_MAX_LEN; k++) {
                    const i64 p = r->pos + k;
                    u32 bit = 0;
                    if (p < fragEnd && p < r->nbits)
                        bit = (r->buf[p >> 3] >> (7 - (p & 7))) & 1;
                    v = (v << 1) | bit;
                }
                const uint16_t e = table[v];
                if ((e & 0xFF) == 0)
                    return -1;
                out[j * frag + i] = (u8)(e >> 8);
                r->pos += e & 0xFF;
            }
            if (r->pos != fragEnd)
                return -1;
        }
        for (u32 i = 4 * frag; i < sz; i++)
            out[i] = (u8)br_get(r, 8);
        start += sz;
    }
    return (int)count;
}

/* ------------------------------------------------------------------ ZRLT
 * transform/ZRLT.cpp:27-117 forward.  Returns 1 (ok), 0 (stage refused).     */
static int zrlt_forward(const u8* src, int n, u8* dst, int cap, int* outLen)
{
    *outLen = 0;
    if (n == 0)
        return 1;
    if (cap < n) /* getMaxEncodedLength(n) == n, ZRLT.hpp:43 */
        return 0;
    u32 s = 0, d = 0;
    const u32 sEnd = (u32)n, dEnd = (u32)cap;
    while (s < sEnd) {
        if (src[s] == 0) {
            u32 run = 1;
            while (s + run < sEnd && src[s + run] == 0)
                run++;
            s += run;
            run++;
            int lg = ilog2(run);
            if ((u32)lg > dEnd - d)
                return 0;
            while (lg > 0) {
                lg--;
                dst[d++] = (u8)((run >> lg) & 1);
            }
            continue;
        }
        const int v = src[s];
        const u32 need = (v >= 0xFE) ? 2u : 1u;
        if (need > dEnd - d)
            return 0;
        if (v >= 0xFE) {
            dst[d++] = 0xFF;
            dst[d++] = (u8)(v - 0xFE);
        } else {
            dst[d++] = (u8)(v + 1);
        }
        s++;
    }
    *outLen = (int)d;
    return 1;
}

/* transform/ZRLT.cpp:119-215 inverse */
static int zrlt_inverse(const u8* src, int n, u8* dst, int cap, int* outLen)
{
    *outLen = 0;
    if (n == 0)
        return 1;
    u32 s = 0, d = 0, run = 0;
    const u32 sEnd = (u32)n, dEnd = (u32)cap;
    for (;;) {
        u32 v = src[s];
        if (v <= 1) {
            run = 1;
            int ended = 0;
            do {
                run += run + v;
                s++;
                if (s >= sEnd) {
                    ended = 1;
                    break;
                }
                v = src[s];
            } while (v <= 1);
            if (ended)
                break; /* goto End with run pending */
            run--;
            if (run > 0) {
                if (run >= dEnd - d)
                    break; /* goto End (run still > 0) */
                memset(dst + d, 0, run);
                d += run;
                run = 0;
                continue;
            }
        }
        if (d >= dEnd)
            return 0;
        if (v == 0xFF) {
            s++;
            if (s >= sEnd)
                return 0;
            dst[d] = (u8)(0xFE + src[s]);
        } else {
            dst[d] = (u8)(v - 1);
        }
        s++;
        d++;
        if (s >= sEnd || d >= dEnd)
            break;
    }
    if (run > 0) {
        run--;
        if (run > dEnd - d)
            return 0;
        if (run > 0) {
            memset(dst + d, 0, run);
            d += run;
        }
    }
    *outLen = (int)d;
    return s == sEnd;
}

/* ------------------------------------------------------------------ SBRT
 * transform/SBRT.cpp:46-97 forward / :99-145 inverse; mode 1 MTFT, 2 RANK,
 * 3 TIMESTAMP (:20-32 masks).                                                */
static void sbrt_masks(int mode, int* m1, int* m2, int* sh)
{
    *m1 = (mode == 3) ? 0 : -1;
    *m2 = (mode == 1) ? 0 : -1;
    *sh = (mode == 2) ? 1 : 0;
}

static void sbrt_forward(const u8* src, int n, u8* dst, int mode)
{
    int m1, m2, sh, p[256] = { 0 }, q[256] = { 0 };
    u8 s2r[256], r2s[256];
    sbrt_masks(mode, &m1, &m2, &sh);
    for (int i = 0; i < 256; i++)
        s2r[i] = r2s[i] = (u8)i;
    for (int i = 0; i < n; i++) {
        const u8 c = src[i];
        int r = s2r[c];
        dst[i] = (u8)r;
        const int qc = ((i & m1) + (p[c] & m2)) >> sh;
        p[c] = i;
        q[c] = qc;
        while (r > 0 && q[r2s[r - 1]] <= qc) {
            r2s[r] = r2s[r - 1];
            s2r[r2s[r]] = (u8)r;
            r--;
        }
        r2s[r] = c;
        s2r[c] = (u8)r;
    }
}

static void sbrt_inverse(const u8* src, int n, u8* dst, int mode)
{
    int m1, m2, sh, p[256] = { 0 }, q[256] = { 0 };
    u8 r2s[256];
    sbrt_masks(mode, &m1, &m2, &sh);
    for (int i = 0; i < 256; i++)
        r2s[i] = (u8)i;
    for (int i = 0; i < n; i++) {
        int r = src[i];
        const int c = r2s[r];
        dst[i] = (u8)c;
        const int qc = ((i & m1) + (p[c] & m2)) >> sh;
        p[c] = i;
        q[c] = qc;
        while (r > 0 && q[r2s[r - 1]] <= qc) {
            r2s[r] = r2s[r - 1];
            r--;
        }
        r2s[r] = (u8)c;
    }
}

/* ------------------------------------------------------------------ BWT
 * The BWT is canonical: any correct suffix sorter yields the reference's bytes.
 * Restates the OUTPUT CONTRACT of BWT::forward (transform/BWT.cpp:92-134) and
 * DivSufSort::computeBWT/constructBWT (transform/DivSufSort.cpp:171-295):
 *   out[0] = in[n-1]; suffix rank r (end-of-string smallest) of suffix p>=1 goes to
 *   out[r + (r < rank(0))] = in[p-1];  primaryIndex[k] = rank(k*step)+1,
 *   step = ceil(n/chunks), chunks = 8 if n >= 256 else 1 (BWT.hpp:130).
 * Suffix ranks by prefix doubling with two counting-sort passes per round.   */
static void suffix_ranks(const u8* s, int n, int* rank)
{
    int* sa = (int*)malloc(sizeof(int) * (size_t)n);
    int* tmp = (int*)malloc(sizeof(int) * (size_t)n);
    int* cnt = (int*)malloc(sizeof(int) * ((size_t)n + 2));
    int* nr = (int*)malloc(sizeof(int) * (size_t)n);
    {
        int c[257] = { 0 };
        for (int i = 0; i < n; i++)
            c[s[i] + 1]++;
        for (int i = 0; i < 256; i++)
            c[i + 1] += c[i];
        for (int i = 0; i < n; i++)
            rank[i] = c[s[i]]; /* bucket start */
        int c2[257];
        memcpy(c2, c, sizeof(c));
        for (int i = 0; i < n; i++)
            sa[c2[s[i]]++] = i;
    }
    for (int h = 1;; h <<= 1) {
        /* pass 1: by second key (rank[i+h]+1, 0 if past the end) */
        memset(cnt, 0, sizeof(int) * ((size_t)n + 2));
        for (int i = 0; i < n; i++)
            cnt[((i + h < n) ? rank[i + h] + 1 : 0) + 1]++;
        for (int i = 0; i <= n; i++)
            cnt[i + 1] += cnt[i];
        for (int j = 0; j < n; j++) {
            const int i = sa[j];
            tmp[cnt[(i + h < n) ? rank[i + h] + 1 : 0]++] = i;
        }
        /* pass 2: stable by first key */
        memset(cnt, 0, sizeof(int) * ((size_t)n + 2));
        for (int i = 0; i < n; i++)
            cnt[rank[i] + 1]++;
        for (int i = 0; i < n; i++)
            cnt[i + 1] += cnt[i];
        for (int j = 0; j < n; j++) {
            const int i = tmp[j];
            sa[cnt[rank[i]]++] = i;
        }
        int distinct = 0;
        for (int j = 0; j < n; j++) {
            const int i = sa[j];
            if (j > 0) {
                const int pi = sa[j - 1];
                const int a1 = (i + h < n) ? rank[i + h] : -1;
                const int a0 = (pi + h < n) ? rank[pi + h] : -1;
                if (rank[i] == rank[pi] && a1 == a0) {
                    nr[i] = nr[pi];
                    continue;
                }
            }
            nr[i] = j;
            distinct++;
        }
        memcpy(rank, nr, sizeof(int) * (size_t)n);
        if (distinct == n || h >= n)
            break;
    }
    free(nr);
    free(cnt);
    free(tmp);
    free(sa);
}

static int bwt_chunks(int n) { return (n < 256) ? 1 : 8; }

static void bwt_forward(const u8* in, int n, u8* out, int* pidx /*[8]*/)
{
    memset(pidx, 0, sizeof(int) * 8);
    if (n == 1) { /* BWT.cpp:109-115: single byte copied, no index touched */
        out[0] = in[0];
        return;
    }
    int* rank = (int*)malloc(sizeof(int) * (size_t)n);
    suffix_ranks(in, n, rank);
    const int chunks = bwt_chunks(n);
    const int st = n / chunks;
    const int step = (chunks * st == n) ? st : st + 1;
    out[0] = in[n - 1];
    for (int p = 1; p < n; p++)
        out[rank[p] + (rank[p] < rank[0] ? 1 : 0)] = in[p - 1];
    for (int p = 0; p < n; p += step)
        if (p / step < 8)
            pidx[p / step] = rank[p] + 1;
    free(rank);
}

/* transform/BWT.cpp:169-292 / :295-657: both inverse variants produce the
 * original text; restated as the plain psi walk from primaryIndex[0].        */
static int bwt_inverse(const u8* in, int n, u8* out, const int* pidx)
{
    if (n == 1) {
        out[0] = in[0];
        return 1;
    }
    const int p0 = pidx[0];
    if (p0 <= 0 || p0 > n)
        return 0;
    u32 c[257] = { 0 };
    for (int i = 0; i < n; i++)
        c[in[i] + 1]++;
    for (int i = 0; i < 256; i++)
        c[i + 1] += c[i];
    u32* nxt = (u32*)malloc(sizeof(u32) * (size_t)n);
    for (int i = 0; i < n; i++) {
        const u32 idx = (i == 0) ? 0u : ((i < p0) ? (u32)(i - 1) : (u32)i);
        nxt[c[in[i]]++] = idx;
    }
    /* symbol of sorted position t = F[t]; recover from the bucket bounds */
    u32 start[257] = { 0 };
    for (int i = 0; i < n; i++)
        start[in[i] + 1]++;
    for (int i = 0; i < 256; i++)
        start[i + 1] += start[i];
    u8* F = (u8*)malloc((size_t)n);
    for (int s = 0; s < 256; s++)
        memset(F + start[s], s, start[s + 1] - start[s]);
    u32 t = (u32)(p0 - 1);
    for (int k = 0; k < n; k++) {
        out[k] = F[t];
        t = nxt[t];
    }
    free(F);
    free(nxt);
    return 1;
}

/* transform/BWTBlockCodec.cpp:32-87 forward (v6 header: mode byte then
 * chunks x pIndexSize big-endian bytes of primaryIndex-1).                    */
static int bwtcodec_forward(const u8* in, int n, u8* out, int cap, int* outLen)
{
    *outLen = 0;
    if (n == 0)
        return 1;
    if (cap < n + 33)
        return 0;
    int lg = ilog2((u32)n);
    if (n & (n - 1))
        lg++;
    const int pisz = (lg + 7) >> 3;
    if (pisz <= 0 || pisz >= 5)
        return 0;
    const int chunks = bwt_chunks(n);
    const int hdr = 1 + chunks * pisz;
    int pidx[8];
    bwt_forward(in, n, out + hdr, pidx);
    out[0] = (u8)((ilog2((u32)chunks) << 2) | (pisz - 1));
    int k = 1;
    for (int i = 0; i < chunks; i++) {
        const int v = pidx[i] - 1;
        for (int sh = (pisz - 1) << 3; sh >= 0; sh -= 8)
            out[k++] = (u8)(v >> sh);
    }
    *outLen = n + hdr;
    return 1;
}

/* transform/BWTBlockCodec.cpp:89-168 inverse (bsVersion 6 branch) */
static int bwtcodec_inverse(const u8* in, int n, u8* out, int cap, int* outLen)
{
    *outLen = 0;
    if (n <= 1)
        return n == 0;
    const int mode = in[0];
    const int chunks = 1 << ((mode >> 2) & 7);
    const int pisz = (mode & 3) + 1;
    const int hdr = 1 + chunks * pisz;
    if (n < hdr)
        return 0;
    if (chunks != bwt_chunks(n - hdr))
        return 0;
    int pidx[8] = { 0 };
    int k = 1;
    for (int i = 0; i < chunks; i++) {
        u32 v = 0;
        for (int b = 0; b < pisz; b++)
            v = (v << 8) | in[k++];
        if (v >= 0x7FFFFFFFu)
            return 0;
        if (i < 8)
            pidx[i] = (int)v + 1;
    }
    const int m = n - hdr;
    if (m > cap)
        return 0;
    if (m == 0)
        return 1;
    if (!bwt_inverse(in + hdr, m, out, pidx))
        return 0;
    *outLen = m;
    return 1;
}


/* ------------------------------------------------------------------ SRT
 * transform/SRT.cpp:22-109 forward, :111-204 inverse, :206-244 preprocess,
 * :246-277 encodeHeader, :279-305 decodeHeader.  Sorted Rank Transform: classic
 * move-to-front whose list starts in order of first appearance, every rank written
 * into the bucket of its own symbol (buckets: frequency descending, symbol
 * ascending), behind a header of 256 varint frequencies.                       */
static int srt_preprocess(const u32* freqs, u8* symbols)
{
    int nb = 0;
    for (int i = 0; i < 256; i++)
        if (freqs[i] != 0)
            symbols[nb++] = (u8)i;
    int h = 4;
    while (h < nb)
        h = h * 3 + 1;
    do {
        h /= 3;
        for (int i = h; i < nb; i++) {
            const u8 t = symbols[i];
            int b;
            for (b = i - h; b >= 0; b -= h) {
                const int val = (int)(freqs[symbols[b]] - freqs[t]);
                if ((val >= 0) && ((val != 0) || (t >= symbols[b])))
                    break;
                symbols[b + h] = symbols[b];
            }
            symbols[b + h] = t;
        }
    } while (h != 1);
    return nb;
}

static int srt_forward(const u8* src, int length, u8* out, int cap, int* outLen)
{
    if (cap < length + 1024) /* getMaxEncodedLength, SRT.hpp:38 */
        return 0;
    u32 freqs[256] = { 0 };
    u8 s2r[256] = { 0 }, r2s[256] = { 0 };
    for (int i = 0, b = 0; i < length;) {
        const u8 c = src[i];
        int j = i + 1;
        while ((j < length) && (src[j] == c))
            j++;
        if (freqs[c] == 0) {
            r2s[b] = c;
            s2r[c] = (u8)b;
            b++;
        }
        freqs[c] += (u32)(j - i);
        i = j;
    }
    u8 symbols[256];
    int buckets[256] = { 0 };
    const int nb = srt_preprocess(freqs, symbols);
    for (int i = 0, pos = 0; i < nb; i++) {
        buckets[symbols[i]] = pos;
        pos += (int)freqs[symbols[i]];
    }
    int hdr = 0;
    for (int i = 0; i < 256; i++) {
        u32 f = freqs[i];
        for (int k = 0; k < 4 && f >= 128; k++) {
            out[hdr++] = (u8)(0x80 | f);
            f >>= 7;
        }
        out[hdr++] = (u8)f;
    }
    u8* dst = out + hdr;
    for (int i = 0; i < length;) {
        const u8 c = src[i];
        int r = s2r[c];
        int p = buckets[c];
        dst[p++] = (u8)r;
        if (r != 0) {
            do {
                const u8 t = r2s[r - 1];
                r2s[r] = t;
                s2r[t] = (u8)r;
                r--;
            } while (r != 0);
            r2s[0] = c;
            s2r[c] = 0;
        }
        i++;
        while ((i < length) && (src[i] == c)) {
            dst[p++] = 0;
            i++;
        }
        buckets[c] = p;
    }
    *outLen = length + hdr;
    return 1;
}

static int srt_inverse(const u8* in, int length, u8* dst, int cap, int* outLen)
{
    if (length < 256)
        return 0;
    u32 freqs[256] = { 0 };
    int hdr = 0;
    for (int i = 0; i < 256; i++) {
        u32 res = 0;
        int shift = 0;
        for (int j = 0; j < 5; j++) {
            if (hdr >= length)
                return 0;
            const u32 val = in[hdr++];
            res |= (val & 0x7F) << shift;
            if ((val & 0x80) == 0)
                break;
            if (j == 4)
                return 0;
            shift += 7;
        }
        freqs[i] = res;
    }
    length -= hdr;
    if (length < 0 || length > cap)
        return 0;
    const u8* src = in + hdr;
    u8 symbols[256] = { 0 };
    int nb = srt_preprocess(freqs, symbols);
    int buckets[256] = { 0 }, ends[256] = { 0 };
    u8 r2s[256] = { 0 };
    for (int i = 0, pos = 0; i < nb; i++) {
        const u8 c = symbols[i];
        if (pos < 0 || pos >= length)
            return 0;
        r2s[src[pos]] = c;
        buckets[c] = pos + 1;
        pos += (int)freqs[c];
        ends[c] = pos;
    }
    u8 c = r2s[0];
    for (int i = 0; i < length; i++) {
        dst[i] = c;
        if (buckets[c] < ends[c]) {
            if (buckets[c] >= length)
                return 0; /* the reference would read past the block here */
            const u8 r = src[buckets[c]++];
            if (r == 0)
                continue;
            memmove(&r2s[0], &r2s[1], r);
            r2s[r] = c;
            c = r2s[0];
        } else {
            if (nb == 1)
                continue;
            nb--;
            memmove(&r2s[0], &r2s[1], (size_t)nb);
            c = r2s[0];
        }
    }
    *outLen = length;
    return 1;
}

/* ------------------------------------------------------------------ sequence
 * transform ids: transform/TransformFactory.hpp:49-73.                       */
/* ------------------------------------------------------------------ LZ / LZX / LZP
 * transform/LZCodec.cpp:118-455 (LZXCodec<T>::forward), :470-610 (inverseV6),
 * :771-992 (LZPCodec), helpers transform/LZCodec.hpp:178-248.
 * LZ (extra = 0, 2^16 hash slots) and LZX (extra = 1, 2^19 slots, one more lazy position) share
 * the format: 13-byte header (end of literals, token count, distance bytes as LE int32, flag byte
 * 0000MMMD), then literals (with their long lengths inline) | tokens | distances | match lengths. */
static u64 lz_le64(const u8* p)
{
    u64 v = 0;
    for (int i = 7; i >= 0; i--)
        v = (v << 8) | p[i];
    return v;
}
static u32 lz_le32(const u8* p) { return (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24); }
static void lz_put_le32(u8* p, u32 v)
{
    p[0] = (u8)v, p[1] = (u8)(v >> 8), p[2] = (u8)(v >> 16), p[3] = (u8)(v >> 24);
}
#define LZ_MAX_MATCH (65535 + 254 + 4)
#define LZ_MAXD1 ((1 << 16) - 2)
#define LZ_MAXD2 ((1 << 24) - 2)
static int lz_max_len(int n, int lzp) { return ((n <= 1024) ? n + 16 : n + n / 64) + (lzp ? 0 : 2); }
static u32 lz_hash(const u8* p, int hashLog) /* LZCodec.hpp:188-191 */
{
    return (u32)(((lz_le64(p) << 24) * (u64)0x1E35A7BDu) >> (64 - hashLog));
}
static int lz_match(const u8* src, int a, int b, int maxMatch) /* LZCodec.hpp:229-246 */
{
    int n = 0;
    while (n + 8 <= maxMatch) {
        const u64 diff = lz_le64(src + a + n) ^ lz_le64(src + b + n);
        if (diff) {
            n += __builtin_ctzll(diff) >> 3;
            break;
        }
        n += 8;
    }
    return n;
}
static int lz_emit_len(u8* p, int length) /* LZCodec.hpp:194-210 */
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
static u32 lz_read_len(const u8* p, int* pos) /* LZCodec.hpp:212-227 */
{
    u32 res = p[(*pos)++];
    if (res < 254)
        return res;
    if (res == 254) {
        res += ((u32)p[*pos] << 8) | p[*pos + 1];
        *pos += 2;
        return res;
    }
    res += ((u32)p[*pos] << 16) | ((u32)p[*pos + 1] << 8) | p[*pos + 2];
    *pos += 3;
    return res;
}

static int lzx_forward(const u8* src, int count, u8* dst, int cap, int* outLen, int extra)
{
    if (count == 0) {
        *outLen = 0;
        return 1;
    }
    if (cap < lz_max_len(count, 0) || count < 24)
        return 0;
    const int hashLog = extra ? 19 : 16;
    int* hashes = (int*)calloc((size_t)1 << hashLog, sizeof(int));
    /* the three side streams; the sum of everything emitted stays below count or the stage fails */
    u8* tk = (u8*)malloc((size_t)count + 64);
    u8* mb = (u8*)malloc((size_t)count + 64);
    u8* ml = (u8*)malloc((size_t)count + 64);
    const int srcEnd = count - 16 - 2;
    const int maxDist = (srcEnd < 4 * LZ_MAXD1) ? LZ_MAXD1 : LZ_MAXD2;
    const int minMatch = 4; /* dataType DNA / SMALL_ALPHABET come from pre-transforms that are out of scope */
    dst[12] = (u8)(((maxDist == LZ_MAXD1) ? 0 : 1) | (((minMatch - 2) & 7) << 1));
    int srcIdx = 0, dstIdx = 13, anchor = 0, mIdx = 0, mLenIdx = 0, tkIdx = 0;
    int repd[2] = { count, count };
    int repIdx = 0, srcInc = 0, ok = 1;
    while (srcIdx < srcEnd) {
        int bestLen = 0;
        const u32 h0 = lz_hash(src + srcIdx, hashLog);
        const int ref0 = hashes[h0];
        hashes[h0] = srcIdx;
        const int srcIdx1 = srcIdx + 1;
        int ref = srcIdx1 - repd[repIdx];
        const int minRef = (srcIdx - maxDist > 0) ? srcIdx - maxDist : 0;
        const int lim1 = (srcEnd - srcIdx1 < LZ_MAX_MATCH) ? srcEnd - srcIdx1 : LZ_MAX_MATCH;
        if (ref > minRef && lz_le32(src + srcIdx1) == lz_le32(src + ref)) {
            bestLen = lz_match(src, srcIdx1, ref, lim1); /* most recent repeat distance first */
        } else {
            ref = srcIdx1 - repd[repIdx ^ 1];
            if (ref > minRef && lz_le32(src + srcIdx1) == lz_le32(src + ref))
                bestLen = lz_match(src, srcIdx1, ref, lim1);
        }
        if (bestLen < minMatch) {
            ref = ref0; /* the hash table's candidate */
            if (ref > minRef && lz_le32(src + srcIdx) == lz_le32(src + ref)) {
                const int lim0 = (srcEnd - srcIdx < LZ_MAX_MATCH) ? srcEnd - srcIdx : LZ_MAX_MATCH;
                bestLen = lz_match(src, srcIdx, ref, lim0);
            }
            if (bestLen < minMatch) { /* no match: skip faster and faster through incompressible data */
                srcIdx = srcIdx1 + (srcInc >> 6);
                srcInc++;
                repIdx = 0;
                continue;
            }
            if (srcIdx - ref != repd[0] && srcIdx - ref != repd[1]) {
                /* lazy evaluation: one (LZ) or two (LZX) positions further */
                const u32 h1 = lz_hash(src + srcIdx1, hashLog);
                const int ref1 = hashes[h1];
                hashes[h1] = srcIdx1;
                if (ref1 > minRef + 1 && lz_le32(src + srcIdx1 + bestLen - 3) == lz_le32(src + ref1 + bestLen - 3)) {
                    const int bl1 = lz_match(src, srcIdx1, ref1, lim1);
                    if (bl1 >= bestLen) {
                        ref = ref1;
                        bestLen = bl1;
                        srcIdx = srcIdx1;
                    }
                }
                if (extra) {
                    const int srcIdx2 = srcIdx1 + 1;
                    const u32 h2 = lz_hash(src + srcIdx2, hashLog);
                    const int ref2 = hashes[h2];
                    hashes[h2] = srcIdx2;
                    if (ref2 > minRef + 2 && lz_le32(src + srcIdx2 + bestLen - 3) == lz_le32(src + ref2 + bestLen - 3)) {
                        const int lim2 = (srcEnd - srcIdx2 < LZ_MAX_MATCH) ? srcEnd - srcIdx2 : LZ_MAX_MATCH;
                        const int bl2 = lz_match(src, srcIdx2, ref2, lim2);
                        if (bl2 >= bestLen) {
                            ref = ref2;
                            bestLen = bl2;
                            srcIdx = srcIdx2;
                        }
                    }
                }
            }
            while (srcIdx > anchor && ref > minRef && src[srcIdx - 1] == src[ref - 1]) { /* extend backwards */
                bestLen++;
                ref--;
                srcIdx--;
            }
            if (bestLen > LZ_MAX_MATCH) {
                ref += bestLen - LZ_MAX_MATCH;
                srcIdx += bestLen - LZ_MAX_MATCH;
                bestLen = LZ_MAX_MATCH;
            }
        } else { /* repeat-distance match found at srcIdx + 1 */
            if (bestLen >= LZ_MAX_MATCH || src[srcIdx] != src[ref - 1]) {
                srcIdx++;
                hashes[lz_hash(src + srcIdx, hashLog)] = srcIdx;
            } else {
                bestLen++;
                ref--;
            }
        }
        srcInc = 0;
        /* token LLLFFMMM (new distance on FF = 1..3 bytes) or LLLFFFMM (FFF = 000 / 001: repeat distance 0 / 1) */
        const int dist = srcIdx - ref;
        int token, mLenTh;
        if (dist == repd[0]) {
            token = 0x00;
            mLenTh = 3;
        } else if (dist == repd[1]) {
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
            mLenIdx += lz_emit_len(ml + mLenIdx, mLen - mLenTh);
        } else {
            token += mLen;
        }
        repd[1] = repd[0];
        repd[0] = dist;
        repIdx = 1;
        const int litLen = srcIdx - anchor;
        if (litLen == 0) {
            tk[tkIdx++] = (u8)token;
        } else {
            if (litLen >= 7) {
                if (litLen >= (1 << 24)) {
                    ok = 0;
                    break;
                }
                tk[tkIdx++] = (u8)((7 << 5) | token);
                dstIdx += lz_emit_len(dst + dstIdx, litLen - 7);
            } else {
                tk[tkIdx++] = (u8)((litLen << 5) | token);
            }
            memcpy(dst + dstIdx, src + anchor, (size_t)litLen);
            dstIdx += litLen;
        }
        if (dstIdx + tkIdx + mIdx + mLenIdx >= count) { /* every term only grows: the final test (:421) must fail */
            ok = 0;
            break;
        }
        anchor = srcIdx + bestLen;
        while (++srcIdx < anchor) /* index the positions the match covers */
            hashes[lz_hash(src + srcIdx, hashLog)] = srcIdx;
    }
    int produced = 0;
    if (ok) {
        const int litLen = count - anchor;
        if (dstIdx + litLen + tkIdx + mIdx + mLenIdx >= count) {
            ok = 0;
        } else {
            if (litLen >= 7) {
                tk[tkIdx++] = (u8)(7 << 5);
                dstIdx += lz_emit_len(dst + dstIdx, litLen - 7);
            } else {
                tk[tkIdx++] = (u8)(litLen << 5);
            }
            memcpy(dst + dstIdx, src + anchor, (size_t)litLen);
            dstIdx += litLen;
            lz_put_le32(dst, (u32)dstIdx);
            lz_put_le32(dst + 4, (u32)tkIdx);
            lz_put_le32(dst + 8, (u32)mIdx);
            memcpy(dst + dstIdx, tk, (size_t)tkIdx);
            dstIdx += tkIdx;
            memcpy(dst + dstIdx, mb, (size_t)mIdx);
            dstIdx += mIdx;
            memcpy(dst + dstIdx, ml, (size_t)mLenIdx);
            dstIdx += mLenIdx;
            produced = dstIdx;
            ok = dstIdx <= count - count / 100; /* :455 */
        }
    }
    free(hashes);
    free(tk);
    free(mb);
    free(ml);
    *outLen = produced;
    return ok;
}

static int lzx_inverse(const u8* src, int count, u8* dst, int dstEnd, int* outLen)
{
    *outLen = 0;
    if (count == 0)
        return 1;
    if (count < 13)
        return 0;
    int tkIdx = (int)lz_le32(src), mIdx = (int)lz_le32(src + 4), mLenIdx = (int)lz_le32(src + 8);
    if (tkIdx < 0 || mIdx < 0 || mLenIdx < 0)
        return 0;
    if (tkIdx < 13 || tkIdx > count || mIdx > count - tkIdx || mLenIdx > count - tkIdx - mIdx)
        return 0;
    mIdx += tkIdx;
    mLenIdx += mIdx;
    const int srcEnd = tkIdx - 13, litEnd = tkIdx;
    const int maxDist = ((src[12] & 1) == 0) ? LZ_MAXD1 : LZ_MAXD2;
    const int minMatch = ((src[12] >> 1) & 7) + 2;
    int srcIdx = 13, dstIdx = 0, repd0 = count, repd1 = count, ok = 1;
    for (;;) {
        /* (the reference reads past `count` on truncated input; refuse instead) */
        if (tkIdx >= count + 2 || mIdx > count + 2 || mLenIdx > count + 2) {
            ok = 0;
            break;
        }
        const int token = src[tkIdx++];
        int mLen, dist;
        if ((token & 0x18) == 0) {
            mLen = token & 3;
            mLen += (mLen == 3) ? minMatch + (int)lz_read_len(src, &mLenIdx) : minMatch;
            dist = (token & 4) ? repd1 : repd0;
        } else {
            mLen = token & 7;
            mLen += (mLen == 7) ? minMatch + (int)lz_read_len(src, &mLenIdx) : minMatch;
            dist = src[mIdx++];
            if (token & 0x10) {
                dist = (dist << 8) | src[mIdx++];
                if (token & 0x08)
                    dist = (dist << 8) | src[mIdx++];
            }
        }
        if (token >= 32) {
            const u32 litLen = (token >= 0xE0) ? 7u + lz_read_len(src, &srcIdx) : (u32)(token >> 5);
            if (litLen > (u32)(dstEnd - dstIdx) || litLen > (u32)(litEnd - srcIdx)) {
                ok = 0;
                break;
            }
            memcpy(dst + dstIdx, src + srcIdx, litLen);
            srcIdx += (int)litLen;
            dstIdx += (int)litLen;
            if (srcIdx >= srcEnd)
                break;
        }
        repd1 = repd0;
        repd0 = dist;
        const int mEnd = dstIdx + mLen;
        int ref = dstIdx - dist;
        if (ref < 0 || dist > maxDist || mEnd > dstEnd) {
            ok = 0;
            break;
        }
        while (dstIdx < mEnd)
            dst[dstIdx++] = dst[ref++];
    }
    *outLen = dstIdx;
    return ok && (srcIdx == srcEnd + 13);
}

static int lzp_forward(const u8* src, int count, u8* dst, int cap, int* outLen)
{
    *outLen = 0;
    if (count == 0)
        return 1;
    if (count < 4 || cap < lz_max_len(count, 1) || count < 128)
        return 0;
    const int srcEnd = count, dstEnd = count - (count >> 6);
    int* hashes = (int*)calloc(1 << 16, sizeof(int));
    memcpy(dst, src, 4);
    u32 ctx = lz_le32(src);
    int srcIdx = 4, dstIdx = 4, ok = 1;
    while (srcIdx < srcEnd - 64 && dstIdx < dstEnd) {
        const u32 h = (0x7FEB352Du * ctx) >> 16;
        const int ref = hashes[h];
        hashes[h] = srcIdx;
        int bestLen = 0;
        if (ref != 0 && lz_le64(src + ref + 56) == lz_le64(src + srcIdx + 56))
            bestLen = lz_match(src, srcIdx, ref, srcEnd - srcIdx);
        if (bestLen < 64) {
            const u32 val = src[srcIdx];
            ctx = (ctx << 8) | val;
            dst[dstIdx++] = src[srcIdx++];
            if (ref != 0 && val == 0xFC) { /* a literal that looks like the match flag is escaped */
                if (dstIdx >= dstEnd) {
                    ok = 0;
                    break;
                }
                dst[dstIdx++] = 0xFF;
            }
            continue;
        }
        srcIdx += bestLen;
        ctx = lz_le32(src + srcIdx - 4);
        dst[dstIdx++] = 0xFC;
        bestLen -= 64;
        while (bestLen >= 254 && dstIdx < dstEnd) {
            bestLen -= 254;
            dst[dstIdx++] = 0xFE;
        }
        if (dstIdx >= dstEnd) {
            ok = 0;
            break;
        }
        dst[dstIdx++] = (u8)bestLen;
    }
    while (ok && srcIdx < srcEnd && dstIdx < dstEnd) {
        const u32 h = (0x7FEB352Du * ctx) >> 16;
        const int ref = hashes[h];
        hashes[h] = srcIdx;
        const u32 val = src[srcIdx];
        ctx = (ctx << 8) | val;
        dst[dstIdx++] = src[srcIdx++];
        if (ref != 0 && val == 0xFC) {
            if (dstIdx >= dstEnd) {
                ok = 0;
                break;
            }
            dst[dstIdx++] = 0xFF;
        }
    }
    free(hashes);
    *outLen = dstIdx;
    return ok && srcIdx == count && dstIdx < dstEnd;
}

static int lzp_inverse(const u8* src, int count, u8* dst, int dstEnd, int* outLen)
{
    *outLen = 0;
    if (count == 0)
        return 1;
    if (count < 4 || dstEnd < count)
        return 0;
    int* hashes = (int*)calloc(1 << 16, sizeof(int));
    memcpy(dst, src, 4);
    u32 ctx = lz_le32(dst);
    int srcIdx = 4, dstIdx = 4, ok = 1;
    const int srcEnd = count;
    while (srcIdx < srcEnd) {
        const u32 h = (0x7FEB352Du * ctx) >> 16;
        int ref = hashes[h];
        hashes[h] = dstIdx;
        if (src[srcIdx] != 0xFC || ref == 0) {
            if (dstIdx >= dstEnd) {
                ok = 0;
                break;
            }
            ctx = (ctx << 8) | src[srcIdx];
            dst[dstIdx++] = src[srcIdx++];
            continue;
        }
        srcIdx++;
        if (srcIdx >= srcEnd) {
            ok = 0;
            break;
        }
        if (src[srcIdx] == 0xFF) {
            if (dstIdx >= dstEnd) {
                ok = 0;
                break;
            }
            ctx = (ctx << 8) | 0xFC;
            dst[dstIdx++] = 0xFC;
            srcIdx++;
            continue;
        }
        int mLen = 64;
        while (srcIdx < srcEnd && src[srcIdx] == 0xFE) {
            srcIdx++;
            mLen += 254;
        }
        if (srcIdx >= srcEnd) {
            ok = 0;
            break;
        }
        mLen += src[srcIdx++];
        if (dstIdx + mLen > dstEnd) {
            ok = 0;
            break;
        }
        for (int i = 0; i < mLen; i++)
            dst[dstIdx + i] = dst[ref + i];
        dstIdx += mLen;
        ctx = lz_le32(dst + dstIdx - 4);
    }
    free(hashes);
    *outLen = dstIdx;
    return ok && srcIdx == srcEnd;
}

enum { T_NONE = 0, T_BWT = 1, T_LZ = 3, T_ZRLT = 6, T_MTFT = 7, T_RANK = 8, T_SRT = 13, T_LZP = 14, T_LZX = 16 };

static int stage_max_len(int t, int n) /* getMaxEncodedLength of each stage */
{
    /* BWTBlockCodec.hpp:47-50 n + 33; SRT.hpp:38 n + 1024; LZCodec.hpp:91-95,158-161; others srcLen */
    if (t == T_LZ || t == T_LZX || t == T_LZP)
        return lz_max_len(n, t == T_LZP);
    return (t == T_BWT) ? n + 33 : (t == T_SRT) ? n + 1024 : n;
}

static int stage_forward(int t, const u8* in, int n, u8* out, int cap, int* outLen)
{
    switch (t) {
    case T_NONE: /* NullTransform: plain copy */
        if (cap < n)
            return 0;
        memcpy(out, in, (size_t)n);
        *outLen = n;
        return 1;
    case T_BWT:
        return bwtcodec_forward(in, n, out, cap, outLen);
    case T_ZRLT:
        return zrlt_forward(in, n, out, cap, outLen);
    case T_MTFT:
    case T_RANK:
        if (n > cap)
            return 0;
        sbrt_forward(in, n, out, (t == T_MTFT) ? 1 : 2);
        *outLen = n;
        return 1;
    case T_SRT:
        return srt_forward(in, n, out, cap, outLen);
    case T_LZ:
        return lzx_forward(in, n, out, cap, outLen, 0);
    case T_LZX:
        return lzx_forward(in, n, out, cap, outLen, 1);
    case T_LZP:
        return lzp_forward(in, n, out, cap, outLen);
    default:
        return -1;
    }
}

static int stage_inverse(int t, const u8* in, int n, u8* out, int cap, int* outLen)
{
    switch (t) {
    case T_NONE:
        if (cap < n)
            return 0;
        memcpy(out, in, (size_t)n);
        *outLen = n;
        return 1;
    case T_BWT:
        return bwtcodec_inverse(in, n, out, cap, outLen);
    case T_ZRLT:
        return zrlt_inverse(in, n, out, cap, outLen);
    case T_MTFT:
    case T_RANK:
        if (n > cap)
            return 0;
        sbrt_inverse(in, n, out, (t == T_MTFT) ? 1 : 2);
        *outLen = n;
        return 1;
    case T_SRT:
        return srt_inverse(in, n, out, cap, outLen);
    case T_LZ:
    case T_LZX:
        return lzx_inverse(in, n, out, cap, outLen);
    case T_LZP:
        return lzp_inverse(in, n, out, cap, outLen);
    default:
        return -1;
    }
}

/* Split the 48-bit transform word (first stage in the top 6 bits) the way
 * TransformFactory::newTransform does (TransformFactory.hpp:208-223): slot 0 is
 * always instantiated, later NONE slots are dropped.                         */
static int split_types(u64 ttype, int* types)
{
    int n = 0;
    for (int i = 0; i < 8; i++) {
        const int t = (int)((ttype >> (42 - 6 * i)) & 63);
        if (t != T_NONE || i == 0)
            types[n++] = t;
    }
    return n;
}

/* transform/TransformSequence.hpp:88-162 forward.  The reference ping-pongs
 * between the caller's two buffers (capacities inCap / outCap) and falls back
 * to a private buffer of `required` bytes when one is too small; capacities
 * matter because ZRLT refuses to run when its output would not fit.
 * Returns post-transform length, sets *skipFlags (bit 7-i = stage i skipped). */
static int sequence_forward(u64 ttype, const u8* in, int n, int inCap, u8* out, int outCap, int* skipFlags)
{
    int types[8];
    const int nt = split_types(ttype, types);
    int required = n;
    for (int i = 0; i < nt; i++) {
        const int m = stage_max_len(types[i], required);
        if (m > required)
            required = m;
    }
    const int big = (required > inCap ? required : inCap) > outCap ? (required > inCap ? required : inCap) : outCap;
    u8* bufs[3];
    int caps[3];
    bufs[0] = (u8*)malloc((size_t)big + 64); /* plays the role of `input`  */
    bufs[1] = (u8*)malloc((size_t)big + 64); /* plays the role of `output` */
    bufs[2] = NULL;                           /* private `buffer`           */
    caps[0] = inCap;
    caps[1] = outCap;
    caps[2] = 0;
    memcpy(bufs[0], in, (size_t)n);
    int ci = 0, co = 1, count = n, swaps = 0, flags = 0xFF;
    for (int i = 0; i < nt; i++) {
        if (caps[co] < required) {
            if (co == 0 || co == 1)
                co = 2;
            if (caps[2] < required) {
                free(bufs[2]);
                bufs[2] = (u8*)malloc((size_t)required + 64);
                caps[2] = required;
            }
        }
        int produced = 0;
        if (stage_forward(types[i], bufs[ci], count, bufs[co], caps[co], &produced) != 1)
            continue;
        flags &= ~(1 << (7 - i));
        count = produced;
        const int t = ci;
        ci = co;
        co = t;
        swaps++;
    }
    int result = count;
    if ((swaps & 1) == 0) {
        if (count > outCap || count > caps[ci])
            flags = 0xFF;
        else
            memmove(out, bufs[ci], (size_t)count);
    } else {
        /* odd number of swaps: the last stage wrote straight into `output`
         * (index 1) because EncodingTask sizes it >= required (:733-739).      */
        memcpy(out, bufs[ci], (size_t)count);
    }
    *skipFlags = flags;
    free(bufs[0]);
    free(bufs[1]);
    free(bufs[2]);
    return result;
}

/* transform/TransformSequence.hpp:165-247 inverse */
static int sequence_inverse(u64 ttype, int skipFlags, const u8* in, int n, u8* out, int outCap, int* outLen)
{
    int types[8];
    const int nt = split_types(ttype, types);
    *outLen = 0;
    if (n == 0)
        return 1;
    if (n > outCap)
        return 0;
    if ((skipFlags & 0xFF) == 0xFF) {
        memmove(out, in, (size_t)n);
        *outLen = n;
        return 1;
    }
    u8* a = (u8*)malloc((size_t)outCap + 64);
    u8* b = (u8*)malloc((size_t)outCap + 64);
    u8* cur = a;
    u8* nxt = b;
    const int aCap = outCap; /* every working buffer is at least output._length */
    if (n > outCap) {
        free(a);
        free(b);
        return 0;
    }
    memcpy(cur, in, (size_t)n);
    int count = n, ok = 1;
    for (int i = nt - 1; i >= 0 && ok; i--) {
        if (skipFlags & (1 << (7 - i)))
            continue;
        int produced = 0;
        if (stage_inverse(types[i], cur, count, nxt, aCap, &produced) != 1) {
            ok = 0;
            break;
        }
        count = produced;
        u8* t = cur;
        cur = nxt;
        nxt = t;
    }
    if (ok) {
        memcpy(out, cur, (size_t)count);
        *outLen = count;
    }
    free(a);
    free(b);
    return ok;
}


/* ------------------------------------------------------------------ FPAQ
 * entropy/FPAQEncoder.cpp:58-103 encode, FPAQEncoder.hpp:72-94 encodeBit/flush,
 * entropy/FPAQDecoder.cpp:61-120 decode, FPAQDecoder.hpp:74-118 decodeBit/read.
 * Binary arithmetic coder, 56-bit interval, 4 x 256 adaptive 16-bit probabilities
 * (row = top two bits of the previous byte, restarted at 0 for every 4 MiB chunk;
 * probabilities and the interval persist across the chunks of a block).  Per
 * chunk: varint byte count, the flushed 32-bit words, then (between chunks and at
 * dispose) 56 bits of low | 0xFFFFFF.                                           */
#define FPAQ_TOP 0x00FFFFFFFFFFFFFFull
#define FPAQ_CHUNK (4u << 20)
static void fpaq_encode(BitW* w, const u8* block, u32 count)
{
    u64 low = 0, high = FPAQ_TOP;
    u16 probs[4][256];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 256; j++)
            probs[i][j] = 32768;
    u8* buf = (u8*)malloc((size_t)FPAQ_CHUNK + (FPAQ_CHUNK >> 3) + 1024);
    u32 start = 0;
    while (start < count) {
        const u32 chunk = (FPAQ_CHUNK < count - start) ? FPAQ_CHUNK : count - start;
        u32 idx = 0;
        u16* p = probs[0];
        for (u32 i = start; i < start + chunk; i++) {
            const int val = block[i];
            const int bits = val + 256;
            for (int k = 7; k >= 0; k--) {
                u16* pr = &p[(k == 7) ? 1 : (bits >> (k + 1))];
                const int bit = (val >> k) & 1;
                const u64 split = (((high - low) >> 8) * (u64)(*pr)) >> 8;
                if (bit == 0) {
                    low = low + split + 1;
                    *pr -= (u16)(*pr >> 6);
                } else {
                    high = low + split;
                    *pr -= (u16)(((int)*pr - 65536 + 64) >> 6);
                }
                if (((low ^ high) >> 24) == 0) {
                    const u32 v = (u32)(high >> 24);
                    buf[idx++] = (u8)(v >> 24);
                    buf[idx++] = (u8)(v >> 16);
                    buf[idx++] = (u8)(v >> 8);
                    buf[idx++] = (u8)v;
                    low <<= 32;
                    high = (high << 32) | 0xFFFFFFFFull;
                }
            }
            p = probs[val >> 6];
        }
        put_varint(w, idx);
        bw_put_bytes(w, buf, 8 * (i64)idx);
        start += chunk;
        if (start < count)
            bw_put(w, (low | 0xFFFFFFull) & FPAQ_TOP, 56);
    }
    bw_put(w, (low | 0xFFFFFFull) & FPAQ_TOP, 56); /* dispose() */
    free(buf);
}

static int fpaq_decode(BitR* r, u8* block, u32 count)
{
    u64 low = 0, high = FPAQ_TOP, current = 0;
    u16 probs[4][256];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 256; j++)
            probs[i][j] = 32768;
    u32 start = 0;
    u8* buf = NULL;
    while (start < count) {
        u32 sz = 0;
        if (get_varint(r, &sz) < 0) {
            free(buf);
            return 0;
        }
        if (sz >= 2 * count) {
            free(buf);
            return 0;
        }
        free(buf);
        buf = (u8*)calloc((size_t)sz + (sz >> 3) + 8192, 1);
        current = br_get(r, 56);
        br_get_bytes(r, buf, sz);
        u32 idx = 0;
        const u32 chunk = (FPAQ_CHUNK < count - start) ? FPAQ_CHUNK : count - start;
        u16* p = probs[0];
        for (u32 i = start; i < start + chunk; i++) {
            int ctx = 1;
            for (int k = 0; k < 8; k++) {
                const u64 split = ((((high - low) >> 8) * (u64)p[ctx]) >> 8) + low;
                if (split >= current) {
                    high = split;
                    p[ctx] -= (u16)(((int)p[ctx] - 65536 + 64) >> 6);
                    ctx += ctx + 1;
                } else {
                    low = split + 1;
                    p[ctx] -= (u16)(p[ctx] >> 6);
                    ctx += ctx;
                }
                if (((low ^ high) >> 24) == 0) {
                    low = (low << 32) & FPAQ_TOP;
                    high = ((high << 32) | 0xFFFFFFFFull) & FPAQ_TOP;
                    if (idx + 4 > sz) {
                        current = (current << 32) & FPAQ_TOP;
                        idx = sz + 1;
                    } else {
                        const u64 val = ((u64)buf[idx] << 24) | ((u64)buf[idx + 1] << 16) | ((u64)buf[idx + 2] << 8) | buf[idx + 3];
                        current = ((current << 32) | val) & FPAQ_TOP;
                        idx += 4;
                    }
                }
            }
            block[i] = (u8)ctx;
            if (idx > sz) {
                free(buf);
                return 0;
            }
            p = probs[(ctx & 0xFF) >> 6];
        }
        start += chunk;
    }
    free(buf);
    return (int)count;
}

/* ------------------------------------------------------------------ blocks
 * entropy ids: entropy/EntropyEncoderFactory.hpp:37-52.                      */
enum { E_NONE = 0, E_HUFFMAN = 1, E_FPAQ = 2, E_ANS0 = 5, E_ANS1 = 8 };

static int entropy_encode(BitW* w, int etype, const u8* p, u32 n)
{
    switch (etype) {
    case E_NONE: /* NullEntropyEncoder: raw bytes */
        bw_put_bytes(w, p, 8 * (i64)n);
        return 0;
    case E_ANS0:
        ans_encode(w, p, n, 0);
        return 0;
    case E_ANS1:
        ans_encode(w, p, n, 1);
        return 0;
    case E_HUFFMAN:
        huf_encode(w, p, n);
        return 0;
    case E_FPAQ:
        fpaq_encode(w, p, n);
        return 0;
    default:
        return -1;
    }
}

static int entropy_decode(BitR* r, int etype, u8* p, u32 n)
{
    switch (etype) {
    case E_NONE:
        br_get_bytes(r, p, n);
        return (int)n;
    case E_ANS0:
        return ans_decode(r, p, n, 0);
    case E_ANS1:
        return ans_decode(r, p, n, 1);
    case E_HUFFMAN:
        return huf_decode(r, p, n);
    case E_FPAQ:
        return fpaq_decode(r, p, n);
    default:
        return -1;
    }
}

/* ================================================================== block checksums
 * util/XXHash.hpp:61-121 (XXHash32::hash) and :151-231 (XXHash64::hash), seed = BITSTREAM_TYPE
 * (io/CompressedOutputStream.cpp:99-110).  The 64-bit variant's first merge rotates by
 * (1, 7, 12, 18) with 32-bit complements on 64-bit values -- written here as the reference has it. */
static u32 xx_le32(const u8* p) { return (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24); }
static u64 xx_le64(const u8* p) { return (u64)xx_le32(p) | ((u64)xx_le32(p + 4) << 32); }
static u32 xx32_round(u32 acc, u32 val)
{
    acc += val * 0x85EBCA77u;
    return ((acc << 13) | (acc >> 19)) * 0x9E3779B1u;
}
static u32 xxhash32(const u8* data, int length)
{
    const u32 P1 = 0x9E3779B1u, P2 = 0x85EBCA77u, P3 = 0xC2B2AE3Du, P4 = 0x27D4EB2Fu, P5 = 0x165667B1u;
    const u32 seed = 0x4B414E5Au;
    u32 h;
    int idx = 0;
    if (length >= 16) {
        u32 v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
        do {
            v1 = xx32_round(v1, xx_le32(data + idx));
            v2 = xx32_round(v2, xx_le32(data + idx + 4));
            v3 = xx32_round(v3, xx_le32(data + idx + 8));
            v4 = xx32_round(v4, xx_le32(data + idx + 12));
            idx += 16;
        } while (idx <= length - 16);
        h = ((v1 << 1) | (v1 >> 31)) + ((v2 << 7) | (v2 >> 25)) + ((v3 << 12) | (v3 >> 20)) + ((v4 << 18) | (v4 >> 14));
    } else {
        h = seed + P5;
    }
    h += (u32)length;
    for (; idx <= length - 4; idx += 4) {
        h += xx_le32(data + idx) * P3;
        h = ((h << 17) | (h >> 15)) * P4;
    }
    for (; idx < length; idx++) {
        h += (u32)data[idx] * P5;
        h = ((h << 11) | (h >> 21)) * P1;
    }
    h ^= h >> 15;
    h *= P2;
    h ^= h >> 13;
    h *= P3;
    return h ^ (h >> 16);
}
static u64 xx64_round(u64 acc, u64 val)
{
    acc += val * 0xC2B2AE3D27D4EB4Full;
    return ((acc << 31) | (acc >> 33)) * 0x9E3779B185EBCA87ull;
}
static u64 xx64_merge(u64 acc, u64 val) { return (acc ^ xx64_round(0, val)) * 0x9E3779B185EBCA87ull + 0x85EBCA77C2B2AE63ull; }
static u64 xxhash64(const u8* data, int length)
{
    const u64 P1 = 0x9E3779B185EBCA87ull, P2 = 0xC2B2AE3D27D4EB4Full, P3 = 0x165667B19E3779F9ull,
              P4 = 0x85EBCA77C2B2AE63ull, P5 = 0x27D4EB2F165667C5ull;
    const u64 seed = 0x4B414E5Aull;
    u64 h;
    int idx = 0;
    if (length >= 32) {
        u64 v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
        do {
            v1 = xx64_round(v1, xx_le64(data + idx));
            v2 = xx64_round(v2, xx_le64(data + idx + 8));
            v3 = xx64_round(v3, xx_le64(data + idx + 16));
            v4 = xx64_round(v4, xx_le64(data + idx + 24));
            idx += 32;
        } while (idx <= length - 32);
        h = ((v1 << 1) | (v1 >> 31)) + ((v2 << 7) | (v2 >> 25)) + ((v3 << 12) | (v3 >> 20)) + ((v4 << 18) | (v4 >> 14));
        h = xx64_merge(h, v1);
        h = xx64_merge(h, v2);
        h = xx64_merge(h, v3);
        h = xx64_merge(h, v4);
    } else {
        h = seed + P5;
    }
    h += (u64)length;
    for (; idx + 8 <= length; idx += 8) {
        h ^= xx64_round(0, xx_le64(data + idx));
        h = ((h << 27) | (h >> 37)) * P1 + P4;
    }
    for (; idx + 4 <= length; idx += 4) {
        h ^= (u64)xx_le32(data + idx) * P1;
        h = ((h << 23) | (h >> 41)) * P2 + P3;
    }
    for (; idx < length; idx++) {
        h ^= (u64)data[idx] * P5;
        h = ((h << 11) | (h >> 53)) * P1;
    }
    h ^= h >> 33;
    h *= P2;
    h ^= h >> 29;
    h *= P3;
    return h ^ (h >> 32);
}
static u64 block_hash(const u8* data, int length, int ckBits) { return (ckBits == 32) ? (u64)xxhash32(data, length) : xxhash64(data, length); }

/* One block exactly as EncodingTask::run builds it in its private buffer
 * (io/CompressedOutputStream.cpp:652-898, skipBlocks off):
 * mode byte, [skip-flag byte], post-transform length, [block checksum], entropy payload.
 * dataCap / bufCap model _data->_length / _buffer->_length (the two ping-pong
 * capacities handed to TransformSequence::forward).  Returns bit count.       */
static i64 encode_block_ck(const u8* in, int n, u64 ttype, int etype, int dataCap, int bufCap, u8* out, i64 outCap,
                           int ckBits)
{
    BitW w;
    bw_init(&w, out, outCap);
    int mode = 0, flags = 0xFF, post = n, ntr = 1;
    u8* tbuf = NULL;
    const u8* payload = in;
    if (n <= 15) { /* SMALL_BLOCK_SIZE :38,691-695 */
        mode |= 0x80;
        ttype = 0;
        etype = E_NONE;
    }
    {
        int types[8];
        ntr = split_types(ttype, types);
        int required = n;
        for (int i = 0; i < ntr; i++) {
            const int m = stage_max_len(types[i], required);
            if (m > required)
                required = m;
        }
        if (bufCap < required)
            bufCap = required; /* :733-739 */
        tbuf = (u8*)malloc((size_t)bufCap + 64);
        post = sequence_forward(ttype, in, n, dataCap, tbuf, bufCap, &flags);
        payload = tbuf;
    }
    const int dataSize = (post < 256) ? 1 : (ilog2((u32)post) >> 3) + 1;
    mode |= ((dataSize - 1) & 3) << 5;
    if ((mode & 0x80) || ntr <= 4) {
        mode |= (flags >> 4) & 0x0F;
        bw_put(&w, (u64)mode, 8);
    } else {
        mode |= 0x10;
        bw_put(&w, (u64)mode, 8);
        bw_put(&w, (u64)flags, 8);
    }
    bw_put(&w, (u64)post, 8 * dataSize);
    if (ckBits) /* :674-682 (hash of the block before the transforms), :804-807 */
        bw_put(&w, block_hash(in, n, ckBits), ckBits);
    entropy_encode(&w, etype, payload, (u32)post);
    free(tbuf);
    return w.overflow ? -1 : w.bits;
}

static i64 encode_block(const u8* in, int n, u64 ttype, int etype, int dataCap, int bufCap, u8* out, i64 outCap)
{
    return encode_block_ck(in, n, ttype, etype, dataCap, bufCap, out, outCap, 0);
}

/* io/CompressedInputStream.cpp:791-1041 DecodingTask::run (payload part) */
static int decode_block_ck(const u8* in, i64 nbits, u64 ttype, int etype, int blockSize, u8* out, int outCap, int ckBits)
{
    BitR r;
    br_init(&r, in, nbits);
    const int mode = (int)br_get(&r, 8);
    int flags = 0;
    if (mode & 0x80) {
        ttype = 0;
        etype = E_NONE;
    } else if (mode & 0x10) {
        flags = (int)br_get(&r, 8);
    } else {
        flags = ((mode << 4) | 0x0F) & 0xFF;
    }
    const int dataSize = 1 + ((mode >> 5) & 3);
    const int pre = (int)br_get(&r, 8 * dataSize);
    const u64 ck = ckBits ? br_get(&r, ckBits) : 0; /* :895-900 */
    /* DecodingTask's _blockLength = blockSize + max(512, blockSize/16)
     * (CompressedInputStream.cpp:275): capacity of the task's data buffer.     */
    const int blkLen = blockSize + ((blockSize >> 4) > 512 ? (blockSize >> 4) : 512);
    int maxT = blkLen + blkLen / 2;
    if (maxT < 2048)
        maxT = 2048;
    if (pre <= 0 || pre > maxT)
        return -1;
    const int tmpCap = (pre + 512 > blkLen) ? pre + 512 : blkLen; /* :935 */
    u8* tmp = (u8*)malloc((size_t)tmpCap + 64);
    u8* data = (u8*)malloc((size_t)blkLen + 64);
    int rc = -1;
    if (entropy_decode(&r, etype, tmp, (u32)pre) == pre && !r.underflow) {
        int produced = 0;
        if (sequence_inverse(ttype, flags, tmp, pre, data, blkLen, &produced) && produced <= outCap) {
            memcpy(out, data, (size_t)produced);
            rc = produced;
            if (ckBits && block_hash(data, produced, ckBits) != ck) /* :1003-1022 */
                rc = -2;
        }
    }
    free(data);
    free(tmp);
    return rc;
}

static int decode_block(const u8* in, i64 nbits, u64 ttype, int etype, int blockSize, u8* out, int outCap)
{
    return decode_block_ck(in, nbits, ttype, etype, blockSize, out, outCap, 0);
}

/* Stream header, io/CompressedOutputStream.cpp:277-342 */
static void put_stream_header(BitW* w, u64 ttype, int etype, int blockSize, i64 inputSize, int ckBits)
{
    const u32 ckSize = (u32)(ckBits >> 5); /* 0, 1 (32 bits), 2 (64 bits) */
    bw_put(w, 0x4B414E5A, 32);
    bw_put(w, 6, 4);
    bw_put(w, ckSize, 2);
    bw_put(w, (u64)etype, 5);
    bw_put(w, ttype, 48);
    bw_put(w, (u64)(blockSize >> 4), 28);
    int szMask = 0;
    if (inputSize != 0 && inputSize < ((i64)1 << 48)) {
        int lg = 0;
        u64 x = (u64)inputSize;
        while (x > 1) {
            x >>= 1;
            lg++;
        }
        szMask = (lg >> 4) + 1;
    }
    bw_put(w, (u64)szMask, 2);
    if (szMask)
        bw_put(w, (u64)inputSize, 16 * szMask);
    bw_put(w, 0, 15);
    const u32 HASH = 0x1E35A7BDu;
    u32 ck = HASH * (0x01030507u * 6u);
    ck ^= HASH * (u32)~ckSize;
    ck ^= HASH * (u32)~(u32)etype;
    ck ^= HASH * (u32)((~ttype) >> 32);
    ck ^= HASH * (u32)(~ttype);
    ck ^= HASH * (u32)~(u32)blockSize;
    if (szMask) {
        ck ^= HASH * (u32)((~(u64)inputSize) >> 32);
        ck ^= HASH * (u32)(~(u64)inputSize);
    }
    ck = (ck >> 23) ^ (ck >> 3);
    bw_put(w, ck & 0xFFFFFFu, 24);
}

/* ================================================================== exports */
#define API __attribute__((visibility("default")))

API i64 ko_entropy_encode(int etype, const u8* in, int n, u8* out, i64 cap)
{
    BitW w;
    bw_init(&w, out, cap);
    if (entropy_encode(&w, etype, in, (u32)n) < 0 || w.overflow)
        return -1;
    return w.bits;
}

API int ko_entropy_decode(int etype, const u8* in, i64 nbits, u8* out, int n)
{
    BitR r;
    br_init(&r, in, nbits);
    const int rc = entropy_decode(&r, etype, out, (u32)n);
    return r.underflow ? -1 : rc;
}

API int ko_stage_forward(int t, const u8* in, int n, u8* out, int cap, int* outLen) { return stage_forward(t, in, n, out, cap, outLen); }
API int ko_stage_inverse(int t, const u8* in, int n, u8* out, int cap, int* outLen) { return stage_inverse(t, in, n, out, cap, outLen); }
API void ko_bwt_forward(const u8* in, int n, u8* out, int* pidx) { bwt_forward(in, n, out, pidx); }
API int ko_bwt_inverse(const u8* in, int n, u8* out, const int* pidx) { return bwt_inverse(in, n, out, pidx); }
API void ko_suffix_ranks(const u8* in, int n, int* rank) { suffix_ranks(in, n, rank); }

API int ko_sequence_forward(u64 ttype, const u8* in, int n, int inCap, u8* out, int outCap, int* skipFlags)
{
    return sequence_forward(ttype, in, n, inCap, out, outCap, skipFlags);
}

API int ko_sequence_inverse(u64 ttype, int skipFlags, const u8* in, int n, u8* out, int outCap, int* outLen)
{
    return sequence_inverse(ttype, skipFlags, in, n, out, outCap, outLen);
}

API i64 ko_encode_block(const u8* in, int n, u64 ttype, int etype, int dataCap, int bufCap, u8* out, i64 outCap)
{
    return encode_block(in, n, ttype, etype, dataCap, bufCap, out, outCap);
}

API int ko_decode_block(const u8* in, i64 nbits, u64 ttype, int etype, int blockSize, u8* out, int outCap)
{
    return decode_block(in, nbits, ttype, etype, blockSize, out, outCap);
}

/* Whole stream with the jobs=1 buffer model of CompressedOutputStream
 * (:138-146, :447-474): _buffers[0] = max(bs + bs/8, 256 KiB) is the data buffer
 * of every block; the transform buffer keeps the largest `required` seen.     */
API i64 ko_stream_compress_ck(const u8* in, i64 n, u64 ttype, int etype, int blockSize, int ckBits, u8* out, i64 cap)
{
    BitW w;
    bw_init(&w, out, cap);
    put_stream_header(&w, ttype, etype, blockSize, n, ckBits);
    const int dataCap = (blockSize + (blockSize >> 3) > 262144) ? blockSize + (blockSize >> 3) : 262144;
    int bufCap = 0;
    const i64 tmpCap = (i64)blockSize + (blockSize >> 1) + 65536;
    u8* tmp = (u8*)malloc((size_t)tmpCap);
    for (i64 off = 0; off < n; off += blockSize) {
        const int len = (n - off < blockSize) ? (int)(n - off) : blockSize;
        /* required size of this block's sequence */
        int types[8];
        u64 tt = (len <= 15) ? 0 : ttype;
        const int nt = split_types(tt, types);
        int required = len;
        for (int i = 0; i < nt; i++) {
            const int m = stage_max_len(types[i], required);
            if (m > required)
                required = m;
        }
        if (bufCap < required)
            bufCap = required;
        const i64 bits = encode_block_ck(in + off, len, ttype, etype, dataCap, bufCap, tmp, tmpCap, ckBits);
        if (bits < 0) {
            free(tmp);
            return -1;
        }
        const u32 lw = (bits < 8) ? 3u : (u32)ilog2((u32)(bits >> 3)) + 4u;
        bw_put(&w, lw - 3, 5);
        bw_put(&w, (u64)bits, (int)lw);
        bw_put_bytes(&w, tmp, bits);
    }
    bw_put(&w, 0, 5);
    bw_put(&w, 0, 3);
    free(tmp);
    if (w.overflow)
        return -1;
    return (w.bits + 7) >> 3;
}

API i64 ko_stream_compress(const u8* in, i64 n, u64 ttype, int etype, int blockSize, u8* out, i64 cap)
{
    return ko_stream_compress_ck(in, n, ttype, etype, blockSize, 0, out, cap);
}

API u64 ko_block_hash(const u8* in, int n, int ckBits) { return block_hash(in, n, ckBits); }

/* io/CompressedInputStream.cpp:511-663 readHeader (+ block loop); -5 = block checksum mismatch */
API i64 ko_stream_decompress(const u8* in, i64 n, u8* out, i64 cap)
{
    BitR r;
    br_init(&r, in, 8 * n);
    if (br_get(&r, 32) != 0x4B414E5A)
        return -1;
    if (br_get(&r, 4) != 6)
        return -2;
    const int ckBits = 32 * (int)br_get(&r, 2);
    if (ckBits > 64)
        return -3;
    const int etype = (int)br_get(&r, 5);
    const u64 ttype = br_get(&r, 48);
    const int blockSize = (int)br_get(&r, 28) << 4;
    const int szMask = (int)br_get(&r, 2);
    if (szMask)
        br_get(&r, 16 * szMask);
    br_get(&r, 15);
    br_get(&r, 24);
    i64 produced = 0;
    u8* tmp = NULL;
    i64 tmpCap = 0;
    for (;;) {
        const int lr = 3 + (int)br_get(&r, 5);
        const u64 bits = br_get(&r, lr);
        if (bits == 0 || r.underflow)
            break;
        const i64 nb = (i64)((bits + 7) >> 3);
        if (nb > tmpCap) {
            free(tmp);
            tmpCap = nb + 64;
            tmp = (u8*)malloc((size_t)tmpCap);
        }
        memset(tmp, 0, (size_t)nb);
        for (u64 k = 0; k < bits; k++) { /* re-align the block to byte 0 */
            if (br_get(&r, 1))
                tmp[k >> 3] |= (u8)(0x80 >> (k & 7));
        }
        const i64 room = cap - produced;
        const int rc = decode_block_ck(tmp, (i64)bits, ttype, etype, blockSize, out + produced,
            (int)((room < blockSize) ? room : blockSize), ckBits);
        if (rc < 0) {
            free(tmp);
            return (rc == -2) ? -5 : -4;
        }
        produced += rc;
    }
    free(tmp);
    return produced;
}
