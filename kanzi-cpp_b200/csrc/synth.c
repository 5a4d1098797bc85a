// Deterministic synthetic inputs for the kanzi block-pipeline benchmarks
// (SURVEY.md §8(d)).  Host-only helper (libknzsynth.so); shared by bench.py,
// the tests and the golden-fixture generator so every measurement names the
// same bytes.  Not part of the hot path.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint64_t s; } rng_t;

static inline uint64_t next64(rng_t* r)
{
    uint64_t z = (r->s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

#define VOCAB 32768
static const char LETTERS[] = "etaoinshrdlcumwfgypbvkjxqz";

// Zipf-like words over a skewed 26-letter alphabet, ' ' separated, '\n' every 16 words.
void knz_synth_text(uint8_t* out, int64_t n, uint64_t seed)
{
    if (n <= 0)
        return;

    rng_t rv = { seed ^ 0x5EED0001ULL };
    uint8_t* vb = (uint8_t*)malloc((size_t)VOCAB * 12);
    int* voff = (int*)malloc(sizeof(int) * (VOCAB + 1));
    int tot = 0;

    for (int w = 0; w < VOCAB; w++) {
        const int len = 2 + (int)(next64(&rv) % 9);
        voff[w] = tot;

        for (int k = 0; k < len; k++) {
            const uint64_t x = next64(&rv) >> 48; // 16 bits
            vb[tot++] = (uint8_t)LETTERS[(x * x * 26) >> 32];
        }
    }

    voff[VOCAB] = tot;
    rng_t rw = { seed ^ 0x7E47ULL };
    int64_t pos = 0;
    uint64_t wcount = 0;

    while (pos < n) {
        const uint64_t x = next64(&rw) >> 48;
        const int idx = (int)((x * x * x * VOCAB) >> 48);
        const int len = voff[idx + 1] - voff[idx];

        for (int k = 0; k < len && pos < n; k++)
            out[pos++] = vb[voff[idx] + k];

        if (pos < n)
            out[pos++] = ((wcount & 15) == 15) ? (uint8_t)'\n' : (uint8_t)' ';

        wcount++;
    }

    free(vb);
    free(voff);
}

// 32-byte records: LE u32 counter, two slowly varying LE u16, 8 bytes from a
// 16-symbol alphabet, 16 padding bytes (0x00, or 0xFF for 1 record in 4).
static void records(uint8_t* out, int64_t n, uint64_t seed)
{
    rng_t r = { seed ^ 0x4EC0ULL };
    static const char SYM[] = "0123456789ABCDEF";
    uint8_t rec[32];

    for (int64_t i = 0, k = 0; i < n; i += 32, k++) {
        const uint64_t x = next64(&r);
        const uint32_t ctr = (uint32_t)(k + (seed << 8));
        const uint16_t a = (uint16_t)((k >> 6) * 3 + seed);
        const uint16_t b = (uint16_t)((k >> 9) * 7 + 11);
        memcpy(rec, &ctr, 4);
        memcpy(rec + 4, &a, 2);
        memcpy(rec + 6, &b, 2);

        for (int j = 0; j < 8; j++)
            rec[8 + j] = (uint8_t)SYM[(x >> (4 * j)) & 15];

        memset(rec + 16, (((x >> 40) & 3) == 0) ? 0xFF : 0x00, 16);
        const int64_t take = (n - i < 32) ? (n - i) : 32;
        memcpy(out + i, rec, (size_t)take);
    }
}

static void walk(uint8_t* out, int64_t n, uint64_t seed)
{
    rng_t r = { seed ^ 0x3A1CULL };
    uint8_t b = 128;

    for (int64_t i = 0; i < n; i++) {
        b = (uint8_t)(b + (int)((next64(&r) >> 33) % 7) - 3);
        out[i] = b;
    }
}

// 50 % text, 25 % structured records, 25 % byte random walk, segment lengths
// uniform in [64 KiB, 1 MiB); then every 64 KiB a 4 KiB copy from
// (1 MiB - 8 KiB) earlier (bounded-LCP repeats; sources are never targets).
void knz_synth_compressible(uint8_t* out, int64_t n, uint64_t seed)
{
    rng_t rs = { seed ^ 0x5E60ULL };
    int64_t pos = 0;
    uint64_t seg = 0;

    while (pos < n) {
        const uint64_t r = next64(&rs);
        const int64_t length = 65536 + (int64_t)((r >> 8) % (1048576 - 65536));
        const int kind = (int)(r & 3);
        const int64_t take = (length < n - pos) ? length : (n - pos);

        if (kind <= 1)
            knz_synth_text(out + pos, take, seed * 1315423911ULL + seg);
        else if (kind == 2)
            records(out + pos, take, seed + seg);
        else
            walk(out + pos, take, seed + seg);

        pos += take;
        seg++;
    }

    for (int64_t p = 1 << 20; p + 4096 <= n; p += 65536)
        memcpy(out + p, out + p - (1 << 20) + 8192, 4096);
}

void knz_synth_incompressible(uint8_t* out, int64_t n, uint64_t seed)
{
    rng_t r = { seed ^ 0x1C0ULL };

    for (int64_t i = 0; i < n; i++)
        out[i] = (uint8_t)(next64(&r) >> 56);
}

// ---- silesia-shaped mix (SURVEY 8 d, config 3): 45 % text / mark-up, 25 % executable-like, 12 % database
// records, 18 % smooth 16-bit little-endian samples, in segments of [256 KiB, 2 MiB).
static void markup(uint8_t* out, int64_t n, uint64_t seed)
{
    // synth_text words, sentences with a capital and a full stop, wrapped in tags now and then
    uint8_t* words = (uint8_t*)malloc((size_t)n + 16);
    knz_synth_text(words, n, seed);
    rng_t r = { seed ^ 0x7A65ULL };
    static const char* TAGS[] = { "p", "title", "item", "row", "name", "div" };
    int64_t i = 0, o = 0;
    int startSentence = 1;

    while (o < n) {
        // next word of the plain text
        int64_t e = i;
        while (e < n && words[e] != ' ' && words[e] != '\n')
            e++;
        if (e == i) { // separator
            if (i >= n)
                i = 0;
            else
                i++;
            continue;
        }
        const uint64_t x = next64(&r);
        const char* tag = ((x & 31) == 0) ? TAGS[(x >> 8) % 6] : NULL;
        if (tag) {
            out[o++] = '<';
            for (const char* t = tag; *t && o < n; t++)
                out[o++] = (uint8_t)*t;
            if (o < n)
                out[o++] = '>';
        }
        for (int64_t k = i; k < e && o < n; k++) {
            uint8_t c = words[k];
            if (k == i && startSentence)
                c = (uint8_t)(c - 32);
            out[o++] = c;
        }
        startSentence = 0;
        if (tag && o < n) {
            out[o++] = '<';
            if (o < n)
                out[o++] = '/';
            for (const char* t = tag; *t && o < n; t++)
                out[o++] = (uint8_t)*t;
            if (o < n)
                out[o++] = '>';
        }
        if (o < n) {
            const int k = (int)((x >> 16) & 31);
            if (k == 0) {
                out[o++] = '.';
                startSentence = 1;
                if (o < n)
                    out[o++] = ((x >> 24) & 3) ? ' ' : '\n';
            } else if (k == 1 && o + 5 < n) {
                memcpy(out + o, "&amp;", 5);
                o += 5;
            } else {
                out[o++] = ' ';
            }
        }
        i = e;
    }
    free(words);
}

static void exe_like(uint8_t* out, int64_t n, uint64_t seed)
{
    // opcode bytes from a skewed table, every few instructions a call / jump with an increasing LE32 address
    rng_t r = { seed ^ 0xE8E9ULL };
    static const uint8_t OPS[16] = { 0x8B, 0x89, 0x48, 0x83, 0xE8, 0xFF, 0x0F, 0x85, 0x74, 0xC3, 0x55, 0x5D, 0x31, 0xC0, 0x4C, 0x8D };
    uint32_t addr = 0x00401000u + (uint32_t)(seed << 12);
    int64_t o = 0;

    while (o < n) {
        const uint64_t x = next64(&r);
        const int len = 1 + (int)(x & 3);
        for (int k = 0; k < len && o < n; k++) {
            const uint64_t y = (x >> (8 + 8 * k)) & 0xFF;
            out[o++] = OPS[(y * y) >> 12];
        }
        if (((x >> 48) & 7) == 0 && o + 5 <= n) {
            out[o++] = 0xE8;
            addr += (uint32_t)((x >> 52) & 0xFF) * 4;
            memcpy(out + o, &addr, 4);
            o += 4;
        } else if (o < n) {
            out[o++] = (uint8_t)(x >> 40); // modrm / immediate
        }
    }
}

static void samples16(uint8_t* out, int64_t n, uint64_t seed)
{
    rng_t r = { seed ^ 0x16B17ULL };
    int32_t v = 12000, slope = 0;
    for (int64_t i = 0; i + 1 < n; i += 2) {
        const uint64_t x = next64(&r);
        slope += (int32_t)(x % 9) - 4;
        if (slope > 60)
            slope = 60;
        if (slope < -60)
            slope = -60;
        v += slope + (int32_t)((x >> 16) % 5) - 2;
        const uint16_t u = (uint16_t)v;
        out[i] = (uint8_t)u;
        out[i + 1] = (uint8_t)(u >> 8);
    }
    if (n & 1)
        out[n - 1] = 0;
}

void knz_synth_silesia(uint8_t* out, int64_t n, uint64_t seed)
{
    rng_t rs = { seed ^ 0x511E51AULL };
    int64_t pos = 0;
    uint64_t seg = 0;

    while (pos < n) {
        const uint64_t r = next64(&rs);
        const int64_t length = 262144 + (int64_t)((r >> 8) % (2097152 - 262144));
        const int64_t take = (length < n - pos) ? length : (n - pos);
        const int pct = (int)(r % 100);

        if (pct < 45)
            markup(out + pos, take, seed * 2654435761ULL + seg);
        else if (pct < 70)
            exe_like(out + pos, take, seed + seg);
        else if (pct < 82)
            records(out + pos, take, seed + seg);
        else
            samples16(out + pos, take, seed + seg);

        pos += take;
        seg++;
    }
}
