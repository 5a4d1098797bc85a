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
