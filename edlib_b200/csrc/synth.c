/* synth.c -- seeded synthetic DNA workloads for bench.py and the tests (tooling, not part of
 * the alignment path).  PRNG: splitmix64 stream per sequence, so the data only depends on
 * (seed, index).  Shapes follow SURVEY.md section 8d / BASELINE.json configs:
 *   reads : uniform start in the target, per-base events with probability `rate`
 *           (substitution / 1-bp insertion / 1-bp deletion, one third each), cut at readLen.
 */
#include <stdint.h>

static const unsigned char DNA[4] = {'A', 'C', 'G', 'T'};

static inline uint64_t splitmix(uint64_t* s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

void synth_dna(unsigned char* out, int64_t n, uint64_t seed) {
    uint64_t s = seed * 0xD1342543DE82EF95ull + 1;
    int64_t i = 0;
    while (i < n) {
        uint64_t r = splitmix(&s);
        for (int k = 0; k < 32 && i < n; ++k, r >>= 2) out[i++] = DNA[r & 3];
    }
}

/* Reads [first, last) of the batch keyed by `seed` (read i only depends on (seed, i)): out holds (last - first) x
 * readLen bytes; every read is exactly readLen long. */
void synth_reads_range(const unsigned char* target, int64_t tlen, unsigned char* out, int64_t first, int64_t last, int readLen,
                       double rate, uint64_t seed);

/* out: nreads x readLen bytes: reads [0, nreads) of the batch. */
void synth_reads(const unsigned char* target, int64_t tlen, unsigned char* out, int64_t nreads, int readLen,
                 double rate, uint64_t seed) {
    synth_reads_range(target, tlen, out, 0, nreads, readLen, rate, seed);
}

void synth_reads_range(const unsigned char* target, int64_t tlen, unsigned char* out, int64_t first, int64_t last, int readLen,
                       double rate, uint64_t seed) {
    const uint64_t thresh = (uint64_t)(rate * 18446744073709551615.0);
    for (int64_t i = first; i < last; ++i) {
        uint64_t s = (seed + 1) * 0x9E3779B97F4A7C15ull ^ ((uint64_t)i * 0xC2B2AE3D27D4EB4Full);
        s = splitmix(&s) ^ (uint64_t)i;          /* decorrelate neighbouring streams */
        int64_t span = tlen - readLen - 16;
        if (span < 1) span = 1;
        int64_t pos = (int64_t)(splitmix(&s) % (uint64_t)span);
        unsigned char* o = out + (i - first) * readLen;
        int len = 0;
        while (len < readLen) {
            unsigned char c = target[pos < tlen ? pos : tlen - 1];
            ++pos;
            if (splitmix(&s) < thresh) {
                uint64_t r = splitmix(&s);
                int kind = (int)(r % 3);
                unsigned char x = DNA[(r >> 8) & 3];
                if (kind == 0) {
                    o[len++] = x;                 /* substitution (may re-draw the same base) */
                } else if (kind == 1) {
                    o[len++] = c;                 /* insertion after the base */
                    if (len < readLen) o[len++] = x;
                }                                 /* kind 2: deletion */
            } else {
                o[len++] = c;
            }
        }
    }
}

/* Mutated copy of src[0..n): returns the output length (<= 2n). */
int64_t synth_mutate(const unsigned char* src, int64_t n, unsigned char* out, double rate, uint64_t seed) {
    const uint64_t thresh = (uint64_t)(rate * 18446744073709551615.0);
    uint64_t s = (seed + 7) * 0x9E3779B97F4A7C15ull;
    s = splitmix(&s) ^ seed;                     /* decorrelate neighbouring seeds */
    int64_t len = 0;
    for (int64_t i = 0; i < n; ++i) {
        unsigned char c = src[i];
        if (splitmix(&s) < thresh) {
            uint64_t r = splitmix(&s);
            int kind = (int)(r % 3);
            unsigned char x = DNA[(r >> 8) & 3];
            if (kind == 0) out[len++] = x;
            else if (kind == 1) { out[len++] = c; out[len++] = x; }
        } else {
            out[len++] = c;
        }
    }
    return len;
}

/* config 3 shape, pairs [i0, i1): query i = the `len` symbols of the genome at a seeded start, written to
 * qbuf + i*len; target i = its mutated copy (same per-base events as the reads), written to tbuf + i*tstride with
 * its length in tlens[i] (tstride >= 2*len). */
void synth_pairs(const unsigned char* genome, int64_t glen, int64_t i0, int64_t i1, int len, double rate, uint64_t seed,
                 unsigned char* qbuf, unsigned char* tbuf, int64_t tstride, int* tlens) {
    for (int64_t i = i0; i < i1; ++i) {
        uint64_t s = (seed + 3) * 0x9E3779B97F4A7C15ull ^ ((uint64_t)i * 0xC2B2AE3D27D4EB4Full);
        s = splitmix(&s) ^ (uint64_t)i;
        int64_t span = glen - len - 64;
        if (span < 1) span = 1;
        const int64_t start = (int64_t)(splitmix(&s) % (uint64_t)span);
        unsigned char* q = qbuf + i * (int64_t)len;
        for (int k = 0; k < len; ++k) q[k] = genome[start + k < glen ? start + k : glen - 1];
        tlens[i] = (int)synth_mutate(q, len, tbuf + i * tstride, rate, seed * 1000003ull + (uint64_t)i);
    }
}
