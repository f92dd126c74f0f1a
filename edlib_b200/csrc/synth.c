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

/* out: nreads x readLen bytes.  Returns nothing; every read is exactly readLen long. */
void synth_reads(const unsigned char* target, int64_t tlen, unsigned char* out, int64_t nreads, int readLen,
                 double rate, uint64_t seed) {
    const uint64_t thresh = (uint64_t)(rate * 18446744073709551615.0);
    for (int64_t i = 0; i < nreads; ++i) {
        uint64_t s = (seed + 1) * 0x9E3779B97F4A7C15ull ^ ((uint64_t)i * 0xC2B2AE3D27D4EB4Full);
        s = splitmix(&s) ^ (uint64_t)i;          /* decorrelate neighbouring streams */
        int64_t span = tlen - readLen - 16;
        if (span < 1) span = 1;
        int64_t pos = (int64_t)(splitmix(&s) % (uint64_t)span);
        unsigned char* o = out + i * readLen;
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
