/*
 * TEST INFRASTRUCTURE -- NOT PART OF THE PRODUCT (see edlib_oracle.h).
 *
 * Plain-C restatement of the reference algorithm behind edlibAlign(), reference
 * edlib/src/edlib.cpp (v1.2.6).  It restates WHAT the reference computes, with the
 * reference's speed devices (Ukkonen band, k doubling, padded last block) removed:
 *
 *   - distances and end locations come from an UNBANDED Myers bit-vector sweep over
 *     64-bit blocks (the block update restates calculateBlock, ref cpp:412-447; the sweeps
 *     restate myersCalcEditDistanceSemiGlobal cpp:550-704 and myersCalcEditDistanceNW
 *     cpp:730-928).  The band of the reference only prunes cells whose value exceeds k, so an
 *     unbanded sweep followed by the "<= k" test gives the same answer;
 *   - the reference reads the score of the last query row through W = 64*ceil(m/64)-m
 *     wildcard padding rows, which shifts end positions by W and is the origin of its "-1"
 *     end location (cpp:670, 681-693).  Here the last row is read directly and the -1 entry
 *     is restated as an explicit rule (see semi_global());
 *   - start locations restate cpp:228-272, the path restates obtainAlignment cpp:1161-1213,
 *     obtainAlignmentTraceback cpp:942-1141 (as a plain DP traceback with the same move
 *     priority) and obtainAlignmentHirschberg cpp:1231-1396 (same split column, same
 *     candidate order, same recursion), using cell-by-cell DP columns.
 *
 * Every rule above is pinned against the unmodified reference build (oracle/_ref) by
 * tests/test_oracle.py.
 */
#include "edlib_oracle.h"

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint64_t word_t;
#define WBITS 64

/* ------------------------------------------------------------------------------------ */
/* Alphabet and equality relation                                                        */
/* ------------------------------------------------------------------------------------ */

typedef struct {
    int size;                 /* number of distinct byte values seen: result.alphabetLength */
    unsigned char code[256];  /* byte value -> dense code, in order of first appearance     */
    unsigned char seen[256];
    unsigned char* eq;        /* size x size, eq[a*size+b] != 0 when codes a and b match     */
} Alphabet;

/* ref cpp:1417-1462 (transformSequences): codes are handed out in order of first
 * appearance, scanning the query first and the target second. */
static void alphabet_scan(Alphabet* A, const char* q, int m, const char* t, int n,
                          unsigned char* qc, unsigned char* tc) {
    memset(A, 0, sizeof(*A));
    for (int pass = 0; pass < 2; pass++) {
        const unsigned char* s = (const unsigned char*)(pass ? t : q);
        unsigned char* out = pass ? tc : qc;
        int len = pass ? n : m;
        for (int i = 0; i < len; i++) {
            unsigned char b = s[i];
            if (!A->seen[b]) {
                A->seen[b] = 1;
                A->code[b] = (unsigned char)A->size++;
            }
            out[i] = A->code[b];
        }
    }
}

/* ref cpp:63-94 (EqualityDefinition): identity, plus each extra pair made symmetric; a pair
 * naming a byte that occurs in neither sequence is ignored. */
static void alphabet_equalities(Alphabet* A, const EdlibEqualityPair* pairs, int numPairs) {
    int s = A->size;
    A->eq = (unsigned char*)calloc((size_t)(s > 0 ? s * s : 1), 1);
    for (int i = 0; i < s; i++) A->eq[i * s + i] = 1;
    if (!pairs) return;
    for (int i = 0; i < numPairs; i++) {
        unsigned char a = (unsigned char)pairs[i].first, b = (unsigned char)pairs[i].second;
        if (A->seen[a] && A->seen[b]) {
            int ca = A->code[a], cb = A->code[b];
            A->eq[ca * s + cb] = A->eq[cb * s + ca] = 1;
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* Myers bit-vector machinery                                                            */
/* ------------------------------------------------------------------------------------ */

static int ceil_div(int a, int b) { return (a + b - 1) / b; } /* ref cpp:453-455 */

/* Query profile, ref cpp:358-384 (buildPeq) without the wildcard padding: bit i of
 * peq[c*nb + b] is set when query row b*64+i matches code c; rows >= m stay 0. */
static word_t* build_profile(const Alphabet* A, const unsigned char* qc, int m) {
    int nb = ceil_div(m, WBITS);
    word_t* peq = (word_t*)calloc((size_t)A->size * nb + 1, sizeof(word_t));
    for (int r = 0; r < m; r++)
        for (int c = 0; c < A->size; c++)
            if (A->eq[qc[r] * A->size + c]) peq[c * nb + r / WBITS] |= (word_t)1 << (r % WBITS);
    return peq;
}

/* One 64-row x 1-column update, ref cpp:412-447 (calculateBlock).  pv/mv are the vertical
 * +1/-1 delta sets of the block, eq the match set for this column's target symbol, hin the
 * horizontal delta entering above the block.  Outputs the UNSHIFTED horizontal delta sets
 * (bit i = delta of row i) so a caller can read any row, and returns the delta of row 63. */
static int block_step(word_t* pv, word_t* mv, word_t eq, int hin, word_t* phOut, word_t* mhOut) {
    const word_t P = *pv, M = *mv;
    const word_t inNeg = hin < 0, inPos = hin > 0;
    const word_t xv = eq | M;
    eq |= inNeg;
    const word_t xh = (((eq & P) + P) ^ P) | eq;
    const word_t ph = M | ~(xh | P);
    const word_t mh = P & xh;
    const word_t phs = (ph << 1) | inPos;
    const word_t mhs = (mh << 1) | inNeg;
    *pv = mhs | ~(xv | phs);
    *mv = phs & xv;
    *phOut = ph;
    *mhOut = mh;
    return (int)(ph >> (WBITS - 1)) - (int)(mh >> (WBITS - 1));
}

typedef struct {
    int* v;
    int n, cap;
} IntVec;

static void vec_push(IntVec* x, int val) {
    if (x->n == x->cap) {
        x->cap = x->cap ? x->cap * 2 : 16;
        x->v = (int*)realloc(x->v, sizeof(int) * (size_t)x->cap);
    }
    x->v[x->n++] = val;
}

/* Unbanded column sweep.  Column -1 is D[r][-1] = r+1 (all vertical deltas +1, ref
 * cpp:575-579 / 759-763); the horizontal delta entering the top row is topHin (0 for HW,
 * +1 for SHW/NW: ref cpp:584,590 and 779).  For every column c the value of the LAST query
 * row, D[m-1][c], is handed to `visit`.  tstep = +1 walks the target forward from tc[0],
 * tstep = -1 walks it backward from tc[0] (used for the reversed sweeps of cpp:253-257). */
typedef void (*ColumnVisitor)(void* ctx, int column, int lastRowScore);

static void sweep(const word_t* peq, int nb, int m, const unsigned char* tc, int n, int tstep,
                  int topHin, ColumnVisitor visit, void* ctx) {
    word_t* P = (word_t*)malloc(sizeof(word_t) * (size_t)nb);
    word_t* M = (word_t*)malloc(sizeof(word_t) * (size_t)nb);
    for (int b = 0; b < nb; b++) { P[b] = ~(word_t)0; M[b] = 0; }
    const int lastBit = (m - 1) % WBITS;
    int score = m;
    for (int c = 0; c < n; c++) {
        const word_t* col = peq + (size_t)tc[(long)c * tstep] * nb;
        int h = topHin;
        word_t ph = 0, mh = 0;
        for (int b = 0; b < nb; b++) h = block_step(&P[b], &M[b], col[b], h, &ph, &mh);
        score += (int)((ph >> lastBit) & 1) - (int)((mh >> lastBit) & 1);
        visit(ctx, c, score);
    }
    free(P);
    free(M);
}

/* Collects min over columns of D[m-1][c] and every column attaining it, restating the
 * best-score bookkeeping of ref cpp:658-673: a strictly better score clears the list. */
typedef struct {
    int best;
    IntVec pos;
} BestTracker;

static void track_best(void* ctx, int column, int s) {
    BestTracker* b = (BestTracker*)ctx;
    if (b->best < 0 || s < b->best) { b->best = s; b->pos.n = 0; }
    if (s == b->best) vec_push(&b->pos, column);
}

typedef struct { int last; } LastTracker;
static void track_last(void* ctx, int column, int s) { (void)column; ((LastTracker*)ctx)->last = s; }

/* HW / SHW distance and end locations, ref cpp:550-704.
 * Result rule (band and padding removed):
 *   best = min_c D[m-1][c]; not found (-1) when k >= 0 and best > k;
 *   end locations = every c with D[m-1][c] == best, ascending, and additionally a leading
 *   -1 when best == m and W = 64*ceil(m/64)-m > 0.  That -1 is the reference's padded
 *   bottom cell of column W-1 (value m, reported as position (W-1)-W, cpp:670 / tail loop
 *   cpp:681-693); with W == 0 there is no padded row and no -1. */
static int semi_global(const word_t* peq, int nb, int m, const unsigned char* tc, int n, int tstep,
                       int k, int isHW, int** positions, int* numPositions) {
    BestTracker bt = { -1, { NULL, 0, 0 } };
    sweep(peq, nb, m, tc, n, tstep, isHW ? 0 : 1, track_best, &bt);
    *positions = NULL;
    *numPositions = 0;
    if (bt.best < 0 || (k >= 0 && bt.best > k)) { free(bt.pos.v); return -1; }
    int W = nb * WBITS - m;
    int lead = (bt.best == m && W > 0) ? 1 : 0;
    int* out = (int*)malloc(sizeof(int) * (size_t)(bt.pos.n + lead));
    if (lead) out[0] = -1;
    memcpy(out + lead, bt.pos.v, sizeof(int) * (size_t)bt.pos.n);
    *positions = out;
    *numPositions = bt.pos.n + lead;
    free(bt.pos.v);
    return bt.best;
}

/* NW distance, ref cpp:730-928: rejected outright when k < |n-m| (cpp:744), otherwise the
 * bottom-right cell, accepted when <= k (cpp:916-917). */
static int global_distance(const word_t* peq, int nb, int m, const unsigned char* tc, int n, int k) {
    if (k >= 0 && k < abs(n - m)) return -1;
    LastTracker lt = { -1 };
    sweep(peq, nb, m, tc, n, 1, 1, track_last, &lt);
    if (k >= 0 && lt.last > k) return -1;
    return lt.last;
}

/* ------------------------------------------------------------------------------------ */
/* Alignment path                                                                        */
/* ------------------------------------------------------------------------------------ */

static int obtain_alignment(const unsigned char* q, int m, const unsigned char* t, int n,
                            const Alphabet* A, int best, unsigned char** aln, int* alnLen);

/* One NW DP column set: out[i] = distance(q[0..i), t[0..n)), i = 0..m.  qstep/tstep = -1
 * read the sequences backward from q[0] / t[0] (the reversed halves of cpp:1258-1260). */
static void nw_last_column(const unsigned char* q, int qstep, int m, const unsigned char* t, int tstep,
                           int n, const Alphabet* A, int* out) {
    for (int i = 0; i <= m; i++) out[i] = i;
    for (int j = 1; j <= n; j++) {
        int diag = out[0];
        out[0] = j;
        int tcj = t[(long)(j - 1) * tstep];
        for (int i = 1; i <= m; i++) {
            int up = out[i - 1], left = out[i];
            int cand = diag + (A->eq[q[(long)(i - 1) * qstep] * A->size + tcj] ? 0 : 1);
            if (up + 1 < cand) cand = up + 1;
            if (left + 1 < cand) cand = left + 1;
            diag = left;
            out[i] = cand;
        }
    }
}

/* Restates obtainAlignmentTraceback (ref cpp:942-1141) on a fully materialised NW matrix:
 * walk from the bottom-right cell; prefer UP (INSERT, cpp:1020), then LEFT (DELETE,
 * cpp:1054), then the diagonal (MATCH when the score does not change, else MISMATCH,
 * cpp:1085-1086); when a matrix edge is reached run out along it (cpp:1025-1029,
 * 1059-1065, 1090-1103); the collected ops are reversed at the end (cpp:1139). */
static void traceback_full(const unsigned char* q, int m, const unsigned char* t, int n,
                           const Alphabet* A, unsigned char** aln, int* alnLen) {
    size_t cols = (size_t)n + 1;
    int* D = (int*)malloc(sizeof(int) * (size_t)(m + 1) * cols);
    for (int j = 0; j <= n; j++) D[j] = j;
    for (int i = 1; i <= m; i++) {
        D[i * cols] = i;
        for (int j = 1; j <= n; j++) {
            int cand = D[(i - 1) * cols + j - 1] + (A->eq[q[i - 1] * A->size + t[j - 1]] ? 0 : 1);
            int up = D[(i - 1) * cols + j] + 1, left = D[i * cols + j - 1] + 1;
            if (up < cand) cand = up;
            if (left < cand) cand = left;
            D[i * cols + j] = cand;
        }
    }
    unsigned char* ops = (unsigned char*)malloc((size_t)m + n + 1);
    int len = 0, i = m, j = n;
    for (;;) {
        int cur = D[i * cols + j];
        if (D[(i - 1) * cols + j] + 1 == cur) {
            ops[len++] = EDLIB_EDOP_INSERT;
            if (--i == 0) { while (j-- > 0) ops[len++] = EDLIB_EDOP_DELETE; break; }
        } else if (D[i * cols + j - 1] + 1 == cur) {
            ops[len++] = EDLIB_EDOP_DELETE;
            if (--j == 0) { while (i-- > 0) ops[len++] = EDLIB_EDOP_INSERT; break; }
        } else {
            ops[len++] = D[(i - 1) * cols + j - 1] == cur ? EDLIB_EDOP_MATCH : EDLIB_EDOP_MISMATCH;
            --i; --j;
            if (j == 0) { while (i-- > 0) ops[len++] = EDLIB_EDOP_INSERT; break; }
            if (i == 0) { while (j-- > 0) ops[len++] = EDLIB_EDOP_DELETE; break; }
        }
    }
    for (int a = 0, b = len - 1; a < b; a++, b--) { unsigned char x = ops[a]; ops[a] = ops[b]; ops[b] = x; }
    free(D);
    *aln = ops;
    *alnLen = len;
}

/* Restates obtainAlignmentHirschberg (ref cpp:1231-1396).  Split the target at n/2
 * (cpp:1247-1248).  With L[h] = distance(q[0..h), left half) and R[h] = distance(q[h..m),
 * right half), the split height h is the first of
 *     h = 1 .. m-1   (cpp:1327-1335: rows queryIdx = h-1, ascending),
 *     h = 0          (top boundary, cpp:1337-1344),
 *     h = m          (bottom boundary, cpp:1345-1353)
 * with L[h] + R[h] == best.  Both quadrants are solved recursively through
 * obtain_alignment() with L[h] and R[h] as their known scores (cpp:1372-1380). */
static int hirschberg(const unsigned char* q, int m, const unsigned char* t, int n,
                      const Alphabet* A, int best, unsigned char** aln, int* alnLen) {
    const int leftW = n / 2, rightW = n - leftW;
    int* L = (int*)malloc(sizeof(int) * (size_t)(m + 1));
    int* Rr = (int*)malloc(sizeof(int) * (size_t)(m + 1));
    nw_last_column(q, 1, m, t, 1, leftW, A, L);
    nw_last_column(q + m - 1, -1, m, t + n - 1, -1, rightW, A, Rr); /* Rr[s] : suffix of length s */
    int h = -1;
    for (int cand = 1; cand <= m - 1 && h < 0; cand++)
        if (L[cand] + Rr[m - cand] == best) h = cand;
    if (h < 0 && L[0] + Rr[m] == best) h = 0;
    if (h < 0 && L[m] + Rr[0] == best) h = m;
    int status = EDLIB_STATUS_ERROR;
    if (h >= 0) {
        unsigned char *ul = NULL, *lr = NULL;
        int ulLen = 0, lrLen = 0;
        int s1 = obtain_alignment(q, h, t, leftW, A, L[h], &ul, &ulLen);
        int s2 = obtain_alignment(q + h, m - h, t + leftW, rightW, A, Rr[m - h], &lr, &lrLen);
        if (s1 == EDLIB_STATUS_OK && s2 == EDLIB_STATUS_OK) {
            *alnLen = ulLen + lrLen;
            *aln = (unsigned char*)malloc((size_t)*alnLen + 1);
            memcpy(*aln, ul, (size_t)ulLen);
            memcpy(*aln + ulLen, lr, (size_t)lrLen);
            status = EDLIB_STATUS_OK;
        }
        free(ul);
        free(lr);
    }
    free(L);
    free(Rr);
    return status;
}

/* Restates obtainAlignment (ref cpp:1161-1213): an empty side gives a run of one op
 * (cpp:1168-1175); the stored-matrix traceback is used while its estimated size
 * (2*8+4)*ceil(m/64)*n + 2*4*n stays below 1 MiB, otherwise Hirschberg (cpp:1188-1211). */
static int obtain_alignment(const unsigned char* q, int m, const unsigned char* t, int n,
                            const Alphabet* A, int best, unsigned char** aln, int* alnLen) {
    if (m == 0 || n == 0) {
        *alnLen = m + n;
        *aln = (unsigned char*)malloc((size_t)*alnLen + 1);
        memset(*aln, m == 0 ? EDLIB_EDOP_DELETE : EDLIB_EDOP_INSERT, (size_t)*alnLen);
        return EDLIB_STATUS_OK;
    }
    long long matrixBytes = 20LL * ceil_div(m, WBITS) * n + 8LL * n;
    if (matrixBytes < 1024 * 1024) {
        traceback_full(q, m, t, n, A, aln, alnLen);
        return EDLIB_STATUS_OK;
    }
    return hirschberg(q, m, t, n, A, best, aln, alnLen);
}

/* ------------------------------------------------------------------------------------ */
/* Public entry points                                                                   */
/* ------------------------------------------------------------------------------------ */

EdlibAlignResult oracleAlign(const char* query, int m, const char* target, int n,
                             const EdlibAlignConfig config) {
    EdlibAlignResult res;
    memset(&res, 0, sizeof(res));
    res.status = EDLIB_STATUS_OK;
    res.editDistance = -1;

    unsigned char* qc = (unsigned char*)malloc((size_t)m + 1);
    unsigned char* tc = (unsigned char*)malloc((size_t)n + 1);
    Alphabet A;
    alphabet_scan(&A, query, m, target, n, qc, tc);
    res.alphabetLength = A.size; /* ref cpp:162: set before anything can fail */

    /* ref cpp:166-184: an empty sequence short-circuits everything, including LOC/PATH. */
    if (m == 0 || n == 0) {
        if (config.mode == EDLIB_MODE_NW || config.mode == EDLIB_MODE_SHW || config.mode == EDLIB_MODE_HW) {
            res.editDistance = config.mode == EDLIB_MODE_NW ? (m > n ? m : n) : m;
            res.endLocations = (int*)malloc(sizeof(int));
            res.endLocations[0] = config.mode == EDLIB_MODE_NW ? n - 1 : -1;
            res.numLocations = 1;
        } else {
            res.status = EDLIB_STATUS_ERROR;
        }
        free(qc);
        free(tc);
        return res;
    }

    alphabet_equalities(&A, config.additionalEqualities, config.additionalEqualitiesLength);
    const int nb = ceil_div(m, WBITS);
    word_t* peq = build_profile(&A, qc, m);
    const int isHW = config.mode == EDLIB_MODE_HW, isSHW = config.mode == EDLIB_MODE_SHW;

    /* ref cpp:199-217: a negative k means "no bound" (the doubling loop always terminates
     * with the true distance); any other mode value falls through to NW (cpp:210). */
    if (isHW || isSHW) {
        res.editDistance = semi_global(peq, nb, m, tc, n, 1, config.k, isHW, &res.endLocations, &res.numLocations);
    } else {
        res.editDistance = global_distance(peq, nb, m, tc, n, config.k);
        if (res.editDistance >= 0) { /* ref cpp:221-225 */
            res.endLocations = (int*)malloc(sizeof(int));
            res.endLocations[0] = n - 1;
            res.numLocations = 1;
        }
    }

    if (res.editDistance >= 0 && (config.task == EDLIB_TASK_LOC || config.task == EDLIB_TASK_PATH)) {
        res.startLocations = (int*)malloc(sizeof(int) * (size_t)res.numLocations);
        if (isHW) {
            /* ref cpp:230-266: reversed query against the reversed target prefix ending at
             * each end location, SHW with k = editDistance; the LAST best end of that sweep is
             * the earliest start.  Only m+editDistance columns can hold a score <= k. */
            unsigned char* rq = (unsigned char*)malloc((size_t)m);
            for (int i = 0; i < m; i++) rq[i] = qc[m - 1 - i];
            word_t* rpeq = build_profile(&A, rq, m);
            for (int i = 0; i < res.numLocations; i++) {
                int e = res.endLocations[i];
                if (e < 0) { res.startLocations[i] = 0; continue; } /* cpp:237-249 */
                int span = e + 1;
                if (span > m + res.editDistance) span = m + res.editDistance;
                int* pos = NULL;
                int npos = 0;
                semi_global(rpeq, nb, m, tc + e, span, -1, res.editDistance, 0, &pos, &npos);
                res.startLocations[i] = e - pos[npos - 1];
                free(pos);
            }
            free(rpeq);
            free(rq);
        } else {
            for (int i = 0; i < res.numLocations; i++) res.startLocations[i] = 0; /* cpp:267-271 */
        }
    }

    if (res.editDistance >= 0 && config.task == EDLIB_TASK_PATH) {
        /* ref cpp:276-289: path only for the first (start,end) pair; status is ignored. */
        int s0 = res.startLocations[0], e0 = res.endLocations[0];
        obtain_alignment(qc, m, tc + s0, e0 - s0 + 1, &A, res.editDistance, &res.alignment, &res.alignmentLength);
    }

    free(peq);
    free(A.eq);
    free(qc);
    free(tc);
    return res;
}

/* ref cpp:303-350: run-length encoding; standard format folds match and mismatch into 'M'. */
char* oracleAlignmentToCigar(const unsigned char* alignment, int alignmentLength, EdlibCigarFormat fmt) {
    if (fmt != EDLIB_CIGAR_EXTENDED && fmt != EDLIB_CIGAR_STANDARD) return NULL;
    const char* letters = fmt == EDLIB_CIGAR_EXTENDED ? "=IDX" : "MIDM";
    for (int i = 0; i < alignmentLength; i++)
        if (alignment[i] > 3) return NULL;
    size_t cap = (size_t)alignmentLength * 2 + 2, len = 0;
    char* out = (char*)malloc(cap);
    int i = 0;
    while (i < alignmentLength) {
        char c = letters[alignment[i]];
        int run = 0;
        while (i < alignmentLength && letters[alignment[i]] == c) { run++; i++; }
        char digits[16];
        int nd = 0;
        for (int r = run; r; r /= 10) digits[nd++] = (char)('0' + r % 10);
        while (nd) out[len++] = digits[--nd];
        out[len++] = c;
    }
    out[len] = 0;
    return out;
}

void oracleFreeAlignResult(EdlibAlignResult r) {
    free(r.endLocations);
    free(r.startLocations);
    free(r.alignment);
}
