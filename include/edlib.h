/*
 * edlib.h -- C ABI of the B200-native batched edit-distance engine.
 *
 * Drop-in boundary: every type and function below keeps the binary layout and the
 * behaviour of the reference interface it replaces (reference edlib/include/edlib.h,
 * cited per declaration as ref h:LINE), so callers of the reference (aligner.cpp:169,
 * edlib.pyx:129, runTests.cpp, helloWorld.c) re-link unchanged.  The computation behind
 * edlibAlign()/edlibAlignBatch() runs in hand-written sm_100a CUDA kernels; there is no
 * CPU compute path -- without a usable CUDA device the calls return EDLIB_STATUS_ERROR.
 *
 * Layout facts the ABI depends on (x86-64 SysV):
 *   sizeof(EdlibEqualityPair) == 2, sizeof(EdlibAlignConfig) == 32,
 *   sizeof(EdlibAlignResult) == 48; enum values NW=0 SHW=1 HW=2, DISTANCE=0 LOC=1 PATH=2,
 *   STANDARD=0 EXTENDED=1.  Config and result travel BY VALUE.
 *
 * One entry point is new: edlibAlignBatch() (contract at the bottom of this file).
 */
#ifndef EDLIB_H
#define EDLIB_H

#if defined(EDLIB_SHARED) && !defined(_WIN32)
#  define EDLIB_API __attribute__((visibility("default")))
#elif defined(EDLIB_SHARED) && defined(EDLIB_BUILD)
#  define EDLIB_API __declspec(dllexport)
#elif defined(EDLIB_SHARED)
#  define EDLIB_API __declspec(dllimport)
#else
#  define EDLIB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* status field of EdlibAlignResult (ref h:30-31) */
#define EDLIB_STATUS_OK    0
#define EDLIB_STATUS_ERROR 1

/* Where gaps are free (ref h:36-62).
 * NW  : global, nothing free.
 * SHW : prefix, target characters after the query's end are free.
 * HW  : infix, target characters before the start and after the end are free. */
typedef enum { EDLIB_MODE_NW, EDLIB_MODE_SHW, EDLIB_MODE_HW } EdlibAlignMode;

/* How much to compute (ref h:67-71): distance+ends / +starts / +edit script. */
typedef enum { EDLIB_TASK_DISTANCE, EDLIB_TASK_LOC, EDLIB_TASK_PATH } EdlibAlignTask;

/* CIGAR alphabets (ref h:78-81): standard "MID", extended "=XID". */
typedef enum { EDLIB_CIGAR_STANDARD, EDLIB_CIGAR_EXTENDED } EdlibCigarFormat;

/* Edit-script byte codes stored in EdlibAlignResult.alignment (ref h:84-87). */
#define EDLIB_EDOP_MATCH    0 /* query char == target char (under the equality relation) */
#define EDLIB_EDOP_INSERT   1 /* query char with no target char */
#define EDLIB_EDOP_DELETE   2 /* target char with no query char */
#define EDLIB_EDOP_MISMATCH 3 /* substitution */

/* Declares `first` and `second` interchangeable; symmetric, not transitive (ref h:92-95). */
typedef struct {
    char first;
    char second;
} EdlibEqualityPair;

/* Call options, passed by value (ref h:100-140). */
typedef struct {
    int k;                 /* >=0: report -1 if distance exceeds k.  <0: no bound. */
    EdlibAlignMode mode;
    EdlibAlignTask task;
    const EdlibEqualityPair* additionalEqualities; /* may be NULL */
    int additionalEqualitiesLength;
} EdlibAlignConfig;

/* ref h:146-150 */
EDLIB_API EdlibAlignConfig edlibNewAlignConfig(int k, EdlibAlignMode mode, EdlibAlignTask task,
                                               const EdlibEqualityPair* additionalEqualities,
                                               int additionalEqualitiesLength);

/* k=-1, NW, DISTANCE, no extra equalities (ref h:156). */
EDLIB_API EdlibAlignConfig edlibDefaultAlignConfig(void);

/* Result, returned by value (ref h:162-218).  The three arrays are libc-malloc memory owned
 * by the caller: release with edlibFreeAlignResult() or plain free(). */
typedef struct {
    int status;             /* EDLIB_STATUS_*; on ERROR nothing else is meaningful */
    int editDistance;       /* -1 when a bound k>=0 was given and is exceeded */
    int* endLocations;      /* 0-based target indices where optimal alignments end; NULL if -1 */
    int* startLocations;    /* matching starts; NULL unless task>=LOC and a solution exists */
    int numLocations;
    unsigned char* alignment; /* EDLIB_EDOP_* codes for the first (start,end); NULL unless PATH */
    int alignmentLength;
    int alphabetLength;     /* distinct byte values in query and target together */
} EdlibAlignResult;

/* free() the three arrays of one result (ref h:224). */
EDLIB_API void edlibFreeAlignResult(EdlibAlignResult result);

/* Levenshtein alignment of one (query,target) pair (ref h:242-246).  Sequences are
 * length-delimited raw bytes (any value 0..255).  Thread-safe. */
EDLIB_API EdlibAlignResult edlibAlign(const char* query, int queryLength,
                                      const char* target, int targetLength,
                                      const EdlibAlignConfig config);

/* Run-length encode an edit script; NUL-terminated malloc'd string, NULL for an unknown
 * format or an op code > 3 (ref h:268-271). */
EDLIB_API char* edlibAlignmentToCigar(const unsigned char* alignment, int alignmentLength,
                                      EdlibCigarFormat cigarFormat);

/* NEW -- batched entry.  results[i] is field-for-field what
 *   edlibAlign(queries[i], queryLengths[i], targets[i], targetLengths[i], config)
 * returns, each with individually free()-able arrays.  Pairs whose targets[i] pointer and
 * length are identical share one upload / one encoding of that target (the reference's own
 * batch shape: aligner.cpp:142,162-170 loops all queries over one target).
 * Returns EDLIB_STATUS_OK, or EDLIB_STATUS_ERROR if the device path failed (then every
 * results[i].status is EDLIB_STATUS_ERROR and no arrays are allocated). */
EDLIB_API int edlibAlignBatch(const char* const* queries, const int* queryLengths,
                              const char* const* targets, const int* targetLengths,
                              int numPairs, const EdlibAlignConfig config,
                              EdlibAlignResult* results);

#ifdef __cplusplus
}
#endif
#endif /* EDLIB_H */
