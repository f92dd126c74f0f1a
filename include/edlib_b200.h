/*
 * edlib_b200.h -- engine-specific additions to the edlib C ABI (plain C, no CUDA or torch
 * types).  Nothing here exists in the reference; the reference-facing surface is edlib.h.
 *
 * The staged calls split edlibAlignBatch() into its three phases so that a caller (bench.py,
 * a multi-GPU driver running one process per device) can keep a batch resident in HBM and
 * time the device work alone:
 *
 *     b = edlibB200BatchPrepare(...)    pack + upload + alphabet/encoding kernels
 *     edlibB200BatchCompute(b, &st)     every DP kernel; may be repeated on the same batch
 *     edlibB200BatchResults(b, res)     malloc'd EdlibAlignResult per pair (as edlibAlignBatch)
 *     edlibB200BatchFree(b)
 *
 * All calls use the CUDA device that is current for the calling thread at the first call
 * (cudaSetDevice / torch.cuda.set_device before it); one process drives one GPU.
 */
#ifndef EDLIB_B200_H
#define EDLIB_B200_H

#include "edlib.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct EdlibB200Batch EdlibB200Batch; /* opaque */

typedef struct {
    double kernelMs;      /* device time of all kernel launches of the last compute (CUDA events) */
    double k1Ms;          /* ... of the lane-per-alignment sweep kernel alone */
    int launches;         /* kernel launches of the last compute */
    int filterWindows;     /* window sweeps planned by the candidate filter (all stages) */
    long long h2dBytes;   /* host->device bytes since the batch was prepared */
    long long d2hBytes;   /* device->host bytes */
    long long k1Cells;    /* nominal DP cells (sum queryLength*targetLength) swept by that kernel */
    long long wCells;     /* nominal DP cells of the distance pass swept by the warp kernel */
    long long filterDecided;  /* alignments settled by the candidate filter (prefix sweep + window) */
    long long filterFallback; /* alignments the filter could not decide (took the plain full sweep) */
} EdlibB200Stats;

/* Selects the CUDA device this process will use; call before any other entry point (one
 * process drives one GPU).  Returns EDLIB_STATUS_OK, or EDLIB_STATUS_ERROR if the engine was
 * already initialised on another device or the device does not exist. */
EDLIB_API int edlibB200SetDevice(int device);

/* NUMA node the selected device hangs off (-1: unknown).  Host buffers a caller hands to edlibAlignBatch travel fastest
 * when they live on that node (page-locked buffers are read by the device directly). */
EDLIB_API int edlibB200DeviceNumaNode(void);

/* Optional, for callers that receive millions of results per call: every result owns malloc'd arrays (the reference's
 * ownership rule), and glibc by default hands freed heap memory back to the kernel, so each batch pays the page faults
 * of tens of megabytes of fresh heap again -- serialised across the threads that build the results.  This call tells the
 * allocator to keep freed memory (mallopt M_TRIM_THRESHOLD / M_TOP_PAD); it changes nothing but the process's resident
 * set.  Returns EDLIB_STATUS_OK when the allocator took the settings. */
EDLIB_API int edlibB200TuneHostAllocator(void);

/* free() the arrays of `n` results at once (same effect as n edlibFreeAlignResult calls). */
EDLIB_API void edlibB200FreeResults(EdlibAlignResult* results, int n);

/* 1 when a CUDA device and the kernels are usable, else 0 (then every align call fails). */
EDLIB_API int edlibB200Available(void);

/* Text of the last engine error on this process (valid until the next call). */
EDLIB_API const char* edlibB200LastError(void);

EDLIB_API EdlibB200Batch* edlibB200BatchPrepare(const char* const* queries, const int* queryLengths,
                                                const char* const* targets, const int* targetLengths,
                                                int numPairs, const EdlibAlignConfig config);
EDLIB_API int edlibB200BatchCompute(EdlibB200Batch* batch, EdlibB200Stats* statsOut);
EDLIB_API int edlibB200BatchResults(EdlibB200Batch* batch, EdlibAlignResult* results);
EDLIB_API void edlibB200BatchFree(EdlibB200Batch* batch);

/* A target kept resident on the device.  edlibAlignBatch calls of read sets (HW, short queries, plain equality) whose
 * targets[i] all equal (target, targetLength) of a live handle skip the target's upload, its encoding and the build of
 * its seed index: a caller that aligns many batches to one genome pays them once.  The bytes at `target` must not
 * change while the handle lives; results are identical with and without a handle.  Returns NULL on failure. */
typedef struct EdlibB200Target EdlibB200Target;
EDLIB_API EdlibB200Target* edlibB200TargetPrepare(const char* target, int targetLength);
EDLIB_API void edlibB200TargetFree(EdlibB200Target* target);

/* edlibAlignmentToCigar (edlib.h) for n results at once, on the engine's host threads: cigars[i] receives a
 * malloc'd C string (caller frees each with free()), or NULL where results[i] holds no alignment.  Returns
 * EDLIB_STATUS_OK, or EDLIB_STATUS_ERROR on a bad format / operation code (then every cigars[i] is NULL). */
EDLIB_API int edlibB200AlignmentsToCigar(const EdlibAlignResult* results, int n, EdlibCigarFormat cigarFormat, char** cigars);

/* free() n strings of edlibB200AlignmentsToCigar at once (on the engine's host threads, like edlibB200FreeResults). */
EDLIB_API void edlibB200FreeCigars(char** cigars, int n);

/* Stats of the most recent edlibAlign / edlibAlignBatch / BatchCompute on this process. */
EDLIB_API void edlibB200LastStats(EdlibB200Stats* statsOut);

/* Device time per kernel of the most recent compute, as text "name:milliseconds:launches;..." written
 * to buf (NUL-terminated, truncated to bufLen); returns the full length. */
EDLIB_API int edlibB200LastKernelReport(char* buf, int bufLen);

#ifdef __cplusplus
}
#endif
#endif /* EDLIB_B200_H */
