// eb_capi.cpp -- the extern "C" surface declared in include/edlib.h and include/edlib_b200.h.
// Thin: argument checks, one process-wide engine behind a mutex (the reference API is
// re-entrant and is called with the GIL released, ref bindings/python/edlib.pyx:128-129), and
// the two pure-host helpers that carry no DP work (config constructors, CIGAR run-length
// encoding, free).
#include <malloc.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <functional>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/edlib.h"
#include "../../include/edlib_b200.h"
#include "eb_engine.h"

namespace eb {
void host_parallel_ranges(size_t n, size_t grain, const std::function<void(size_t, size_t)>& fn);  // eb_engine.cpp
Backend* create_backend(std::string* err);  // provided by the backend object linked into this library
int select_device(int device, std::string* err);  // 0 on success
}

namespace {

std::mutex g_mu;
eb::Backend* g_backend = nullptr;
eb::Engine* g_engine = nullptr;
std::string g_initError;
bool g_initTried = false;

// Small calls (single edlibAlign calls, batches of a few pairs) do not queue behind each other or behind a large batch:
// beside the main engine (large batches, the staged API, target handles; under g_mu) a few more engines, each with
// its own streams, scratch and lock, take them round robin -- concurrent callers (the reference's Python binding
// releases the GIL around edlibAlign, bindings/python/edlib.pyx:128-129) overlap their launches and copies.
struct SideEngine {
    std::mutex mu;
    eb::Backend* be = nullptr;
    eb::Engine* eng = nullptr;
};
constexpr int kMaxSide = 8;
SideEngine g_side[kMaxSide];
int g_numSide = -1;  // -1: not decided yet
std::atomic<unsigned> g_nextSide{0};
constexpr int kSmallBatch = 256;  // pairs; larger batches fill the device on their own and use the main engine
thread_local eb::Engine* t_lastEngine = nullptr;  // engine the calling thread used last (LastStats / LastKernelReport)
thread_local std::string t_lastError;              // text of the last failure of a call made by this thread

eb::Engine* engine_locked() {
    if (!g_initTried) {
        g_initTried = true;
        g_backend = eb::create_backend(&g_initError);
        if (g_backend) g_engine = new eb::Engine(g_backend);
    }
    // the caller may be any host thread: CUDA's current device is per thread
    if (g_backend) {
        try {
            g_backend->bind_thread();
        } catch (const std::exception& e) {
            if (g_engine) g_engine->lastError = e.what();
            return nullptr;
        }
    }
    return g_engine;
}

// A side engine for a small call, locked (nullptr: none configured / the device is unusable: use the main engine).
SideEngine* side_engine_acquire() {
    {
        std::lock_guard<std::mutex> lock(g_mu);
        if (!engine_locked()) return nullptr;
        if (g_numSide < 0) {
            const char* e = getenv("EDLIB_B200_ENGINES");
            g_numSide = std::max(0, std::min(kMaxSide, e && *e ? atoi(e) - 1 : 3));
        }
    }
    if (g_numSide <= 0) return nullptr;
    SideEngine* s = &g_side[g_nextSide.fetch_add(1, std::memory_order_relaxed) % (unsigned)g_numSide];
    s->mu.lock();
    if (!s->eng) {
        std::string err;
        s->be = eb::create_backend(&err);
        if (s->be) s->eng = new eb::Engine(s->be);
        if (!s->eng) {
            s->mu.unlock();
            return nullptr;
        }
    }
    try {
        s->be->bind_thread();
    } catch (...) {
        s->mu.unlock();
        return nullptr;
    }
    return s;
}

void fail_results(EdlibAlignResult* results, int n) {
    for (int i = 0; i < n; ++i) {
        memset(&results[i], 0, sizeof(results[i]));
        results[i].status = EDLIB_STATUS_ERROR;
        results[i].editDistance = -1;
    }
}

}  // namespace

extern "C" {

// ref edlib.cpp:1465-1475
EDLIB_API EdlibAlignConfig edlibNewAlignConfig(int k, EdlibAlignMode mode, EdlibAlignTask task,
                                               const EdlibEqualityPair* additionalEqualities,
                                               int additionalEqualitiesLength) {
    EdlibAlignConfig c;
    c.k = k;
    c.mode = mode;
    c.task = task;
    c.additionalEqualities = additionalEqualities;
    c.additionalEqualitiesLength = additionalEqualitiesLength;
    return c;
}

// ref edlib.cpp:1477-1479
EDLIB_API EdlibAlignConfig edlibDefaultAlignConfig(void) {
    return edlibNewAlignConfig(-1, EDLIB_MODE_NW, EDLIB_TASK_DISTANCE, NULL, 0);
}

// ref edlib.cpp:1481-1485
EDLIB_API void edlibFreeAlignResult(EdlibAlignResult result) {
    free(result.endLocations);
    free(result.startLocations);
    free(result.alignment);
}

EDLIB_API int edlibAlignBatch(const char* const* queries, const int* queryLengths,
                              const char* const* targets, const int* targetLengths,
                              int numPairs, const EdlibAlignConfig config, EdlibAlignResult* results) {
    if (numPairs < 0 || (numPairs > 0 && (!queries || !queryLengths || !targets || !targetLengths || !results)))
        return EDLIB_STATUS_ERROR;
    if (numPairs == 0) return EDLIB_STATUS_OK;
    eb::BatchInput in{queries, queryLengths, targets, targetLengths, numPairs, config};
    if (numPairs <= kSmallBatch) {
        if (SideEngine* s = side_engine_acquire()) {
            t_lastEngine = s->eng;
            const int rc = s->eng->align_batch(in, results);
            if (rc != EDLIB_STATUS_OK) t_lastError = s->eng->lastError;  // (copied while the engine is still ours)
            s->mu.unlock();
            return rc;
        }
    }
    std::lock_guard<std::mutex> lock(g_mu);
    eb::Engine* e = engine_locked();
    if (!e) {  // no usable device: fail loudly, there is no CPU path
        fail_results(results, numPairs);
        return EDLIB_STATUS_ERROR;
    }
    t_lastEngine = e;
    const int rc = e->align_batch(in, results);
    if (rc != EDLIB_STATUS_OK) t_lastError = e->lastError;
    return rc;
}

// ref edlib.cpp:146-301
EDLIB_API EdlibAlignResult edlibAlign(const char* query, int queryLength, const char* target, int targetLength,
                                      const EdlibAlignConfig config) {
    EdlibAlignResult r;
    const char* q = query ? query : "";
    const char* t = target ? target : "";
    edlibAlignBatch(&q, &queryLength, &t, &targetLength, 1, config, &r);
    return r;
}

// ref edlib.cpp:303-350.  Pure formatting of an existing edit script (no DP): kept on the host.
// One pass: runs are found eight operations at a time (EXTENDED: a run is a stretch of equal bytes) and written into a
// scratch buffer of the worst-case size (every run "1X": two characters per operation), then copied into one exact malloc.
static inline int cigar_run_end(const unsigned char* a, int i, int len, unsigned char op) {
    int j = i + 1;
    const uint64_t pat = 0x0101010101010101ull * op;
    while (j + 8 <= len) {
        uint64_t w;
        memcpy(&w, a + j, 8);
        const uint64_t x = w ^ pat;
        if (x) return j + (__builtin_ctzll(x) >> 3);  // first byte that differs (little endian)
        j += 8;
    }
    while (j < len && a[j] == op) ++j;
    return j;
}

// STANDARD format: MATCH (0) and MISMATCH (3) both print 'M' -- the two codes whose low bits agree.  Any other byte
// (I, D, or a bad code) ends the run.
static inline int cigar_m_run_end(const unsigned char* a, int i, int len) {
    int j = i + 1;
    const uint64_t ones = 0x0101010101010101ull;
    while (j + 8 <= len) {
        uint64_t w;
        memcpy(&w, a + j, 8);
        const uint64_t x = ((w ^ (w >> 1)) & ones) | (w & ~(3 * ones));  // per byte: bit0 != bit1, or a bit above them
        if (x) return j + (__builtin_ctzll(x) >> 3);
        j += 8;
    }
    while (j < len && (a[j] == 0 || a[j] == 3)) ++j;
    return j;
}

EDLIB_API char* edlibAlignmentToCigar(const unsigned char* alignment, int alignmentLength, EdlibCigarFormat cigarFormat) {
    if (cigarFormat != EDLIB_CIGAR_EXTENDED && cigarFormat != EDLIB_CIGAR_STANDARD) return NULL;
    const bool ext = cigarFormat == EDLIB_CIGAR_EXTENDED;
    const char* sym = ext ? "=IDX" : "MIDM";
    const int len = alignmentLength < 0 ? 0 : alignmentLength;
    char stackBuf[2048];
    const size_t worst = 2 * (size_t)len + 1;
    char* buf = worst <= sizeof(stackBuf) ? stackBuf : static_cast<char*>(malloc(worst));
    if (!buf) return NULL;
    char* w = buf;
    bool bad = false;
    for (int i = 0; i < len;) {
        const unsigned char op = alignment[i];
        if (op > 3) {
            bad = true;
            break;
        }
        const char c = sym[op];
        // STANDARD: MATCH and MISMATCH share 'M' (ref cpp:311-314); insertions and deletions are plain runs in both formats
        const int j = (ext || c != 'M') ? cigar_run_end(alignment, i, len, op) : cigar_m_run_end(alignment, i, len);
        int run = j - i;
        char digits[12];
        int nd = 0;
        for (; run; run /= 10) digits[nd++] = (char)('0' + run % 10);
        while (nd) *w++ = digits[--nd];
        *w++ = c;
        i = j;
    }
    char* res = NULL;
    if (!bad) {
        const size_t bytes = (size_t)(w - buf) + 1;
        res = static_cast<char*>(malloc(bytes));
        if (res) {
            memcpy(res, buf, bytes - 1);
            res[bytes - 1] = 0;
        }
    }
    if (buf != stackBuf) free(buf);
    return res;
}

// ---- include/edlib_b200.h ------------------------------------------------------------------

EDLIB_API const char* edlibB200LastError(void) {
    // every calling thread reads its own copy: a later call on another thread cannot change it under the reader
    static thread_local std::string copy;
    std::lock_guard<std::mutex> lock(g_mu);
    // failures of edlibAlign / edlibAlignBatch are recorded per calling thread (side engines run concurrently); the
    // staged / handle entry points all use the main engine under this lock
    if (t_lastEngine && t_lastEngine != g_engine) copy = t_lastError;
    else copy = g_engine ? g_engine->lastError : g_initError;
    return copy.c_str();
}

EDLIB_API int edlibB200SetDevice(int device) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_initTried) return EDLIB_STATUS_ERROR;
    // remembered by the backend factory: the engine binds to THIS device whichever thread makes the first call
    return eb::select_device(device, &g_initError) == 0 ? EDLIB_STATUS_OK : EDLIB_STATUS_ERROR;
}

EDLIB_API int edlibB200DeviceNumaNode(void) {
    std::lock_guard<std::mutex> lock(g_mu);
    return engine_locked() ? g_backend->numa_node() : -1;
}

EDLIB_API int edlibB200TuneHostAllocator(void) {
#if defined(__GLIBC__)
    const int a = mallopt(M_TRIM_THRESHOLD, 1 << 30);  // freed heap memory stays with the allocator
    const int b = mallopt(M_TOP_PAD, 64 << 20);        // heaps grow in large steps
    return (a && b) ? EDLIB_STATUS_OK : EDLIB_STATUS_ERROR;
#else
    return EDLIB_STATUS_ERROR;
#endif
}

// Large result sets are freed on the host pool, the way they were built: glibc re-serves chunks that ONE thread freed
// (cold in every other core's cache, threaded through its bins one by one) several times slower than chunks the worker
// threads freed themselves -- measured 17 ns vs 4 ns per malloc with eight threads (profiles/README.md).
EDLIB_API void edlibB200FreeResults(EdlibAlignResult* results, int n) {
    if (!results || n <= 0) return;
    auto free_range = [results](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) {
            free(results[i].endLocations);
            free(results[i].startLocations);
            free(results[i].alignment);
            results[i].endLocations = results[i].startLocations = NULL;
            results[i].alignment = NULL;
        }
    };
    if (n < 65536) {
        free_range(0, (size_t)n);
        return;
    }
    std::lock_guard<std::mutex> lock(g_mu);  // the host pool serves one client at a time
    eb::host_parallel_ranges((size_t)n, 16384, free_range);
}

EDLIB_API int edlibB200Available(void) {
    std::lock_guard<std::mutex> lock(g_mu);
    return engine_locked() ? 1 : 0;
}

EDLIB_API EdlibB200Batch* edlibB200BatchPrepare(const char* const* queries, const int* queryLengths,
                                                const char* const* targets, const int* targetLengths,
                                                int numPairs, const EdlibAlignConfig config) {
    std::lock_guard<std::mutex> lock(g_mu);
    eb::Engine* e = engine_locked();
    if (!e || numPairs <= 0) return NULL;
    t_lastEngine = e;
    try {
        eb::BatchInput in{queries, queryLengths, targets, targetLengths, numPairs, config};
        return reinterpret_cast<EdlibB200Batch*>(e->prepare(in));
    } catch (const std::exception& ex) {
        e->lastError = ex.what();
        return NULL;
    }
}

EDLIB_API int edlibB200BatchCompute(EdlibB200Batch* batch, EdlibB200Stats* statsOut) {
    std::lock_guard<std::mutex> lock(g_mu);
    eb::Engine* e = engine_locked();
    if (!e || !batch) return EDLIB_STATUS_ERROR;
    t_lastEngine = e;
    try {
        e->stats = eb::EngineStats();
        e->compute(reinterpret_cast<eb::Prepared*>(batch));
    } catch (const std::exception& ex) {
        e->lastError = ex.what();
        return EDLIB_STATUS_ERROR;
    }
    if (statsOut) {
        e->finish_stats();
        statsOut->kernelMs = e->stats.kernelMs;
        statsOut->k1Ms = e->stats.k1Ms;
        statsOut->launches = e->stats.launches;
        statsOut->h2dBytes = e->stats.h2dBytes;
        statsOut->d2hBytes = e->stats.d2hBytes;
        statsOut->k1Cells = e->stats.k1Cells;
        statsOut->wCells = e->stats.wCells;
        statsOut->filterDecided = e->stats.filterDecided;
        statsOut->filterFallback = e->stats.filterFallback;
        statsOut->filterWindows = (int)std::min<long long>(e->stats.filterWindows, 0x7fffffff);
    }
    return EDLIB_STATUS_OK;
}

EDLIB_API int edlibB200BatchResults(EdlibB200Batch* batch, EdlibAlignResult* results) {
    std::lock_guard<std::mutex> lock(g_mu);
    eb::Engine* e = engine_locked();
    if (!e || !batch || !results) return EDLIB_STATUS_ERROR;
    t_lastEngine = e;
    try {
        e->materialize(reinterpret_cast<eb::Prepared*>(batch), results);
    } catch (const std::exception& ex) {
        e->lastError = ex.what();
        return EDLIB_STATUS_ERROR;
    }
    return EDLIB_STATUS_OK;
}

EDLIB_API void edlibB200BatchFree(EdlibB200Batch* batch) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (batch && engine_locked()) g_engine->release(reinterpret_cast<eb::Prepared*>(batch));
}

EDLIB_API EdlibB200Target* edlibB200TargetPrepare(const char* target, int targetLength) {
    std::lock_guard<std::mutex> lock(g_mu);
    eb::Engine* e = engine_locked();
    if (!e) return NULL;
    t_lastEngine = e;
    try {
        return reinterpret_cast<EdlibB200Target*>(e->target_prepare(target, targetLength));
    } catch (const std::exception& ex) {
        e->lastError = ex.what();
        return NULL;
    }
}

EDLIB_API void edlibB200TargetFree(EdlibB200Target* target) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (target && engine_locked()) g_engine->target_free(reinterpret_cast<eb::TargetHandle*>(target));
}

EDLIB_API int edlibB200AlignmentsToCigar(const EdlibAlignResult* results, int n, EdlibCigarFormat cigarFormat, char** cigars) {
    if (n < 0 || (n > 0 && (!results || !cigars))) return EDLIB_STATUS_ERROR;
    if (cigarFormat != EDLIB_CIGAR_EXTENDED && cigarFormat != EDLIB_CIGAR_STANDARD) return EDLIB_STATUS_ERROR;
    std::lock_guard<std::mutex> lock(g_mu);  // the host pool serves one client at a time
    std::atomic<int> bad(0);
    eb::host_parallel_ranges((size_t)n, 4096, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) {
            cigars[i] = NULL;
            if (!results[i].alignment || results[i].alignmentLength <= 0) continue;
            cigars[i] = edlibAlignmentToCigar(results[i].alignment, results[i].alignmentLength, cigarFormat);
            if (!cigars[i]) bad.store(1, std::memory_order_relaxed);
        }
    });
    if (!bad.load()) return EDLIB_STATUS_OK;
    for (int i = 0; i < n; ++i) {
        free(cigars[i]);
        cigars[i] = NULL;
    }
    return EDLIB_STATUS_ERROR;
}

EDLIB_API void edlibB200FreeCigars(char** cigars, int n) {
    if (!cigars || n <= 0) return;
    auto free_range = [cigars](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) {
            free(cigars[i]);
            cigars[i] = NULL;
        }
    };
    if (n < 65536) {
        free_range(0, (size_t)n);
        return;
    }
    std::lock_guard<std::mutex> lock(g_mu);  // the host pool serves one client at a time
    eb::host_parallel_ranges((size_t)n, 16384, free_range);
}

EDLIB_API void edlibB200LastStats(EdlibB200Stats* s) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (!s) return;
    memset(s, 0, sizeof(*s));
    if (!g_engine || !engine_locked()) return;
    eb::Engine* e = t_lastEngine ? t_lastEngine : g_engine;  // the engine this thread's last call ran on
    e->finish_stats();
    s->kernelMs = e->stats.kernelMs;
    s->k1Ms = e->stats.k1Ms;
    s->launches = e->stats.launches;
    s->h2dBytes = e->stats.h2dBytes;
    s->d2hBytes = e->stats.d2hBytes;
    s->k1Cells = e->stats.k1Cells;
    s->wCells = e->stats.wCells;
    s->filterDecided = e->stats.filterDecided;
    s->filterFallback = e->stats.filterFallback;
    s->filterWindows = (int)std::min<long long>(e->stats.filterWindows, 0x7fffffff);
}

EDLIB_API int edlibB200LastKernelReport(char* buf, int bufLen) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (!buf || bufLen <= 0) return 0;
    eb::Engine* e = t_lastEngine ? t_lastEngine : g_engine;
    if (e && engine_locked()) e->finish_stats();
    const std::string r = e ? e->stats.kernelReport : std::string();
    const int n = (int)std::min<size_t>(r.size(), (size_t)bufLen - 1);
    memcpy(buf, r.data(), (size_t)n);
    buf[n] = 0;
    return (int)r.size();
}

}  // extern "C"
