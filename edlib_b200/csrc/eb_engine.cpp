// eb_engine.cpp -- batch planner / orchestrator (see eb_engine.h).
//
// Pipeline of one batch (reference driver: edlibAlign, ref edlib.cpp:146-301):
//   prepare : pack + upload raw bytes, byte-presence sets on the device (-> alphabetLength per
//             pair, ref cpp:1417-1462), dense code map, in-place encoding, equality table
//             (ref cpp:63-94)
//   compute : distance + end locations  (K1 lane-per-alignment for groups that share a target,
//             W warp-per-alignment otherwise; ref cpp:199-225)
//             start locations           (reversed SHW sweeps, ref cpp:228-272)
//             alignment path            (stored-matrix NW sweep + traceback kernel inside the
//                                        reference's 1 MiB rule, ref cpp:276-289, 1161-1213)
//   materialize : malloc'd arrays per result (ownership as ref edlib.h:177,186,205)
#include "eb_engine.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <functional>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <chrono>
#include <map>
#include <stdexcept>
#include <thread>
#include <unordered_map>

namespace eb {

static int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

EngineTunables::EngineTunables() {
    k1MinGroup = env_int("EDLIB_B200_K1_MIN_GROUP", k1MinGroup);
    k1MinChunk = env_int("EDLIB_B200_K1_MIN_CHUNK", k1MinChunk);
    ovfCap = env_int("EDLIB_B200_OVF_CAP", ovfCap);
    filterK0 = env_int("EDLIB_B200_FILTER_K0", filterK0);
    filterK1 = env_int("EDLIB_B200_FILTER_K1", filterK1);
    packParallelBytes = (size_t)env_int("EDLIB_B200_PACK_PARALLEL_KB", (int)(packParallelBytes >> 10)) << 10;
    filterSeedK = env_int("EDLIB_B200_FILTER_SEED_K", filterSeedK);
    filterSeedBucket = env_int("EDLIB_B200_FILTER_SEED_BUCKET", filterSeedBucket);
    filterSeedSlack = env_int("EDLIB_B200_FILTER_SEED_SLACK", filterSeedSlack);
    filterSeedLevels = std::min(SEED_LEVELS, env_int("EDLIB_B200_FILTER_SEED_LEVELS", filterSeedLevels));
    filterMaxWindows = env_int("EDLIB_B200_FILTER_MAX_WINDOWS", filterMaxWindows);
    filterMinLen = env_int("EDLIB_B200_FILTER_MIN_LEN", filterMinLen);
    filterSpread = env_int("EDLIB_B200_FILTER_SPREAD", filterSpread);
    filterMinTarget = env_int("EDLIB_B200_FILTER_MIN_TARGET", filterMinTarget);
    const int sliceMb = env_int("EDLIB_B200_SLICE_MB", 0);
    if (sliceMb > 0) sliceBytes = (size_t)sliceMb << 20;
}

namespace {

// EDLIB_B200_TRACE=1: wall-clock of the host phases to stderr (diagnostics only).
struct Trace {
    bool on;
    std::chrono::steady_clock::time_point t0;
    Trace() : on(getenv("EDLIB_B200_TRACE") != nullptr), t0(std::chrono::steady_clock::now()) {}
    void mark(const char* what) {
        if (!on) return;
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[edlib_b200] %-28s %8.2f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

// Persistent host worker threads (spawning threads per loop costs more than most of these loops): run(n, fn)
// executes fn(0) .. fn(n-1), the caller taking part, and returns when all are done.  One client at a time
// (the engine runs under the library's lock).  The workers live until the process ends.
class HostPool {
public:
    static HostPool& get() {
        static HostPool* pool = new HostPool();  // never destroyed: workers may still be parked at exit
        return *pool;
    }
    size_t width() const { return workers_ + 1; }
    void run(size_t n, const std::function<void(size_t)>& fn) {
        if (n == 0) return;
        if (n == 1 || workers_ == 0) {
            for (size_t i = 0; i < n; ++i) fn(i);
            return;
        }
        unsigned long long gen;
        {
            std::lock_guard<std::mutex> lock(mu_);
            fn_ = &fn;
            total_ = n;
            next_ = 0;
            pending_ = n;
            gen = ++generation_;
        }
        cv_.notify_all();
        work(gen);
        std::unique_lock<std::mutex> lock(mu_);
        done_.wait(lock, [this]() { return pending_ == 0; });
        fn_ = nullptr;
    }

private:
    HostPool() {
        const size_t hw = std::max(1u, std::thread::hardware_concurrency());
        workers_ = std::min<size_t>(16, hw) - 1;
        for (size_t i = 0; i < workers_; ++i) std::thread([this]() { loop(); }).detach();
    }
    // Tasks are few and coarse, so they are claimed under the lock; a worker only ever claims tasks of
    // the generation it woke up for.
    void work(unsigned long long gen) {
        for (;;) {
            const std::function<void(size_t)>* fn;
            size_t i;
            {
                std::lock_guard<std::mutex> lock(mu_);
                if (generation_ != gen || next_ >= total_) return;
                i = next_++;
                fn = fn_;
            }
            (*fn)(i);
            std::lock_guard<std::mutex> lock(mu_);
            if (--pending_ == 0) done_.notify_all();
        }
    }
    void loop() {
        unsigned long long seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lock(mu_);
                cv_.wait(lock, [&]() { return generation_ != seen; });
                seen = generation_;
            }
            work(seen);
        }
    }
    size_t workers_ = 0;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    const std::function<void(size_t)>* fn_ = nullptr;
    size_t total_ = 0, pending_ = 0, next_ = 0;
    unsigned long long generation_ = 0;
};

// number of parts a loop over n items is cut into (every part gets >= grain items)
inline size_t host_parts(size_t n, size_t grain) {
    return std::max<size_t>(1, std::min<size_t>(HostPool::get().width(), n / std::max<size_t>(grain, 1)));
}

// fn(begin, end) over [0, n) on the host workers (only when every part gets >= grain items).
template <class F>
void parallel_ranges(size_t n, size_t grain, F fn) {
    const size_t nthr = host_parts(n, grain);
    if (nthr <= 1) {
        fn((size_t)0, n);
        return;
    }
    HostPool::get().run(nthr, [&](size_t t) { fn(n * t / nthr, n * (t + 1) / nthr); });
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t round_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// Host staging memory from the backend (pinned and cached on CUDA): transfers run at full PCIe rate and
// the host reads / writes it in place.
template <class T>
struct HostBuf {
    Backend* be = nullptr;
    T* p = nullptr;
    size_t n = 0;
    HostBuf(Backend* b, size_t count) : be(b), p(static_cast<T*>(b->alloc_host(std::max<size_t>(count, 1) * sizeof(T)))), n(count) {}
    HostBuf(const HostBuf&) = delete;
    HostBuf& operator=(const HostBuf&) = delete;
    ~HostBuf() { be->free_host(p); }
    T& operator[](size_t i) { return p[i]; }
    const T& operator[](size_t i) const { return p[i]; }
};

template <class T>
struct DevBuf {
    Backend* be = nullptr;
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    DevBuf(Backend* b, size_t count) { alloc(b, count); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { reset(); }
    void alloc(Backend* b, size_t count) {
        reset();
        be = b;
        n = count;
        p = static_cast<T*>(be->alloc(std::max<size_t>(count, 1) * sizeof(T)));
    }
    void reset() {
        if (p) be->free(p);
        p = nullptr;
        n = 0;
    }
    void upload(const T* src, size_t count) { be->h2d(p, src, count * sizeof(T)); }
    void download(T* dst, size_t count) { be->d2h(dst, p, count * sizeof(T)); }
};

struct Target {
    const char* ptr;
    int len;
    uint64_t off;  // into the packed sequence buffer
};

// One warp-per-alignment sweep as seen by the host.
struct WTask {
    uint64_t qOff = 0, tOff = 0;
    int m = 0, n = 0, mode = 0, flags = 0, kInit = 0, dhi = 0, stopCol = -1, trackFrom = 0;
    int R = 1, nWp = 0;
    int pair = -1, tag = 0;
    bool wantPositions = false;  // the caller needs every end position, not just best/cnt/last
    int splitSide = -1;          // WF_STOPCOL pairs of a Hirschberg node: 0 forward half, 1 reversed half (adjacent tasks)
    int splitBest = 0;           // ... the node's known score
    SplitOut split{};            // ... the split found on the device (stored on the forward task)
    Rec rec{};
    std::vector<int> extra;  // positions past KPOS that attain rec.best, ascending
    long long opsOff = -1;   // into the ops pool (WF_STORE)
    int opsLen = 0;
};

struct WPlan {
    int R, nWp;
    bool slide;
    int dhi;
};

// Window shape of a W job.  Short queries (<= 1024 rows) always fit one fixed window, which is
// exact for any k.  Longer NW jobs with a bound use one window sliding down the Ukkonen band
// (cells with |d| + |delta - d| <= k, d = c - r: ref cpp:755, 799-830 keep the same cells) when
// the band is at most half of the query; everything else is swept unbanded in strips.
WPlan plan_w(int m, int n, int mode, int kBound) {
    WPlan pl;
    const int nW = ceil_div(m, 32);
    pl.slide = false;
    pl.dhi = 0;
    if (nW <= 32) {
        pl.R = 1;
        pl.nWp = nW;
        return pl;
    }
    pl.nWp = (int)round_up((size_t)nW, 8);
    if (mode == MODE_NW && kBound >= 0) {
        const int d = n - m;
        const int ad = d < 0 ? -d : d;
        const long long h = ((long long)kBound - ad) / 2;
        const long long dlo = std::min(0, d) - h, dhi = std::max(0, d) + h;
        const long long height = dhi - dlo + 1;
        for (int R = 1; R <= 8; R *= 2) {
            if (height + 32LL * R <= 1024LL * R && 64 * R <= nW) {
                pl.R = R;
                pl.slide = true;
                pl.dhi = (int)dhi;
                return pl;
            }
        }
    }
    pl.R = 8;
    for (int R = 2; R <= 8; R *= 2)
        if (32 * R >= pl.nWp) {
            pl.R = R;
            break;
        }
    return pl;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// Prepared batch
// ---------------------------------------------------------------------------------------------
class Prepared {
public:
    Backend* be = nullptr;
    int N = 0;
    EdlibAlignConfig cfg{};
    int mode = MODE_NW;  // normalised: anything that is not SHW/HW runs as NW (ref cpp:205-215)
    std::vector<int> qlen, tlen, tidx;
    std::vector<uint64_t> qoff;
    std::vector<Target> tg;
    DevBuf<uint8_t> dSeq;
    DevBuf<uint64_t> dQoff;
    DevBuf<int> dQlen;
    DevBuf<uint8_t> dEqtab;
    bool hasEq = false;
    int ncodes = 0;
    std::vector<int> alphaLen;

    // classification (Engine::classify): pairs per (target, word class) for the lane kernels, the rest
    struct Part {
        std::map<std::pair<int, int>, std::vector<int>> groups;
        std::vector<int> wPairs;
    };
    std::map<std::pair<int, int>, std::vector<int>> groups;
    std::vector<Part> parts;        // per-thread pieces, kept for their storage
    std::vector<int> wPairsBase;    // queries above 256 rows
    bool classified = false;

    // results
    std::vector<int> ed;            // distance or -1
    std::vector<uint8_t> special;   // 1: an empty sequence (ref cpp:166-184)
    std::vector<long long> endStart;
    std::vector<int> endCount;
    std::vector<int> endPool, startPool;
    std::vector<long long> alnStart;  // -1: none
    std::vector<int> alnLen;
    std::vector<uint8_t> alnPool;
    bool computed = false;
};

// ---------------------------------------------------------------------------------------------
// prepare
// ---------------------------------------------------------------------------------------------
Prepared* Engine::prepare(const BatchInput& in) {
    Backend* be = be_;
    Trace trace;
    // the host vectors of the previous batch are reused (their pages stay mapped)
    Prepared* p = spare_ ? spare_ : new Prepared();
    spare_ = nullptr;
    try {
        p->tg.clear();
        p->hasEq = false;
        p->ncodes = 0;
        p->computed = false;
        p->classified = false;
        p->be = be;
        p->N = in.numPairs;
        p->cfg = in.config;
        p->mode = (in.config.mode == EDLIB_MODE_SHW) ? MODE_SHW : (in.config.mode == EDLIB_MODE_HW) ? MODE_HW : MODE_NW;
        const int N = p->N;
        p->qlen.resize(N);
        p->tlen.resize(N);
        p->tidx.resize(N);
        p->qoff.resize(N);
        {
            std::vector<int> bad(1, 0);
            parallel_ranges((size_t)N, 65536, [&](size_t lo, size_t hi) {
                for (size_t i = lo; i < hi; ++i) {
                    p->qlen[i] = in.queryLengths[i];
                    p->tlen[i] = in.targetLengths[i];
                    if (p->qlen[i] < 0 || p->tlen[i] < 0) bad[0] = 1;
                }
            });
            if (bad[0]) throw std::runtime_error("negative sequence length");
        }

        // identical (pointer, length) targets are uploaded and encoded once
        struct Key {
            const char* ptr;
            int len;
            bool operator==(const Key& o) const { return ptr == o.ptr && len == o.len; }
        };
        struct KeyHash {
            size_t operator()(const Key& k) const { return std::hash<const void*>()(k.ptr) * 31 + (size_t)k.len; }
        };
        std::unordered_map<Key, int, KeyHash> seen;
        Key lastKey{nullptr, -1};
        int lastIdx = -1;
        bool oneTarget = N > 0;  // the usual batch shape (reads over one shared target), checked in parallel
        if (N >= 131072) {
            std::vector<int> differs(1, 0);
            const Key first{in.targets[0], in.targetLengths[0]};
            parallel_ranges((size_t)N, 65536, [&](size_t lo, size_t hi) {
                for (size_t i = lo; i < hi; ++i)
                    if (in.targets[i] != first.ptr || in.targetLengths[i] != first.len) differs[0] = 1;
            });
            oneTarget = !differs[0];
        } else {
            oneTarget = false;
        }
        if (oneTarget) {
            p->tg.push_back(Target{in.targets[0], in.targetLengths[0], 0});
            parallel_ranges((size_t)N, 65536, [&](size_t lo, size_t hi) {
                for (size_t i = lo; i < hi; ++i) p->tidx[i] = 0;
            });
        }
        for (int i = 0; i < N && !oneTarget; ++i) {
            Key k{in.targets[i], in.targetLengths[i]};
            if (!(k == lastKey)) {  // neighbours usually share their target
                auto it = seen.find(k);
                if (it == seen.end()) {
                    it = seen.emplace(k, (int)p->tg.size()).first;
                    p->tg.push_back(Target{k.ptr, k.len, 0});
                }
                lastKey = k;
                lastIdx = it->second;
            }
            p->tidx[i] = lastIdx;
        }
        const int T = (int)p->tg.size();
        trace.mark("prepare: lengths + distinct targets");

        // pack: queries back to back, then every target 16-aligned with >= 16 bytes of slack
        size_t total = 0;
        std::vector<int> longQueries;  // beyond one presence-set work item
        for (int i = 0; i < N; ++i) {
            p->qoff[i] = total;
            total += (size_t)p->qlen[i];
            if (p->qlen[i] > 65536) longQueries.push_back(i);
        }
        total = round_up(total, 16);
        for (int t = 0; t < T; ++t) {
            p->tg[t].off = total;
            total += round_up((size_t)p->tg[t].len, 16) + 16;
        }
        total += 16;
        uint8_t* stage = static_cast<uint8_t*>(be->alloc_host(total));
        p->dSeq.alloc(be, total);
        // the small per-pair arrays go first: they would otherwise queue behind the sequences
        p->dQoff.alloc(be, N);
        p->dQoff.upload(p->qoff.data(), N);
        p->dQlen.alloc(be, N);
        p->dQlen.upload(p->qlen.data(), N);
        trace.mark("prepare: offsets + buffers");
        {
            // Pure memcpy work, split by bytes over a few host threads when the batch is large: items
            // 0..N-1 are the queries, N..N+T-1 the distinct targets.  Every thread uploads its own byte
            // range as soon as it is packed, so the host->device copy overlaps the packing.
            const size_t qBytes = N ? (size_t)(p->qoff[N - 1] + (uint64_t)p->qlen[N - 1]) : 0;
            size_t allBytes = qBytes;
            for (int t = 0; t < T; ++t) allBytes += (size_t)p->tg[t].len;
            const int nthr = allBytes > tun.packParallelBytes ? (int)HostPool::get().width() : 1;
            auto copy_item = [&](int it) {
                if (it < N) {
                    if (p->qlen[it]) memcpy(stage + p->qoff[it], in.queries[it], (size_t)p->qlen[it]);
                } else {
                    const Target& g = p->tg[it - N];
                    if (g.len) memcpy(stage + g.off, g.ptr, (size_t)g.len);
                }
            };
            auto item_off = [&](int it) -> size_t {  // first staging byte of item `it` (N + T: the end)
                if (it >= N + T) return total;
                return it < N ? (size_t)p->qoff[it] : p->tg[it - N].off;
            };
            // padding first: it lies between the items and travels with the neighbouring byte ranges
            {
                size_t pos = qBytes;
                size_t end = p->tg.empty() ? total : p->tg[0].off;
                memset(stage + pos, 0, end - pos);
                for (int t = 0; t < T; ++t) {
                    const Target& g = p->tg[t];
                    const size_t next = (t + 1 < T) ? p->tg[t + 1].off : total;
                    memset(stage + g.off + g.len, 0, next - g.off - (size_t)g.len);
                }
            }
            if (nthr > 1) {
                // contiguous item ranges of roughly equal byte counts
                std::vector<int> cut(nthr + 1, N + T);
                cut[0] = 0;
                {
                    size_t tAcc = qBytes;  // bytes before target `tt`
                    int tt = 0;
                    for (int c = 1; c < nthr; ++c) {
                        const size_t want = allBytes * c / nthr;
                        if (want < qBytes) {  // qoff is the running byte count of the queries
                            cut[c] = (int)(std::upper_bound(p->qoff.begin(), p->qoff.end(), (uint64_t)want) - p->qoff.begin());
                        } else {
                            while (tt < T && tAcc + (size_t)p->tg[tt].len <= want) tAcc += (size_t)p->tg[tt++].len;
                            cut[c] = N + tt;
                        }
                        if (cut[c] < cut[c - 1]) cut[c] = cut[c - 1];
                    }
                }
                std::vector<std::string> errs(nthr);
                HostPool::get().run((size_t)nthr, [&](size_t t) {
                    try {
                        for (int it = cut[t]; it < cut[t + 1]; ++it) copy_item(it);
                        const size_t a = t == 0 ? 0 : item_off(cut[t]), b = item_off(cut[t + 1]);
                        be->bind_thread();  // a pool worker: select the backend's device before the copy
                        if (b > a) be->h2d(p->dSeq.p + a, stage + a, b - a);
                    } catch (const std::exception& e) {
                        errs[t] = e.what();
                    }
                });
                for (auto& e : errs)
                    if (!e.empty()) throw std::runtime_error(e);
            } else {
                for (int it = 0; it < N + T; ++it) copy_item(it);
                p->dSeq.upload(stage, total);
            }
        }
        trace.mark("prepare: pack");
        stats.h2dBytes += (long long)total;

        // byte-presence sets: one per query, one per distinct target, one union for the batch.  Queries are
        // implicit work items of the kernel; explicit ones (at most 65536 bytes each) are only needed for
        // the targets and for the pieces of longer queries.
        std::vector<MaskItem> items;
        auto add_items = [&](uint64_t off, int len, int dst) {
            for (int s0 = 0; s0 < len; s0 += 65536) items.push_back(MaskItem{off + (uint64_t)s0, std::min(65536, len - s0), dst});
        };
        for (int i : longQueries) add_items(p->qoff[i], p->qlen[i], i);
        for (int t = 0; t < T; ++t) add_items(p->tg[t].off, p->tg[t].len, N + t);
        HostBuf<int> tset(be, (size_t)N);
        parallel_ranges((size_t)N, 65536, [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; ++i) tset[i] = N + p->tidx[i];
        });
        const int unionSet = N + T;
        DevBuf<uint32_t> dMasks(be, (size_t)(N + T + 1) * 8);
        be->zero(dMasks.p, (size_t)(N + T + 1) * 8 * sizeof(uint32_t));
        DevBuf<MaskItem> dItems(be, items.size());
        if (!items.empty()) dItems.upload(items.data(), items.size());
        {
            MaskParams mp;
            memset(&mp, 0, sizeof(mp));
            mp.raw = p->dSeq.p;
            mp.items = dItems.p;
            mp.numItems = (int)items.size();
            mp.qoff = p->dQoff.p;
            mp.qlen = p->dQlen.p;
            mp.numQueries = N;
            mp.masks = dMasks.p;
            mp.unionSet = unionSet;
            if (mp.numItems + mp.numQueries > 0) be->launch_mask(mp);
        }
        DevBuf<int> dTset(be, N), dAlpha(be, N);
        dTset.upload(tset.p, N);
        trace.mark("prepare: mask items");
        be->launch_alpha_len(dMasks.p, nullptr, dTset.p, N, dAlpha.p);
        classify(p);  // host work while the upload and the alphabet kernels run
        trace.mark("prepare: classification");
        p->alphaLen.resize(N);
        dAlpha.download(p->alphaLen.data(), N);
        uint32_t uni[8];
        be->d2h(uni, dMasks.p + (size_t)unionSet * 8, sizeof(uni));
        stats.d2hBytes += (long long)N * 4 + 32;
        trace.mark("prepare: alphabet lengths back");

        // dense codes in ascending byte order; absent bytes (padding) map to code 0
        uint8_t map[256];
        memset(map, 0, sizeof(map));
        int byteOfCode[256];
        p->ncodes = 0;
        for (int b = 0; b < 256; ++b)
            if (uni[b >> 5] >> (b & 31) & 1u) {
                byteOfCode[p->ncodes] = b;
                map[b] = (uint8_t)p->ncodes++;
            }
        if (p->ncodes == 0) p->ncodes = 1;
        DevBuf<uint8_t> dMap(be, 256);
        dMap.upload(map, 256);
        EncodeParams ep{p->dSeq.p, (uint64_t)total, dMap.p};
        be->launch_encode(ep);

        // equality table over codes (ref cpp:63-94); a pair naming an absent byte changes nothing
        if (in.config.additionalEqualities && in.config.additionalEqualitiesLength > 0) {
            const int s = p->ncodes;
            std::vector<uint8_t> eq((size_t)s * s, 0);
            for (int i = 0; i < s; ++i) eq[(size_t)i * s + i] = 1;
            bool any = false;
            for (int i = 0; i < in.config.additionalEqualitiesLength; ++i) {
                const int a = (unsigned char)in.config.additionalEqualities[i].first;
                const int b = (unsigned char)in.config.additionalEqualities[i].second;
                const bool ha = uni[a >> 5] >> (a & 31) & 1u, hb = uni[b >> 5] >> (b & 31) & 1u;
                if (ha && hb) {
                    eq[(size_t)map[a] * s + map[b]] = eq[(size_t)map[b] * s + map[a]] = 1;
                    any = true;
                }
            }
            if (any) {
                p->dEqtab.alloc(be, eq.size());
                p->dEqtab.upload(eq.data(), eq.size());
                p->hasEq = true;
            }
        }
        (void)byteOfCode;
        be->sync();
        be->free_host(stage);
        trace.mark("prepare: upload+alphabet");
    } catch (...) {
        delete p;
        throw;
    }
    return p;
}

// ---------------------------------------------------------------------------------------------
// W runner: executes a list of tasks in memory-bounded slices
// ---------------------------------------------------------------------------------------------
namespace {

struct WRunner {
    Engine* eng;
    Backend* be;
    Prepared* p;
    std::vector<uint8_t>* opsPool = nullptr;

    size_t task_bytes(const WTask& t) const {
        size_t b = (size_t)p->ncodes * t.nWp * 4 + sizeof(WJob) + sizeof(Rec);
        if (t.flags & WF_STORE) b += (size_t)t.n * t.nWp * 8 + (size_t)t.m + t.n + 64;
        if (t.flags & WF_STOPCOL) b += (size_t)t.m * 4;
        if (!(t.flags & WF_SLIDE) && t.nWp / t.R > 32) b += 2 * (size_t)t.n;
        return b;
    }

    // Tasks whose query fits 256 rows and whose shape one of the lane-kernel classes covers run one
    // alignment per THREAD (lane_kernel); everything else one alignment per warp (w_kernel).
    static int lane_class(const WTask& t) {  // -1: not a lane task
        if (t.m > 256 || (t.flags & (WF_SLIDE | WF_STOPCOL))) return -1;
        const bool qrev = (t.flags & WF_QREV) != 0, trev = (t.flags & WF_TREV) != 0;
        if (t.flags & WF_STORE) return (t.mode == MODE_NW && !qrev && !trev) ? 4 : -1;
        if (qrev != trev) return -1;
        if (qrev) return t.mode == MODE_SHW ? 3 : -1;
        return t.mode;  // 0 NW, 1 SHW, 2 HW, forward
    }

    void run(std::vector<WTask>& tasks) {
        std::vector<int> warp;
        std::map<std::pair<int, int>, std::vector<int>> lanes;  // (word class, lane class) -> tasks
        for (size_t i = 0; i < tasks.size(); ++i) {
            const int lc = lane_class(tasks[i]);
            if (lc < 0) warp.push_back((int)i);
            else lanes[std::make_pair(ceil_div(tasks[i].m, 32), lc)].push_back((int)i);
        }
        for (auto& kv : lanes) {
            int bt = 0, rc = 0;
            be->k1_shape(kv.first.first, p->ncodes, 0x7fffffff, &bt, &rc);
            if (rc <= 0 || (int)kv.second.size() < 8) {  // alphabet too large for per-thread Peq rows / too few to bother
                warp.insert(warp.end(), kv.second.begin(), kv.second.end());
                continue;
            }
            run_lane(tasks, kv.second, kv.first.first, kv.first.second, warp);
        }
        std::vector<int>& order = warp;
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return tasks[a].R < tasks[b].R; });
        size_t i = 0;
        while (i < order.size()) {
            const int R = tasks[order[i]].R;
            size_t bytes = 0, j = i;
            while (j < order.size() && tasks[order[j]].R == R) {
                const size_t tb = task_bytes(tasks[order[j]]);
                if (j > i && bytes + tb > eng->tun.sliceBytes && tasks[order[j]].splitSide != 1) break;
                bytes += tb;
                ++j;
            }
            std::vector<int> slice(order.begin() + i, order.begin() + j);
            run_slice(tasks, slice, R, 0);
            i = j;
        }
    }

    // One class of lane tasks, in memory-bounded slices.  Tasks that need a longer end-location list
    // than a record holds are handed to the warp kernel (`spill`), which owns the list machinery.
    void run_lane(std::vector<WTask>& tasks, const std::vector<int>& idx, int nw, int lc, std::vector<int>& spill) {
        const bool store = lc == 4, rev = lc == 3;
        const int mode = store ? MODE_NW : rev ? MODE_SHW : lc;
        size_t i = 0;
        while (i < idx.size()) {
            size_t bytes = 0, j = i;
            while (j < idx.size()) {
                const WTask& t = tasks[idx[j]];
                const size_t tb = sizeof(LJob) + sizeof(Rec) + (store ? (size_t)t.n * nw * 8 + (size_t)t.m + t.n + 64 : 0);
                if (j > i && bytes + tb > eng->tun.sliceBytes) break;
                bytes += tb;
                ++j;
            }
            const int J = (int)(j - i);
            std::vector<LJob> jobs(J);
            std::vector<TbJob> tb;
            uint64_t matEntries = 0, opsBytes = 0;
            for (int s = 0; s < J; ++s) {
                const WTask& t = tasks[idx[i + s]];
                LJob& lj = jobs[s];
                memset(&lj, 0, sizeof(lj));
                lj.qOff = t.qOff;
                lj.tOff = t.tOff;
                lj.m = t.m;
                lj.n = t.n;
                lj.kInit = t.kInit;
                lj.trackFrom = t.trackFrom;
                if (store) {
                    lj.matOff = matEntries;
                    TbJob b;
                    memset(&b, 0, sizeof(b));
                    b.matOff = matEntries;
                    b.qOff = t.qOff;
                    b.peqOff = ~0ull;
                    b.tOff = t.tOff;
                    b.outOff = opsBytes;
                    b.m = t.m;
                    b.n = t.n;
                    b.nWp = nw;
                    tb.push_back(b);
                    matEntries += (uint64_t)t.n * nw;
                    opsBytes += (uint64_t)t.m + t.n;
                }
            }
            DevBuf<LJob> dJobs(be, J);
            dJobs.upload(jobs.data(), J);
            DevBuf<Rec> dRecs(be, J);
            be->zero(dRecs.p, (size_t)J * sizeof(Rec));
            DevBuf<U2> dMat(be, matEntries);
            LParams lp{dJobs.p, J, p->dSeq.p, p->dSeq.p, p->ncodes, p->hasEq ? p->dEqtab.p : nullptr, dRecs.p, dMat.p};
            be->launch_lane(lp, nw, mode, rev, store);
            DevBuf<TbJob> dTb;
            DevBuf<uint8_t> dOps;
            DevBuf<int> dOpsStart, dOpsLen;
            if (store) {
                dTb.alloc(be, tb.size());
                dTb.upload(tb.data(), tb.size());
                dOps.alloc(be, opsBytes);
                dOpsStart.alloc(be, tb.size());
                dOpsLen.alloc(be, tb.size());
                TbParams tp{dTb.p, (int)tb.size(), dMat.p, nullptr, p->dSeq.p, p->dSeq.p, p->hasEq ? p->dEqtab.p : nullptr, p->ncodes,
                            dOps.p, dOpsStart.p, dOpsLen.p};
                be->launch_traceback(tp);
            }
            std::vector<Rec> recs(J);
            dRecs.download(recs.data(), J);
            eng->stats.d2hBytes += (long long)J * (long long)sizeof(Rec);
            for (int s = 0; s < J; ++s) {
                WTask& t = tasks[idx[i + s]];
                t.rec = recs[s];
                t.extra.clear();
                if (t.wantPositions && t.rec.cnt > KPOS) spill.push_back(idx[i + s]);
            }
            if (store) {
                std::vector<int> st(tb.size()), ln(tb.size());
                dOpsStart.download(st.data(), tb.size());
                dOpsLen.download(ln.data(), tb.size());
                std::vector<uint8_t> ops(opsBytes);
                dOps.download(ops.data(), opsBytes);
                eng->stats.d2hBytes += (long long)opsBytes + 8LL * (long long)tb.size();
                for (int s = 0; s < J; ++s) {
                    WTask& t = tasks[idx[i + s]];
                    t.opsOff = (long long)opsPool->size();
                    t.opsLen = ln[s];
                    opsPool->insert(opsPool->end(), ops.begin() + tb[s].outOff + st[s], ops.begin() + tb[s].outOff + st[s] + ln[s]);
                }
            }
            i = j;
        }
    }

    // ovfCap == 0: first pass (no position list).  ovfCap > 0: second pass over the tasks whose
    // end-location lists exceed KPOS, started from their known minimum with an exact-size list.
    void run_slice(std::vector<WTask>& tasks, const std::vector<int>& slice, int R, int ovfCap) {
        const int J = (int)slice.size();
        std::vector<WJob> jobs(J);
        uint64_t peqWords = 0, matEntries = 0, colInts = 0, hbytes = 0, opsBytes = 0;
        std::vector<TbJob> tb;
        std::vector<int> tbTask;
        for (int s = 0; s < J; ++s) {
            WTask& t = tasks[slice[s]];
            WJob& j = jobs[s];
            memset(&j, 0, sizeof(j));
            j.qOff = t.qOff;
            j.tOff = t.tOff;
            j.m = t.m;
            j.n = t.n;
            j.nWp = t.nWp;
            j.mode = t.mode;
            j.flags = t.flags;
            j.kInit = t.kInit;
            j.dhi = t.dhi;
            j.stopCol = t.stopCol;
            j.trackFrom = t.trackFrom;
            j.rec = s;
            j.peqOff = peqWords;
            peqWords += (uint64_t)p->ncodes * t.nWp;
            if (t.flags & WF_STORE) {
                j.auxOff = matEntries;
                TbJob b;
                memset(&b, 0, sizeof(b));
                b.matOff = matEntries;
                b.qOff = t.qOff;
                b.peqOff = j.peqOff;
                b.tOff = t.tOff;
                b.outOff = opsBytes;
                b.m = t.m;
                b.n = t.n;
                b.nWp = t.nWp;
                tb.push_back(b);
                tbTask.push_back(slice[s]);
                matEntries += (uint64_t)t.n * t.nWp;
                opsBytes += (uint64_t)t.m + t.n;
            } else if (t.flags & WF_STOPCOL) {
                j.auxOff = colInts;
                colInts += (uint64_t)t.m;
            }
            if (!(t.flags & WF_SLIDE) && t.nWp / R > 32) {
                j.hbufOff = hbytes;
                hbytes += 2 * (uint64_t)t.n;
            }
        }
        DevBuf<WJob> dJobs(be, J);
        dJobs.upload(jobs.data(), J);
        DevBuf<uint32_t> dPeq(be, peqWords);
        DevBuf<U2> dMat(be, matEntries);
        DevBuf<int> dCol(be, colInts);
        DevBuf<uint8_t> dH(be, hbytes);
        DevBuf<Rec> dRecs(be, J);
        be->zero(dRecs.p, (size_t)J * sizeof(Rec));
        DevBuf<Ovf> dOvf(be, (size_t)std::max(ovfCap, 1));
        DevBuf<int> dOvfCount(be, 1);
        be->zero(dOvfCount.p, sizeof(int));
        if (colInts) be->fill(dCol.p, 0x3f, (size_t)colInts * sizeof(int));  // rows outside a sliding window: far above any k
        PeqParams pp{dJobs.p, J, p->dSeq.p, dPeq.p, p->ncodes, p->hasEq ? p->dEqtab.p : nullptr};
        be->launch_peq(pp);
        WParams wp{dJobs.p, J, p->dSeq.p, p->dSeq.p, dPeq.p, dH.p, dMat.p, dCol.p, dRecs.p, dOvf.p, dOvfCount.p, ovfCap};
        be->launch_w(wp, R);

        DevBuf<TbJob> dTb;
        DevBuf<uint8_t> dOps;
        DevBuf<int> dOpsStart, dOpsLen;
        if (!tb.empty()) {
            dTb.alloc(be, tb.size());
            dTb.upload(tb.data(), tb.size());
            dOps.alloc(be, opsBytes);
            dOpsStart.alloc(be, tb.size());
            dOpsLen.alloc(be, tb.size());
            TbParams tp{dTb.p, (int)tb.size(), dMat.p, dPeq.p, p->dSeq.p, p->dSeq.p, p->hasEq ? p->dEqtab.p : nullptr, p->ncodes,
                        dOps.p, dOpsStart.p, dOpsLen.p};
            be->launch_traceback(tp);
        }

        std::vector<Rec> recs(J);
        dRecs.download(recs.data(), J);
        int ovfCount = 0;
        dOvfCount.download(&ovfCount, 1);
        eng->stats.d2hBytes += (long long)J * (long long)sizeof(Rec) + 4;
        for (int s = 0; s < J; ++s) tasks[slice[s]].rec = recs[s];
        if (ovfCap > 0) {
            if (ovfCount > ovfCap) throw std::runtime_error("internal: end-location list larger than counted");
            std::vector<Ovf> ov(ovfCount);
            if (ovfCount) dOvf.download(ov.data(), ovfCount);
            eng->stats.d2hBytes += (long long)ovfCount * (long long)sizeof(Ovf);
            for (int s = 0; s < J; ++s) tasks[slice[s]].extra.clear();
            for (const Ovf& o : ov) {
                WTask& t = tasks[slice[o.rec]];
                if (o.score == t.rec.best) t.extra.push_back(o.pos);
            }
        }
        if (!tb.empty()) {
            std::vector<int> st(tb.size()), ln(tb.size());
            dOpsStart.download(st.data(), tb.size());
            dOpsLen.download(ln.data(), tb.size());
            std::vector<uint8_t> ops(opsBytes);
            dOps.download(ops.data(), opsBytes);
            eng->stats.d2hBytes += (long long)opsBytes + 8LL * (long long)tb.size();
            for (size_t k = 0; k < tb.size(); ++k) {
                WTask& t = tasks[tbTask[k]];
                t.opsOff = (long long)opsPool->size();
                t.opsLen = ln[k];
                opsPool->insert(opsPool->end(), ops.begin() + tb[k].outOff + st[k], ops.begin() + tb[k].outOff + st[k] + ln[k]);
            }
        }
        if (colInts) {
            // Hirschberg halves: the split row is searched on the device, only {h, left, right} come back.
            std::vector<SplitNode> nodes;
            std::vector<int> owner;
            for (int s = 0; s + 1 < J; ++s) {
                const WTask& f = tasks[slice[s]];
                const WTask& r = tasks[slice[s + 1]];
                if (f.splitSide != 0 || r.splitSide != 1) continue;
                SplitNode nd;
                nd.colF = jobs[s].auxOff;
                nd.colR = jobs[s + 1].auxOff;
                nd.m = f.m;
                nd.leftW = f.n;
                nd.rightW = r.n;
                nd.best = f.splitBest;
                nodes.push_back(nd);
                owner.push_back(slice[s]);
            }
            if (nodes.empty()) throw std::runtime_error("internal: stop-column tasks without a split pair");
            DevBuf<SplitNode> dNodes(be, nodes.size());
            dNodes.upload(nodes.data(), nodes.size());
            DevBuf<SplitOut> dOut(be, nodes.size());
            SplitParams sp{dNodes.p, (int)nodes.size(), dCol.p, dOut.p};
            be->launch_split(sp);
            std::vector<SplitOut> outs(nodes.size());
            dOut.download(outs.data(), outs.size());
            eng->stats.d2hBytes += (long long)outs.size() * (long long)sizeof(SplitOut);
            for (size_t q = 0; q < outs.size(); ++q) tasks[owner[q]].split = outs[q];
        }
        if (ovfCap == 0) {
            std::vector<int> again;
            long long need = 0;
            for (int s = 0; s < J; ++s) {
                WTask& t = tasks[slice[s]];
                if (t.wantPositions && t.rec.cnt > KPOS) {
                    again.push_back(slice[s]);
                    need += t.rec.cnt - KPOS;
                    t.kInit = t.rec.best;
                }
            }
            if (need > 0x7fffffffLL / 4) throw std::runtime_error("end-location list too large");
            if (!again.empty()) run_slice(tasks, again, R, (int)need + 16);
        }
    }
};

}  // namespace

// ---------------------------------------------------------------------------------------------
// compute
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// One compute() over a prepared batch: shared state + the phases of the reference driver
// ---------------------------------------------------------------------------------------------
namespace {

struct Pass {
    Engine& eng;
    Backend* be;
    Prepared* p;
    EngineTunables& tun;
    EngineStats& stats;
    Trace trace;
    const int N, mode, k;
    // per-pair sweep outcome before the "-1" rule (storage reused from pass to pass: EngineScratch)
    std::vector<int>&best, &cnt;
    std::vector<long long>& posStart;  // end columns of pair i: posPool[posStart[i] .. +posLen[i])
    std::vector<int>&posLen, &posPool;
    std::vector<int> wPairs;          // pairs swept by the warp / lane-job kernels
    std::vector<uint8_t> opsPool;
    WRunner runner;
    int laneOkCache[9] = {-1, -1, -1, -1, -1, -1, -1, -1, -1};

    Pass(Engine& e, Backend* b, Prepared* pr)
        : eng(e), be(b), p(pr), tun(e.tun), stats(e.stats), N(pr->N), mode(pr->mode), k(pr->cfg.k),
          best(e.scratch.best), cnt(e.scratch.cnt), posStart(e.scratch.posStart), posLen(e.scratch.posLen),
          posPool(e.scratch.posPool), runner{&e, b, pr, &opsPool} {
        best.resize((size_t)N);
        cnt.resize((size_t)N);
        posStart.resize((size_t)N);
        posLen.resize((size_t)N);
        parallel_ranges((size_t)N, 65536, [this](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; ++i) {
                best[i] = -1;
                cnt[i] = 0;
                posStart[i] = -1;
                posLen[i] = 0;
            }
        });
        posPool.clear();
        posPool.reserve((size_t)N + 16);
    }

    // ---- direct lane-kernel launches (no per-job host objects): the LOC / PATH phases of large read
    // batches issue millions of tiny sweeps, so their jobs are built straight into LJob arrays. --------
    bool lane_ok(int m) {
        if (m <= 0 || m > 256) return false;
        const int nw = ceil_div(m, 32);
        if (laneOkCache[nw] < 0) {
            int bt = 0, rc = 0;
            be->k1_shape(nw, p->ncodes, 0x7fffffff, &bt, &rc);
            laneOkCache[nw] = rc > 0 ? 1 : 0;
        }
        return laneOkCache[nw] == 1;
    }

    void lane_launch(const std::vector<LJob>& jobs, int nw, int laneMode, bool rev, std::vector<Rec>& recs) {
        const size_t J = jobs.size();
        recs.resize(J);
        const size_t step = 4u << 20;
        for (size_t a = 0; a < J; a += step) {
            const size_t n = std::min(step, J - a);
            DevBuf<LJob> dJobs(be, n);
            dJobs.upload(jobs.data() + a, n);
            DevBuf<Rec> dRecs(be, n);
            be->zero(dRecs.p, n * sizeof(Rec));
            LParams lp{dJobs.p, (int)n, p->dSeq.p, p->dSeq.p, p->ncodes, p->hasEq ? p->dEqtab.p : nullptr, dRecs.p, nullptr};
            be->launch_lane(lp, nw, laneMode, rev, false);
            dRecs.download(recs.data() + a, n);
            stats.d2hBytes += (long long)n * (long long)sizeof(Rec);
        }
    }

    // Matrix-storing NW sweeps + traceback of `jobs` (matOff is assigned here); `sink(jobIndex, ops, len,
    // score)` receives every edit script.  Slices bound the stored matrices to the slice budget.
    template <class Sink>
    void lane_paths(std::vector<LJob>& jobs, int nw, Sink sink) {
        size_t a = 0;
        while (a < jobs.size()) {
            size_t bytes = 0, b = a;
            uint64_t matEntries = 0, opsBytes = 0;
            std::vector<TbJob> tb;
            while (b < jobs.size()) {
                LJob& j = jobs[b];
                const size_t need = (size_t)j.n * nw * 8 + (size_t)j.m + j.n + sizeof(LJob) + sizeof(TbJob) + 64;
                if (b > a && bytes + need > tun.sliceBytes) break;
                j.matOff = matEntries;
                TbJob t;
                memset(&t, 0, sizeof(t));
                t.matOff = matEntries;
                t.qOff = j.qOff;
                t.peqOff = ~0ull;
                t.tOff = j.tOff;
                t.outOff = opsBytes;
                t.m = j.m;
                t.n = j.n;
                t.nWp = nw;
                tb.push_back(t);
                matEntries += (uint64_t)j.n * nw;
                opsBytes += (uint64_t)j.m + j.n;
                bytes += need;
                ++b;
            }
            const size_t n = b - a;
            DevBuf<LJob> dJobs(be, n);
            dJobs.upload(jobs.data() + a, n);
            DevBuf<Rec> dRecs(be, n);
            be->zero(dRecs.p, n * sizeof(Rec));
            DevBuf<U2> dMat(be, matEntries);
            LParams lp{dJobs.p, (int)n, p->dSeq.p, p->dSeq.p, p->ncodes, p->hasEq ? p->dEqtab.p : nullptr, dRecs.p, dMat.p};
            be->launch_lane(lp, nw, MODE_NW, false, true);
            DevBuf<TbJob> dTb(be, n);
            dTb.upload(tb.data(), n);
            DevBuf<uint8_t> dOps(be, opsBytes);
            DevBuf<int> dStart(be, n), dLen(be, n);
            TbParams tp{dTb.p, (int)n, dMat.p, nullptr, p->dSeq.p, p->dSeq.p, p->hasEq ? p->dEqtab.p : nullptr, p->ncodes,
                        dOps.p, dStart.p, dLen.p};
            be->launch_traceback(tp);
            // one pinned staging block for everything that comes back (fast D2H, no zero-fill of vectors)
            const size_t offSt = round_up(n * sizeof(Rec), 64), offLn = offSt + round_up(n * sizeof(int), 64);
            const size_t offOps = offLn + round_up(n * sizeof(int), 64);
            uint8_t* host = static_cast<uint8_t*>(be->alloc_host(offOps + opsBytes));
            const Rec* recs = reinterpret_cast<const Rec*>(host);
            const int* st = reinterpret_cast<const int*>(host + offSt);
            const int* ln = reinterpret_cast<const int*>(host + offLn);
            const uint8_t* ops = host + offOps;
            be->d2h(host, dRecs.p, n * sizeof(Rec));
            be->d2h(host + offSt, dStart.p, n * sizeof(int));
            be->d2h(host + offLn, dLen.p, n * sizeof(int));
            be->d2h(host + offOps, dOps.p, opsBytes);
            stats.d2hBytes += (long long)opsBytes + (long long)n * (long long)(sizeof(Rec) + 8);
            for (size_t q = 0; q < n; ++q) sink(a + q, ops + tb[q].outOff + st[q], ln[q], recs[q].best);
            be->free_host(host);
            a = b;
        }
    }

    // Hash indexes of the seeds of one target (candidate filter, seed stages), one per seed length; kept for
    // the last target used.  Level 0: the shortest L with sigma^L >= filterSeedSlack * n (a fraction of a chance
    // occurrence per seed: every occurrence costs a window sweep);
    // levels 1 and 2: two and four symbols shorter (more seeds fit into a read, so a higher threshold, at the
    // price of more chance occurrences) for the reads the previous level cannot decide.
    struct SeedIndex {
        int target = -1;
        int L = 0, bits = 0;
        DevBuf<int> bucketStart, positions;
    } seed[SEED_LEVELS];
    bool seed_index(int t, int level) {
        SeedIndex& sx = seed[level];
        const Target& tg = p->tg[t];
        const int n = tg.len;
        if (sx.target == t) return sx.L > 0;
        sx.target = t;
        sx.L = 0;
        const double sigma = std::max(2, p->ncodes);
        int L = 8;
        double v = std::pow(sigma, 8);
        while (v < (double)tun.filterSeedSlack * (double)n && L < 32) {
            v *= sigma;
            ++L;
        }
        if (level > 0) {
            if (L - 2 * level < 8) return false;  // seeds shorter than 8 symbols select nothing
            L -= 2 * level;
        }
        if (n < 4 * L) return false;
        int bits = 12;
        while (bits < 27 && (1LL << bits) < 2LL * n) ++bits;
        const size_t B = (size_t)1 << bits;
        sx.bucketStart.alloc(be, B + 1);
        sx.positions.alloc(be, (size_t)(n - L + 1));
        DevBuf<int> cursor(be, B);
        be->zero(sx.bucketStart.p, (B + 1) * sizeof(int));
        be->zero(cursor.p, B * sizeof(int));
        SeedIndexParams ip;
        memset(&ip, 0, sizeof(ip));
        ip.tcodes = p->dSeq.p + tg.off;
        ip.n = n;
        ip.L = L;
        ip.bits = bits;
        ip.bucketStart = sx.bucketStart.p;
        ip.cursor = cursor.p;
        ip.positions = sx.positions.p;
        be->launch_seed_count(ip);
        be->launch_scan(sx.bucketStart.p, (int)B);
        be->launch_seed_fill(ip);
        sx.L = L;
        sx.bits = bits;
        trace.mark("filter: seed index");
        return true;
    }

    // One group of pairs that share a target and a word class (queries <= 256 rows), on its way through
    // the distance pass.  Reads are addressed by their index `s` into `list`.
    struct LaneGroup {
        int t, nw;               // target index, 32-bit words per query
        const std::vector<int>& list;  // the pairs of the group
        const Target& tg;
        int n;                   // target length
        std::vector<int> bound;  // per read: largest distance that still counts as found
        std::vector<int> excl;   // per read: it is known that no distance <= excl[s] exists
        std::vector<int> direct; // reads that take the plain full sweep
    };

    // Chunk geometry: a HW sweep may be cut into target chunks (each re-started 2*m columns
    // early, exact because no HW path spans more than 2*m target symbols) so that a small
    // group still fills the machine.
    void lane_geometry(const LaneGroup& c, int g, int nwL, int& chunks, int& chunkLen, bool perChunkRecs) {
        const int n = c.n;
        int blockThreads = 256, residentCtas = 1;
        be->k1_shape(nwL, p->ncodes, g, &blockThreads, &residentCtas);
        chunks = 1;
        chunkLen = (int)round_up((size_t)n, 16);
        if (mode != MODE_HW) return;
        // the restart lead-in (64 * nwL columns) stays below 1/8 of a chunk; a handful of reads is latency-bound
        // per CTA and may be cut finer (lead-in up to 1/3)
        const int minChunk = std::max(tun.k1MinChunk, (g <= 32 ? 2 : 8) * 64 * nwL);
        long long maxChunks = std::max<long long>(1, n / minChunk);
        // plain sweeps return one record per (chunk, read): keep that below ~64 MB
        if (perChunkRecs) maxChunks = std::min<long long>(maxChunks, std::max<long long>(64, (2LL << 20) / std::max(g, 1)));
        maxChunks = std::min<long long>(maxChunks, 4096);
        const long long tiles = ceil_div(g, blockThreads);
        // CTAs run in waves of `residentCtas`; all CTAs of a launch cost the same, so the launch
        // takes ceil(waves) CTA-times.  Pick the cut with the best (fullness of the last wave) x
        // (1 - halo overhead); more, shorter CTAs fill waves better.
        long long best = 1;
        double bestScore = -1;
        for (long long c = 1; c <= maxChunks; ++c) {
            const double waves = (double)(tiles * c) / residentCtas;
            const double eff = waves / (double)((tiles * c + residentCtas - 1) / residentCtas);
            const double len = (double)n / (double)c;
            const double score = eff * (len / (len + 64.0 * nwL));
            if (score > bestScore + 0.002) {
                bestScore = score;
                best = c;
            }
        }
        chunkLen = (int)round_up((size_t)ceil_div(n, (int)best), 16);
        chunks = ceil_div(n, chunkLen);
    }

    // One launch over the reads `sub` (indices into `list`) with sentinels / thresholds subK.
    void lane_sweep(LaneGroup& c, const std::vector<int>& sub, const std::vector<int>& subK, int nwL, int chunks, int chunkLen,
                    int cap, int prefixLen, int rangeMode, std::vector<Rec>& outRecs, std::vector<Ovf>& outOvf) {
        const std::vector<int>& list = c.list;
        const Target& tg = c.tg;
        const int n = c.n;
        const int g = (int)sub.size();
        std::vector<int> rl(g);
        for (int s = 0; s < g; ++s) rl[s] = list[sub[s]];
        DevBuf<int> dList(be, g), dK(be, g);
        dList.upload(rl.data(), g);
        dK.upload(subK.data(), g);
        const size_t numRecs = rangeMode ? 0 : (size_t)g * chunks;  // range mode reports through the list only
        DevBuf<Rec> dRecs(be, std::max<size_t>(numRecs, 1));
        if (numRecs) be->zero(dRecs.p, numRecs * sizeof(Rec));
        DevBuf<int> dCount(be, 1);
        K1Params kp;
        memset(&kp, 0, sizeof(kp));
        kp.tcodes = p->dSeq.p + tg.off;
        kp.n = n;
        kp.qcodes = p->dSeq.p;
        kp.qoff = p->dQoff.p;
        kp.qlen = p->dQlen.p;
        kp.readList = dList.p;
        kp.kInit = dK.p;
        kp.numReads = g;
        kp.mode = mode;
        kp.ncodes = p->ncodes;
        kp.eqtab = p->hasEq ? p->dEqtab.p : nullptr;
        kp.chunks = chunks;
        kp.chunkLen = chunkLen;
        kp.halo = 64 * nwL;
        kp.recs = dRecs.p;
        kp.ovfCount = dCount.p;
        kp.prefixLen = prefixLen;
        kp.rangeMode = rangeMode;
        for (;;) {
            DevBuf<Ovf> dOvf(be, (size_t)std::max(cap, 1));
            be->zero(dCount.p, sizeof(int));
            kp.ovf = dOvf.p;
            kp.ovfCap = cap;
            be->launch_k1(kp, nwL);
            outOvf.clear();
            if (cap <= 0) break;
            int count = 0;
            dCount.download(&count, 1);
            stats.d2hBytes += 4;
            if (count > cap) {
                if (!rangeMode) throw std::runtime_error("internal: end-location list larger than counted");
                cap = count;  // range list overflow: repeat with the exact size
                continue;
            }
            outOvf.resize(count);
            if (count) dOvf.download(outOvf.data(), count);
            stats.d2hBytes += (long long)count * (long long)sizeof(Ovf);
            break;
        }
        outRecs.resize(numRecs);
        if (numRecs) dRecs.download(outRecs.data(), outRecs.size());
        stats.d2hBytes += (long long)outRecs.size() * (long long)sizeof(Rec);
    }

    // Merge the chunks of every read: the minimum wins; its columns are the inline positions
    // of the chunks attaining it (ascending by construction) plus, in a second pass, the
    // listed ones.  Returns the reads whose lists are incomplete (some chunk holds > KPOS).
    void lane_merge(LaneGroup& c, const std::vector<int>& sub, int chunks, const std::vector<Rec>& rr, const std::vector<Ovf>* oo,
                    std::vector<int>& incomplete, long long& missing) {
        const std::vector<int>& list = c.list;
        const int g = (int)sub.size();
        std::unordered_map<int, std::vector<int>> extra;  // rec index -> listed positions
        if (oo)
            for (const Ovf& o : *oo)
                if (o.score == rr[o.rec].best) extra[o.rec].push_back(o.pos);
        for (int s = 0; s < g; ++s) {
            const int pair = list[sub[s]];
            int b = 0x7fffffff;
            long long total = 0;
            for (int c = 0; c < chunks; ++c) {
                const Rec& r = rr[(size_t)c * g + s];
                if (r.cnt > 0 && r.best < b) {
                    b = r.best;
                    total = 0;
                }
                if (r.cnt > 0 && r.best == b) total += r.cnt;
            }
            best[pair] = (total > 0) ? b : 0x7fffffff;
            if (total > 0x7fffffffLL / 4) throw std::runtime_error("end-location list too large");
            cnt[pair] = (int)total;
            std::vector<int>& dst = posPool;
            posStart[pair] = (long long)posPool.size();
            posLen[pair] = 0;
            if (total == 0) continue;
            bool complete = true;
            for (int c = 0; c < chunks; ++c) {
                const Rec& r = rr[(size_t)c * g + s];
                if (r.cnt <= 0 || r.best != b) continue;
                for (int q = 0; q < std::min(r.cnt, KPOS); ++q) dst.push_back(r.pos[q]);
                if (r.cnt > KPOS) {
                    if (oo) {
                        const std::vector<int>& ex = extra[(int)((size_t)c * g + s)];
                        dst.insert(dst.end(), ex.begin(), ex.end());
                    } else {
                        complete = false;
                    }
                }
            }
            posLen[pair] = (int)((long long)posPool.size() - posStart[pair]);
            if (!complete) {
                incomplete.push_back(sub[s]);
                missing += total;
            }
        }
    }

    // It is now known that read s has no alignment within t: final if t is the caller's bound, else the
    // read moves on to `next`.
    void no_distance_within(LaneGroup& c, int s, int t, std::vector<int>& next) {
        if (t > c.excl[s]) c.excl[s] = t;
        if (t == c.bound[s]) {
            best[c.list[s]] = 0x7fffffff;
            cnt[c.list[s]] = 0;
            stats.filterDecided++;
        } else {
            next.push_back(s);
        }
    }

    // Seed stage: exact seeds of every read looked up in the hash index of the target; windows around
    // the expected end columns are planned, swept and reduced on the device (eb_core.h: seed_plan_read).
    void seed_stage(LaneGroup& c, int level, const std::vector<int>& in, std::vector<int>& next) {
        const std::vector<int>& list = c.list;
        const Target& tg = c.tg;
        const int n = c.n;
        const int nw = c.nw;
        const std::vector<int>& bound = c.bound;
        std::vector<int>& excl = c.excl;
        std::vector<int>& direct = c.direct;
        if (!seed_index(c.t, level)) {
            next = in;
            return;
        }
        const SeedIndex& sx = seed[level];
        const int L = sx.L;
        // every read of `in` gets a slot; thr < 0 marks the ones this stage cannot help (the kernel skips them)
        const std::vector<int>& cand = in;
        const int g = (int)cand.size();
        if (g == 0) return;
        HostBuf<int> rl(be, g), hThr(be, g);
        const int* thr = hThr.p;
        parallel_ranges((size_t)g, 65536, [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; ++i) {
                const int s = cand[i];
                const int m = p->qlen[list[s]];
                const int tt = std::min(std::min(bound[s], tun.filterSeedK), m / L - 1);
                rl[i] = list[s];
                hThr[i] = (m >= 2 * L && tt > excl[s]) ? tt : -1;
            }
        });
        DevBuf<int> dList(be, g), dThr(be, g), dCount(be, 1);
        dList.upload(rl.p, g);
        dThr.upload(hThr.p, g);
        DevBuf<SeedPlan> dPlan(be, g);
        DevBuf<int> wPair, wK, wStart, wLen, wTf;
        // room for the window jobs: sized from what the previous pass of this level needed per read
        int& perRead = eng.scratch.seedWindowsPerRead[level];
        int cap = (int)std::min<long long>((long long)g * std::max(perRead + 2, level == 0 ? 8 : level == 1 ? 96 : 400) + 4096, 1LL << 28), V = 0;
        for (;;) {
            wPair.alloc(be, cap);
            wK.alloc(be, cap);
            wStart.alloc(be, cap);
            wLen.alloc(be, cap);
            wTf.alloc(be, cap);
            be->zero(dCount.p, sizeof(int));
            SeedPlanParams sp;
            memset(&sp, 0, sizeof(sp));
            sp.tcodes = p->dSeq.p + tg.off;
            sp.n = n;
            sp.qcodes = p->dSeq.p;
            sp.qoff = p->dQoff.p;
            sp.qlen = p->dQlen.p;
            sp.readList = dList.p;
            sp.thr = dThr.p;
            sp.numReads = g;
            sp.L = L;
            sp.bits = sx.bits;
            sp.bucketStart = sx.bucketStart.p;
            sp.positions = sx.positions.p;
            sp.maxBucket = tun.filterSeedBucket << (4 * level);  // shorter seeds: longer buckets are normal
            sp.level = level;
            sp.spread = tun.filterSpread;
            sp.winPair = wPair.p;
            sp.winK = wK.p;
            sp.winStart = wStart.p;
            sp.winLen = wLen.p;
            sp.winTf = wTf.p;
            sp.winCap = cap;
            sp.winCount = dCount.p;
            sp.plan = dPlan.p;
            be->launch_seed_plan(sp);
            dCount.download(&V, 1);
            stats.d2hBytes += 4;
            if (V <= cap) break;
            cap = V;  // window list overflow: repeat with the exact size
        }
        perRead = (int)(((long long)V + g - 1) / g);
        stats.filterWindows += V;
        trace.mark("filter: seeds planned");
        DevBuf<WinRec> dWinRecs(be, (size_t)std::max(V, 1));
        if (V > 0) {
            K1WParams wp;
            memset(&wp, 0, sizeof(wp));
            wp.tcodes = p->dSeq.p + tg.off;
            wp.qcodes = p->dSeq.p;
            wp.qoff = p->dQoff.p;
            wp.qlen = p->dQlen.p;
            wp.readList = wPair.p;
            wp.kInit = wK.p;
            wp.winStart = wStart.p;
            wp.winLen = wLen.p;
            wp.trackFrom = wTf.p;
            wp.numReads = V;
            wp.ncodes = p->ncodes;
            wp.eqtab = nullptr;
            wp.recs = dWinRecs.p;
            be->launch_k1w(wp, nw);
        }
        DevBuf<Rec> dOut(be, g);
        const int extraCap = g / 4 + 1024;
        DevBuf<int> dExtra(be, (size_t)extraCap);
        be->zero(dCount.p, sizeof(int));
        WinReduceParams rp;
        rp.plan = dPlan.p;
        rp.thr = dThr.p;
        rp.winRecs = dWinRecs.p;
        rp.numReads = g;
        rp.out = dOut.p;
        rp.extra = dExtra.p;
        rp.extraCount = dCount.p;
        rp.extraCap = extraCap;
        be->launch_win_reduce(rp);
        HostBuf<Rec> out(be, g);
        dOut.download(out.p, g);
        int nExtra = 0;
        dCount.download(&nExtra, 1);
        nExtra = std::min(nExtra, extraCap);  // reads whose run did not fit were marked as long lists
        std::vector<int> extra((size_t)nExtra);
        if (nExtra) dExtra.download(extra.data(), (size_t)nExtra);
        stats.d2hBytes += (long long)g * (long long)sizeof(Rec) + 4 + 4LL * nExtra;
        trace.mark("filter: seed windows");
        // Outcome per read, on a few host threads: records of decided reads go straight to best / cnt;
        // their positions are appended to posPool in slot order (counts first, then the fill).
        struct Part {
            std::vector<int> next, direct;
            long long decided = 0, positions = 0;
            int nSat = 0, nLong = 0;
        };
        std::vector<Part> parts;
        std::vector<size_t> partLo;
        {
            const size_t nparts = host_parts((size_t)g, 65536);
            parts.resize(nparts);
            for (size_t t2 = 0; t2 <= nparts; ++t2) partLo.push_back((size_t)g * t2 / nparts);
        }
        auto for_parts = [&](const std::function<void(size_t)>& fn) { HostPool::get().run(parts.size(), fn); };
        for_parts([&](size_t t2) {
            Part& P = parts[t2];
            for (size_t i = partLo[t2]; i < partLo[t2 + 1]; ++i) {
                const int s = cand[i], pair = list[s];
                const Rec& r = out[i];
                if (thr[i] < 0) {
                    P.next.push_back(s);
                } else if (r.rsv == SEED_WINDOWS) {
                    P.decided++;
                    best[pair] = r.best;
                    cnt[pair] = r.cnt;
                    posLen[pair] = r.cnt;
                    P.positions += r.cnt;
                } else if (r.rsv == SEED_NONE) {
                    if (thr[i] > excl[s]) excl[s] = thr[i];
                    if (thr[i] == bound[s]) {  // nothing within the caller's bound: final
                        best[pair] = 0x7fffffff;
                        cnt[pair] = 0;
                        P.decided++;
                    } else {
                        P.next.push_back(s);
                    }
                } else if (r.rsv == SEED_LONG_LIST) {
                    P.direct.push_back(s);
                    P.nLong++;
                } else {
                    P.next.push_back(s);
                    P.nSat++;
                }
            }
        });
        int nSat = 0, nLong = 0;
        std::vector<long long> partPos(parts.size());
        {
            long long at = (long long)posPool.size();
            for (size_t t2 = 0; t2 < parts.size(); ++t2) {
                partPos[t2] = at;
                at += parts[t2].positions;
                stats.filterDecided += parts[t2].decided;
                nSat += parts[t2].nSat;
                nLong += parts[t2].nLong;
                next.insert(next.end(), parts[t2].next.begin(), parts[t2].next.end());
                direct.insert(direct.end(), parts[t2].direct.begin(), parts[t2].direct.end());
            }
            posPool.resize((size_t)at);
        }
        for_parts([&](size_t t2) {
            long long at = partPos[t2];
            for (size_t i = partLo[t2]; i < partLo[t2 + 1]; ++i) {
                const Rec& r = out[i];
                if (thr[i] < 0 || r.rsv != SEED_WINDOWS) continue;
                const int pair = list[cand[i]];
                posStart[pair] = at;
                for (int q = 0; q < std::min(r.cnt, KPOS); ++q) posPool[(size_t)at++] = r.pos[q];
                for (int q = KPOS; q < r.cnt; ++q) posPool[(size_t)at++] = extra[(size_t)r.last + q - KPOS];
            }
        });
        if (trace.on)
            fprintf(stderr, "[edlib_b200] filter seed stage %d, L=%d: %d reads, %d windows, %d saturated, %d long lists, %zu to the next stage\n",
                    level, L, g, V, nSat, nLong, next.size());
    }

    // Prefix stage over the reads `in` (indices into `list`): a sweep of the first P rows of every read reports
    // the target ranges where that prefix matches within t = min(K0, bound); the whole read is then swept over
    // one window per range.  A read is decided when a window holds a distance <= t (or when t is the caller's
    // bound and none does).  Undecided reads go to `next` (a longer prefix or the plain sweep), reads with
    // long end-location lists to c.direct.
    void prefix_stage(LaneGroup& c, int P, int K0, const std::vector<int>& in, std::vector<int>& next) {
        const std::vector<int>& list = c.list;
        const Target& tg = c.tg;
        const int n = c.n;
        const int nw = c.nw;
        const std::vector<int>& bound = c.bound;
        std::vector<int>& excl = c.excl;
        std::vector<int>& direct = c.direct;
        std::vector<int> cand, thr;
        const int minLen = std::max(tun.filterMinLen * P / 64, P + 1);
        for (int s : in) {
            // worth a sweep only if it can decide clearly more than what is already excluded
            if (p->qlen[list[s]] >= minLen && std::min(K0, bound[s]) > excl[s] && (excl[s] < 0 || K0 >= excl[s] + 4)) {
                cand.push_back(s);
                thr.push_back(std::min(K0, bound[s]));
            } else {
                next.push_back(s);
            }
        }
        if (cand.empty()) return;
        const int g = (int)cand.size();
        int chunksA = 1, chunkLenA = 0;
        lane_geometry(c, g, P / 32, chunksA, chunkLenA, false);
        std::vector<Rec> none;
        std::vector<Ovf> ranges;
        lane_sweep(c, cand, thr, P / 32, chunksA, chunkLenA, (int)std::min<long long>(16LL * g + 4096, 1LL << 28), P, 1, none, ranges);
        trace.mark("filter: prefix sweep");
        auto undecided = [&](int s, int t) { no_distance_within(c, s, t, next); };
        // ranges of every read, ascending (the list is in completion order)
        std::vector<int> start(g + 1, 0);
        std::vector<char> saturated(g, 0);
        for (const Ovf& o : ranges) {
            if (o.score < 0) saturated[o.rec] = 1;
            else start[o.rec + 1]++;
        }
        for (int i = 0; i < g; ++i) start[i + 1] += start[i];
        std::vector<std::pair<int, int>> rg(start[g]);
        {
            std::vector<int> fill(start.begin(), start.end() - 1);
            for (const Ovf& o : ranges)
                if (o.score >= 0) rg[fill[o.rec]++] = std::make_pair(o.score, o.pos);
        }
        std::vector<int> vOwner, vPair, vK, vWs, vLen, vTf;  // windows to verify
        std::vector<int> wFirst(g + 1, 0);
        for (int i = 0; i < g; ++i) {
            wFirst[i] = (int)vOwner.size();
            const int s = cand[i];
            const int pair = list[s], m = p->qlen[pair], t = thr[i];
            if (saturated[i]) {
                next.push_back(s);
                continue;
            }
            if (start[i] == start[i + 1]) {
                undecided(s, t);
                continue;
            }
            std::sort(rg.begin() + start[i], rg.begin() + start[i + 1]);
            // An alignment with distance d <= t ending at column e passes, after its first P rows,
            // through a column c' with prefix score <= d and e in [c'+(m-P)-d, c'+(m-P)+d]: the end
            // columns to examine are [first+(m-P)-t, last+(m-P)+t] of every range.  Ranges close to
            // each other share one window; tracked columns of successive windows are kept disjoint.
            long long prevHi = -1;
            int windows = 0;
            for (int a = start[i]; a < start[i + 1];) {
                const int first = rg[a].first;
                int last = rg[a].second;
                int b = a + 1;
                while (b < start[i + 1] && rg[b].first - last <= K1_RANGE_GAP && rg[b].second - first <= tun.filterSpread) {
                    last = std::max(last, rg[b].second);
                    ++b;
                }
                a = b;
                long long lo = (long long)first + (m - P) - t;
                long long hi = (long long)last + (m - P) + t;
                if (lo <= prevHi) lo = prevHi + 1;
                if (lo < 0) lo = 0;
                if (hi > n - 1) hi = n - 1;
                if (lo > hi) continue;
                prevHi = hi;
                // HW restart: alignments with <= t edits span at most m + t columns (scores <= t stay exact)
                const long long ws = std::max<long long>(0, lo - (long long)(m + t));
                vOwner.push_back(i);
                vPair.push_back(pair);
                vK.push_back(t + 1);
                vWs.push_back((int)ws);
                vLen.push_back((int)(hi - ws + 1));
                vTf.push_back((int)(lo - ws));
                ++windows;
            }
            if (windows == 0) {
                undecided(s, t);
            } else if (windows > tun.filterMaxWindows) {
                vOwner.resize(wFirst[i]);
                vPair.resize(wFirst[i]);
                vK.resize(wFirst[i]);
                vWs.resize(wFirst[i]);
                vLen.resize(wFirst[i]);
                vTf.resize(wFirst[i]);
                next.push_back(s);
            }
        }
        wFirst[g] = (int)vOwner.size();
        trace.mark("filter: windows planned");
        const int V = (int)vOwner.size();
        stats.filterWindows += V;
        if (trace.on) {
            int sat = 0;
            for (char c : saturated) sat += c;
            fprintf(stderr, "[edlib_b200] filter stage P=%d: %d reads, %zu ranges, %d saturated, %d windows, %zu to the next stage\n",
                    P, g, rg.size(), sat, V, next.size());
        }
        if (V == 0) return;
        // Whole reads over their windows, one window per thread (k1w_kernel).
        DevBuf<int> dPair(be, V), dK(be, V), dWs(be, V), dLen(be, V), dTf(be, V);
        dPair.upload(vPair.data(), V);
        dK.upload(vK.data(), V);
        dWs.upload(vWs.data(), V);
        dLen.upload(vLen.data(), V);
        dTf.upload(vTf.data(), V);
        DevBuf<WinRec> dRecs(be, V);
        K1WParams wp;
        memset(&wp, 0, sizeof(wp));
        wp.tcodes = p->dSeq.p + tg.off;
        wp.qcodes = p->dSeq.p;
        wp.qoff = p->dQoff.p;
        wp.qlen = p->dQlen.p;
        wp.readList = dPair.p;
        wp.kInit = dK.p;
        wp.winStart = dWs.p;
        wp.winLen = dLen.p;
        wp.trackFrom = dTf.p;
        wp.numReads = V;
        wp.ncodes = p->ncodes;
        wp.eqtab = p->hasEq ? p->dEqtab.p : nullptr;
        wp.recs = dRecs.p;
        be->launch_k1w(wp, nw);
        std::vector<WinRec> rv(V);
        dRecs.download(rv.data(), V);
        stats.d2hBytes += (long long)V * (long long)sizeof(WinRec);
        trace.mark("filter: window sweeps");
        for (int i = 0; i < g; ++i) {
            if (wFirst[i] == wFirst[i + 1]) continue;
            const int s = cand[i], t = thr[i], pair = list[s];
            int b = 0x7fffffff;
            for (int j = wFirst[i]; j < wFirst[i + 1]; ++j)
                if (rv[j].cnt > 0 && rv[j].best < b) b = rv[j].best;
            if (b > t) {  // every window minimum is above the threshold
                undecided(s, t);
                continue;
            }
            bool longList = false;
            int total = 0;
            for (int j = wFirst[i]; j < wFirst[i + 1]; ++j)
                if (rv[j].cnt > 0 && rv[j].best == b) {
                    total += rv[j].cnt;
                    if (rv[j].cnt > KPOSW) longList = true;
                }
            if (longList) {  // long end-location list: the plain sweep collects it
                direct.push_back(s);
                continue;
            }
            stats.filterDecided++;
            best[pair] = b;
            cnt[pair] = total;
            posStart[pair] = (long long)posPool.size();
            for (int j = wFirst[i]; j < wFirst[i + 1]; ++j)
                if (rv[j].cnt > 0 && rv[j].best == b)
                    for (int q = 0; q < rv[j].cnt; ++q) posPool.push_back(rv[j].pos[q]);
            posLen[pair] = total;
        }
    }

    // The plain lane-per-alignment sweep of the reads in c.direct over the whole target.
    void plain_sweep(LaneGroup& c) {
        const std::vector<int>& list = c.list;
        const std::vector<int>& direct = c.direct;
        if (direct.empty()) return;
        int chunks = 1, chunkLen = 0;
        lane_geometry(c, (int)direct.size(), c.nw, chunks, chunkLen, true);
        std::vector<int> kInit(direct.size());
        for (size_t s = 0; s < direct.size(); ++s) kInit[s] = c.bound[direct[s]] + 1;
        std::vector<Rec> recs;
        std::vector<Ovf> ovf;
        std::vector<int> incomplete;
        long long missing = 0;
        lane_sweep(c, direct, kInit, c.nw, chunks, chunkLen, 0, 0, 0, recs, ovf);
        lane_merge(c, direct, chunks, recs, nullptr, incomplete, missing);
        if (incomplete.empty()) return;
        // Second pass over the few reads with more than KPOS end positions in one chunk: start from the
        // known minimum so that only final positions are recorded, with a list sized from the counts of
        // the first pass, on a finer chunking of the target.
        std::vector<int> subK(incomplete.size());
        for (size_t s = 0; s < incomplete.size(); ++s) subK[s] = best[list[incomplete[s]]];
        int chunks2 = 1, chunkLen2 = 0;
        lane_geometry(c, (int)incomplete.size(), c.nw, chunks2, chunkLen2, true);
        std::vector<Rec> recs2;
        std::vector<Ovf> ovf2;
        std::vector<int> still;
        long long dummy = 0;
        lane_sweep(c, incomplete, subK, c.nw, chunks2, chunkLen2, (int)missing + 16, 0, 0, recs2, ovf2);
        lane_merge(c, incomplete, chunks2, recs2, &ovf2, still, dummy);
    }

    // Distance pass of one group of pairs that share target `t` and word class `nw` (queries <= 256
    // rows): the stages of the candidate filter (HW over a long target), each on the
    // reads the previous ones left undecided, then the plain lane-per-alignment sweep of what is left.
    void lane_group(int t, int nw, const std::vector<int>& list) {
        const Target& tg = p->tg[t];
        const int G = (int)list.size();
        LaneGroup c{t, nw, list, tg, tg.len, std::vector<int>(G), std::vector<int>(G, -1), std::vector<int>()};
        for (int s = 0; s < G; ++s) {
            const int m = p->qlen[list[s]];
            c.bound[s] = (k < 0 || k > m) ? m : k;  // distances never exceed m in HW/SHW (ref cpp:566-568)
            stats.k1Cells += (long long)m * c.n;
        }
        c.direct.reserve(G);
        std::vector<int> cur(G);
        for (int s = 0; s < G; ++s) cur[s] = s;
        const bool filtered = mode == MODE_HW && c.n >= tun.filterMinTarget;
        if (filtered) {
            trace.mark("compute: classify");
            for (int level = 0; level < tun.filterSeedLevels && tun.filterSeedK > 0 && !p->hasEq && !cur.empty(); ++level) {
                std::vector<int> next;
                seed_stage(c, level, cur, next);
                cur.swap(next);
            }
            const int stageP[2] = {32, 64};
            const int stageK[2] = {tun.filterK1, tun.filterK0};
            for (int st = 0; st < 2; ++st) {
                if (stageK[st] <= 0 || cur.empty()) continue;
                if (stageP[st] / 32 >= nw) continue;  // the prefix must be shorter than the read's word class
                std::vector<int> next;
                prefix_stage(c, stageP[st], stageK[st], cur, next);
                cur.swap(next);
            }
        }
        c.direct.insert(c.direct.end(), cur.begin(), cur.end());
        if (filtered) stats.filterFallback += (long long)c.direct.size();
        trace.mark("filter: collect");
        plain_sweep(c);
    }

    // Distance pass of everything else: one alignment per warp (or per thread with its own target).
    void warp_distance() {
        // ---- W distance pass ------------------------------------------------------------------
        {
            std::vector<int> pending = wPairs;
            int kRound = 64;  // ref cpp:201: the doubling schedule only matters for speed
            while (!pending.empty()) {
                std::vector<WTask> tasks;
                std::vector<int> later;
                for (int pair : pending) {
                    const int m = p->qlen[pair], n = p->tlen[pair];
                    int bound = -1;
                    if (mode == MODE_NW) {
                        if (k >= 0) {
                            bound = k;
                        } else if (ceil_div(m, 32) > 32) {
                            bound = kRound;
                            if (bound < abs(n - m)) {
                                later.push_back(pair);
                                continue;
                            }
                        }
                    }
                    WPlan pl = plan_w(m, n, mode, bound);
                    WTask t;
                    t.pair = pair;
                    t.qOff = p->qoff[pair];
                    t.tOff = p->tg[p->tidx[pair]].off;
                    t.m = m;
                    t.n = n;
                    t.mode = mode;
                    t.flags = pl.slide ? WF_SLIDE : 0;
                    t.dhi = pl.dhi;
                    t.R = pl.R;
                    t.nWp = pl.nWp;
                    t.kInit = ((k < 0 || k > m) ? m : k) + 1;
                    t.tag = pl.slide ? bound : -1;  // a sliding result is only valid when <= bound
                    t.wantPositions = (mode != MODE_NW);
                    tasks.push_back(std::move(t));
                }
                runner.run(tasks);
                for (WTask& t : tasks) {
                    stats.wCells += (long long)t.m * t.n;
                    if (t.tag >= 0 && t.rec.best > t.tag) {  // outside the band of this round
                        if (k < 0) later.push_back(t.pair);
                        else best[t.pair] = 0x7fffffff;
                        continue;
                    }
                    best[t.pair] = t.rec.cnt > 0 ? t.rec.best : 0x7fffffff;
                    cnt[t.pair] = t.rec.cnt;
                    posStart[t.pair] = (long long)posPool.size();
                    for (int q = 0; q < std::min(t.rec.cnt, KPOS); ++q) posPool.push_back(t.rec.pos[q]);
                    posPool.insert(posPool.end(), t.extra.begin(), t.extra.end());
                    posLen[t.pair] = (int)((long long)posPool.size() - posStart[t.pair]);
                }
                pending.swap(later);
                if (kRound < (1 << 29)) kRound *= 2;
            }
        }
    }

    // editDistance and endLocations per pair from the sweep outcomes (ref cpp:219-225 and the -1 rule).
    void collect_ends() {
        // ---- distances and end locations: counts per pair, offsets, fill (on a few host threads) -------
        // ref cpp:670, 681-693: the padded bottom cell of column W-1 shows up as end location -1
        auto accepted = [&](int i) -> int {  // number of end locations of pair i, or -1 if it has no result
            if (p->special[i]) return -1;
            if (best[i] < 0 || best[i] == 0x7fffffff) return -1;  // rejected up front or nothing tracked
            if (k >= 0 && best[i] > k) return -1;
            if (mode == MODE_NW) return 1;
            const int m = p->qlen[i];
            if (best[i] > m) return -1;
            const int W64 = ceil_div(m, 64) * 64 - m;
            return posLen[i] + ((best[i] == m && W64 > 0) ? 1 : 0);
        };
        const size_t nparts = host_parts((size_t)N, 65536);
        std::vector<long long> partCount(nparts + 1, 0);
        std::vector<int> bad(nparts, 0);
        auto run = [&](const std::function<void(size_t, size_t, size_t)>& fn) {
            HostPool::get().run(nparts, [&](size_t t) { fn(t, (size_t)N * t / nparts, (size_t)N * (t + 1) / nparts); });
        };
        run([&](size_t t, size_t lo, size_t hi) {
            long long c = 0;
            for (size_t i = lo; i < hi; ++i) {
                const int a = accepted((int)i);
                if (a > 0) c += a;
                if (a >= 0 && mode != MODE_NW && posLen[i] != cnt[i]) bad[t] = 1;
            }
            partCount[t + 1] = c;
        });
        for (size_t t = 0; t < nparts; ++t) {
            if (bad[t]) throw std::runtime_error("internal: end-location count mismatch");
            partCount[t + 1] += partCount[t];
        }
        p->endPool.resize((size_t)partCount[nparts]);
        run([&](size_t t, size_t lo, size_t hi) {
            long long at = partCount[t];
            for (size_t i = lo; i < hi; ++i) {
                p->endStart[i] = at;
                const int a = accepted((int)i);
                if (a < 0) {
                    p->ed[i] = -1;
                    p->endCount[i] = 0;
                    continue;
                }
                p->ed[i] = best[i];
                p->endCount[i] = a;
                if (mode == MODE_NW) {
                    p->endPool[(size_t)at++] = p->tlen[i] - 1;  // ref cpp:221-225
                    continue;
                }
                if (a > posLen[i]) p->endPool[(size_t)at++] = -1;
                if (posLen[i]) memcpy(p->endPool.data() + at, posPool.data() + posStart[i], sizeof(int) * (size_t)posLen[i]);
                at += posLen[i];
            }
        });
    }

    void start_locations() {
        // ---- start locations (ref cpp:228-272) ------------------------------------------------
        const bool wantLoc = p->cfg.task == EDLIB_TASK_LOC || p->cfg.task == EDLIB_TASK_PATH;
        if (wantLoc) {
            p->startPool.assign(p->endPool.size(), 0);
            if (mode == MODE_HW) {
                std::vector<WTask> tasks;
                std::vector<long long> slotOf;
                std::vector<LJob> lj[9];          // short queries: straight to the lane kernel, per word class
                std::vector<long long> lslot[9];
                std::vector<int> lpair[9];
                for (int i = 0; i < N; ++i) {
                    if (p->ed[i] < 0) continue;
                    const int m = p->qlen[i];
                    const bool lane = lane_ok(m);
                    for (int q = 0; q < p->endCount[i]; ++q) {
                        const long long slot = p->endStart[i] + q;
                        const int e = p->endPool[(size_t)slot];
                        if (e < 0) continue;  // ref cpp:237-249: start 0
                        if (lane) {
                            const int nw = ceil_div(m, 32);
                            LJob j;
                            memset(&j, 0, sizeof(j));
                            j.qOff = p->qoff[i];
                            j.tOff = p->tg[p->tidx[i]].off + (uint64_t)e;  // first symbol read, walking backward
                            j.m = m;
                            j.n = (int)std::min<long long>((long long)e + 1, (long long)m + p->ed[i]);
                            j.kInit = p->ed[i] + 1;
                            lj[nw].push_back(j);
                            lslot[nw].push_back(slot);
                            lpair[nw].push_back(i);
                            continue;
                        }
                        WTask t;
                        t.pair = i;
                        t.qOff = p->qoff[i];
                        t.tOff = p->tg[p->tidx[i]].off + (uint64_t)e;  // first symbol read, walking backward
                        t.m = m;
                        t.n = (int)std::min<long long>((long long)e + 1, (long long)m + p->ed[i]);
                        t.mode = MODE_SHW;
                        t.flags = WF_QREV | WF_TREV;
                        t.kInit = p->ed[i] + 1;
                        WPlan pl = plan_w(t.m, t.n, MODE_SHW, -1);
                        t.R = pl.R;
                        t.nWp = pl.nWp;
                        tasks.push_back(std::move(t));
                        slotOf.push_back(slot);
                    }
                }
                trace.mark("starts: jobs built");
                for (int nw = 1; nw <= 8; ++nw) {
                    if (lj[nw].empty()) continue;
                    std::vector<Rec> recs;
                    lane_launch(lj[nw], nw, MODE_SHW, true, recs);
                    for (size_t j = 0; j < recs.size(); ++j) {
                        if (recs[j].cnt <= 0 || recs[j].best != p->ed[lpair[nw][j]])
                            throw std::runtime_error("internal: start-location sweep disagrees");
                        const int e = p->endPool[(size_t)lslot[nw][j]];
                        p->startPool[(size_t)lslot[nw][j]] = e - recs[j].last;  // ref cpp:260
                    }
                }
                trace.mark("starts: lane sweeps");
                runner.run(tasks);
                for (size_t j = 0; j < tasks.size(); ++j) {
                    const WTask& t = tasks[j];
                    if (t.rec.cnt <= 0 || t.rec.best != p->ed[t.pair]) throw std::runtime_error("internal: start-location sweep disagrees");
                    const int e = p->endPool[(size_t)slotOf[j]];
                    p->startPool[(size_t)slotOf[j]] = e - t.rec.last;  // ref cpp:260
                }
            }
        }
    }

    void paths() {
        // ---- alignment path (ref cpp:276-289, 1161-1213, 1231-1396) ---------------------------
        // obtainAlignment as a level-synchronous tree: a node is an NW sub-problem (query slice,
        // target slice, known score).  Inside the reference's 1 MiB rule (cpp:1188-1190) it is a
        // leaf: matrix-storing sweep + traceback kernel.  Otherwise it is split like
        // obtainAlignmentHirschberg: the score column left of the target's middle from a forward
        // sweep and the one right of it from a reversed sweep (both on the device, cpp:1252-1260),
        // the split row chosen by the reference's candidate order (cpp:1321-1353), both halves
        // becoming nodes of the next level (cpp:1372-1380).  All nodes of a level run in one batch.
        if (p->cfg.task == EDLIB_TASK_PATH) {
            struct Node {
                int pair;
                uint64_t qOff, tOff;
                int m, n, best;
                int left = -1, right = -1;
                long long opsOff = -1;  // into opsPool (leaf) ...
                int opsLen = 0;
                int fillOp = -1;        // ... or a run of one op (empty side, cpp:1168-1175)
            };
            std::vector<Node> nodes;
            std::vector<int> rootOf(N, -1), frontier, leaves;
            for (int i = 0; i < N; ++i) {
                if (p->ed[i] < 0) continue;
                const int s0 = p->startPool[(size_t)p->endStart[i]], e0 = p->endPool[(size_t)p->endStart[i]];
                Node nd;
                nd.pair = i;
                nd.qOff = p->qoff[i];
                nd.tOff = p->tg[p->tidx[i]].off + (uint64_t)s0;
                nd.m = p->qlen[i];
                nd.n = e0 - s0 + 1;
                nd.best = p->ed[i];
                rootOf[i] = (int)nodes.size();
                frontier.push_back((int)nodes.size());
                nodes.push_back(nd);
            }
            while (!frontier.empty()) {
                std::vector<int> split;
                for (int id : frontier) {
                    Node& nd = nodes[id];
                    if (nd.m == 0 || nd.n <= 0) {
                        nd.fillOp = (nd.m == 0) ? EDLIB_EDOP_DELETE : EDLIB_EDOP_INSERT;
                        nd.opsLen = nd.m + std::max(nd.n, 0);
                        continue;
                    }
                    const long long matrixBytes = 20LL * ceil_div(nd.m, 64) * nd.n + 8LL * nd.n;  // cpp:1188-1190
                    if (matrixBytes < 1024 * 1024) leaves.push_back(id);
                    else split.push_back(id);
                }
                frontier.clear();
                if (split.empty()) break;
                std::vector<WTask> tasks;
                tasks.reserve(split.size() * 2);
                for (int id : split) {
                    const Node& nd = nodes[id];
                    const int leftW = nd.n / 2, rightW = nd.n - leftW;  // cpp:1247-1248
                    const WPlan pl = plan_w(nd.m, nd.n, MODE_NW, nd.best);  // band of the WHOLE node
                    WTask f;
                    f.pair = id;
                    f.qOff = nd.qOff;
                    f.tOff = nd.tOff;
                    f.m = nd.m;
                    f.n = leftW;
                    f.mode = MODE_NW;
                    f.flags = WF_STOPCOL | (pl.slide ? WF_SLIDE : 0);
                    f.dhi = pl.dhi;
                    f.stopCol = leftW - 1;
                    f.R = pl.R;
                    f.nWp = pl.nWp;
                    f.splitSide = 0;
                    f.splitBest = nd.best;
                    WTask r = f;
                    r.tOff = nd.tOff + (uint64_t)nd.n - 1;  // reversed: first symbol read is the last one
                    r.n = rightW;
                    r.flags |= WF_QREV | WF_TREV;
                    r.stopCol = rightW - 1;
                    r.splitSide = 1;
                    tasks.push_back(std::move(f));
                    tasks.push_back(std::move(r));
                }
                runner.run(tasks);
                for (size_t s = 0; s < split.size(); ++s) {
                    const int id = split[s];
                    const Node nd = nodes[id];
                    const int leftW = nd.n / 2, rightW = nd.n - leftW;
                    const SplitOut& so = tasks[2 * s].split;  // found on the device (split_kernel)
                    const int h = so.h;
                    if (h < 0) throw std::runtime_error("internal: Hirschberg split not found");
                    Node a, b;
                    a.pair = b.pair = nd.pair;
                    a.qOff = nd.qOff;
                    a.tOff = nd.tOff;
                    a.m = h;
                    a.n = leftW;
                    a.best = so.left;
                    b.qOff = nd.qOff + (uint64_t)h;
                    b.tOff = nd.tOff + (uint64_t)leftW;
                    b.m = nd.m - h;
                    b.n = rightW;
                    b.best = so.right;
                    nodes[id].left = (int)nodes.size();
                    frontier.push_back((int)nodes.size());
                    nodes.push_back(a);
                    nodes[id].right = (int)nodes.size();
                    frontier.push_back((int)nodes.size());
                    nodes.push_back(b);
                }
            }
            {
                std::vector<WTask> tasks;
                std::vector<LJob> lj[9];  // leaves with short queries: lane kernel, per word class
                std::vector<int> lnode[9];
                for (int id : leaves) {
                    const Node& nd = nodes[id];
                    if (lane_ok(nd.m)) {
                        const int nw = ceil_div(nd.m, 32);
                        LJob j;
                        memset(&j, 0, sizeof(j));
                        j.qOff = nd.qOff;
                        j.tOff = nd.tOff;
                        j.m = nd.m;
                        j.n = nd.n;
                        lj[nw].push_back(j);
                        lnode[nw].push_back(id);
                        continue;
                    }
                    WTask t;
                    t.pair = id;
                    t.qOff = nd.qOff;
                    t.tOff = nd.tOff;
                    t.m = nd.m;
                    t.n = nd.n;
                    t.mode = MODE_NW;
                    t.flags = WF_STORE;
                    const WPlan pl = plan_w(nd.m, nd.n, MODE_NW, -1);
                    t.R = pl.R;
                    t.nWp = pl.nWp;
                    tasks.push_back(std::move(t));
                }
                trace.mark("paths: tree + leaf jobs built");
                for (int nw = 1; nw <= 8; ++nw) {
                    if (lj[nw].empty()) continue;
                    lane_paths(lj[nw], nw, [&](size_t j, const uint8_t* ops, int len, int score) {
                        Node& nd = nodes[lnode[nw][j]];
                        if (score != nd.best) throw std::runtime_error("internal: path sweep disagrees with the distance");
                        nd.opsOff = (long long)opsPool.size();
                        nd.opsLen = len;
                        opsPool.insert(opsPool.end(), ops, ops + len);
                    });
                }
                runner.run(tasks);
                for (const WTask& t : tasks) {
                    Node& nd = nodes[t.pair];
                    if (t.rec.best != nd.best) throw std::runtime_error("internal: path sweep disagrees with the distance");
                    nd.opsOff = t.opsOff;
                    nd.opsLen = t.opsLen;
                }
            }
            trace.mark("paths: leaf sweeps + tracebacks");
            // in-order concatenation (cpp:1388-1391)
            std::vector<int> stack;
            for (int i = 0; i < N; ++i) {
                if (rootOf[i] < 0) continue;
                p->alnStart[i] = (long long)p->alnPool.size();
                stack.assign(1, rootOf[i]);
                while (!stack.empty()) {
                    const int id = stack.back();
                    stack.pop_back();
                    const Node& nd = nodes[id];
                    if (nd.left >= 0) {
                        stack.push_back(nd.right);
                        stack.push_back(nd.left);
                    } else if (nd.fillOp >= 0) {
                        p->alnPool.insert(p->alnPool.end(), (size_t)nd.opsLen, (uint8_t)nd.fillOp);
                    } else {
                        p->alnPool.insert(p->alnPool.end(), opsPool.begin() + nd.opsOff, opsPool.begin() + nd.opsOff + nd.opsLen);
                    }
                }
                p->alnLen[i] = (int)(p->alnPool.size() - (size_t)p->alnStart[i]);
            }
        }
    }
};

}  // namespace

// Groups the pairs by (target, word class): a pure function of the lengths, the distinct targets and the
// config, so prepare() runs it on the host workers while the sequences travel to the device; the lists stay
// with the batch (and, through the spare batch object, keep their storage from call to call).
void Engine::classify(Prepared* p) {
    const int N = p->N;
    const int mode = p->mode;
    const int k = p->cfg.k;
    p->special.resize(N);
    std::map<std::pair<int, int>, std::vector<int>>& groups = p->groups;  // (target, nw32) -> pairs, ascending
    std::vector<int>& wPairs = p->wPairsBase;
    wPairs.clear();
    for (auto& kv : groups) kv.second.clear();
    {
        // contiguous ranges of pairs are classified on a few host threads and concatenated in order
        const size_t nparts = host_parts((size_t)N, 65536);
        std::vector<Prepared::Part>& parts = p->parts;
        parts.resize(nparts);
        for (auto& P : parts) {
            for (auto& kv : P.groups) kv.second.clear();
            P.wPairs.clear();
        }
        auto classify = [&](size_t t) {
            Prepared::Part& P = parts[t];
            std::pair<int, int> lastKey(-1, -1);
            std::vector<int>* lastList = nullptr;
            const int lo = (int)((size_t)N * t / nparts), hi = (int)((size_t)N * (t + 1) / nparts);
            for (int i = lo; i < hi; ++i) {
                const int m = p->qlen[i], n = p->tlen[i];
                p->special[i] = (m == 0 || n == 0) ? 1 : 0;
                if (p->special[i]) continue;
                if (mode == MODE_NW && k >= 0 && k < abs(n - m)) continue;  // ref cpp:744
                if (m <= 256) {
                    const std::pair<int, int> key(p->tidx[i], ceil_div(m, 32));
                    if (key != lastKey) {  // neighbours usually share their group
                        lastKey = key;
                        lastList = &P.groups[key];
                        if (lastList->empty()) lastList->reserve((size_t)(hi - i));
                    }
                    lastList->push_back(i);
                } else {
                    P.wPairs.push_back(i);
                }
            }
        };
        HostPool::get().run(nparts, classify);
        for (Prepared::Part& P : parts) {
            for (auto& kv : P.groups) {
                if (kv.second.empty()) continue;
                std::vector<int>& dst = groups[kv.first];
                dst.insert(dst.end(), kv.second.begin(), kv.second.end());
            }
            wPairs.insert(wPairs.end(), P.wPairs.begin(), P.wPairs.end());
        }
        // drop the keys this batch does not use (bounded memory across differently shaped batches)
        for (auto it = groups.begin(); it != groups.end();) it = it->second.empty() ? groups.erase(it) : std::next(it);
    }
    p->classified = true;
}

void Engine::compute(Prepared* p) {
    Backend* be = be_;
    be->reset_timing();
    const int N = p->N;
    const int mode = p->mode;
    // ed / endStart / endCount are written for every pair by collect_ends, special by the classification
    p->ed.resize(N);
    p->endStart.resize(N);
    p->endCount.resize(N);
    p->endPool.clear();
    p->startPool.clear();
    if (p->cfg.task == EDLIB_TASK_PATH) {
        p->alnStart.assign(N, -1);
        p->alnLen.assign(N, 0);
    } else {
        p->alnStart.clear();
        p->alnLen.clear();
    }
    p->alnPool.clear();
    stats.k1Cells = stats.wCells = 0;
    stats.filterDecided = stats.filterFallback = stats.filterWindows = 0;

    Pass ps(*this, be, p);
    std::vector<int>& wPairs = ps.wPairs;
    Trace& trace = ps.trace;

    // ---- classification (normally done by prepare() while the upload is in flight) --------
    if (!p->classified) classify(p);
    std::map<std::pair<int, int>, std::vector<int>>& groups = p->groups;
    wPairs = p->wPairsBase;

    // ---- K1 groups ----------------------------------------------------------------------
    for (auto& kv : groups) {
        const std::vector<int>& list = kv.second;
        // Small groups go to the warp kernel, except HW over a long target: there the lane kernel
        // can cut the target into chunks and spread even one alignment over many CTAs.
        if ((int)list.size() < tun.k1MinGroup && !(mode == MODE_HW && p->tg[kv.first.first].len >= 8 * tun.k1MinChunk)) {
            wPairs.insert(wPairs.end(), list.begin(), list.end());
            continue;
        }
        const int t = kv.first.first, nw = kv.first.second;
        {
            int bt = 0, rc = 0;
            be->k1_shape(nw, p->ncodes, (int)list.size(), &bt, &rc);
            if (rc <= 0) {  // alphabet too large for per-thread Peq rows in shared memory
                wPairs.insert(wPairs.end(), list.begin(), list.end());
                continue;
            }
        }
        ps.lane_group(t, nw, list);
    }
    trace.mark("compute: K1 groups done");
    ps.warp_distance();
    trace.mark("compute: W distance pass");
    ps.collect_ends();
    trace.mark("compute: end locations");
    ps.start_locations();
    ps.paths();
    trace.mark("compute: starts + paths");
    be->sync();
    stats.kernelMs = be->kernel_ms(nullptr);
    stats.k1Ms = be->kernel_ms("k1") + be->kernel_ms("k1_prefix");
    stats.kernelReport = be->kernel_report();
    stats.launches = be->launches();
    p->computed = true;
}

// ---------------------------------------------------------------------------------------------
// materialize / release / one-shot
// ---------------------------------------------------------------------------------------------
void Engine::materialize(Prepared* p, EdlibAlignResult* results) {
    Trace trace;
    const int N = p->N;
    parallel_ranges((size_t)N, 65536, [=](size_t lo, size_t hi) {
    for (int i = (int)lo; i < (int)hi; ++i) {
        EdlibAlignResult& r = results[i];
        memset(&r, 0, sizeof(r));
        r.status = EDLIB_STATUS_OK;
        r.editDistance = -1;
        r.alphabetLength = p->alphaLen[i];
        const int m = p->qlen[i], n = p->tlen[i];
        if (p->special[i]) {  // ref cpp:166-184
            const int rawMode = (int)p->cfg.mode;
            if (rawMode == EDLIB_MODE_NW || rawMode == EDLIB_MODE_SHW || rawMode == EDLIB_MODE_HW) {
                r.editDistance = rawMode == EDLIB_MODE_NW ? std::max(m, n) : m;
                r.endLocations = static_cast<int*>(malloc(sizeof(int)));
                r.endLocations[0] = rawMode == EDLIB_MODE_NW ? n - 1 : -1;
                r.numLocations = 1;
            } else {
                r.status = EDLIB_STATUS_ERROR;
            }
            continue;
        }
        if (p->ed[i] < 0) continue;
        r.editDistance = p->ed[i];
        const int c = p->endCount[i];
        r.numLocations = c;
        r.endLocations = static_cast<int*>(malloc(sizeof(int) * (size_t)std::max(c, 1)));
        memcpy(r.endLocations, p->endPool.data() + p->endStart[i], sizeof(int) * (size_t)c);
        if (!p->startPool.empty() || p->cfg.task == EDLIB_TASK_LOC || p->cfg.task == EDLIB_TASK_PATH) {
            r.startLocations = static_cast<int*>(malloc(sizeof(int) * (size_t)std::max(c, 1)));
            memcpy(r.startLocations, p->startPool.data() + p->endStart[i], sizeof(int) * (size_t)c);
        }
        if (!p->alnStart.empty() && p->alnStart[i] >= 0) {
            r.alignmentLength = p->alnLen[i];
            r.alignment = static_cast<unsigned char*>(malloc((size_t)std::max(p->alnLen[i], 1)));
            memcpy(r.alignment, p->alnPool.data() + p->alnStart[i], (size_t)p->alnLen[i]);
        }
    }
    });
    trace.mark("materialize");
}

void Engine::release(Prepared* p) {
    if (!p) return;
    if (spare_) {
        delete p;
        return;
    }
    // keep the object for the next batch: device buffers go back to the pool now
    p->dSeq.reset();
    p->dQoff.reset();
    p->dQlen.reset();
    p->dEqtab.reset();
    spare_ = p;
}

Engine::~Engine() { delete spare_; }

int Engine::align_batch(const BatchInput& in, EdlibAlignResult* results) {
    Prepared* p = nullptr;
    stats = EngineStats();
    try {
        p = prepare(in);
        compute(p);
        materialize(p, results);
        release(p);
        return EDLIB_STATUS_OK;
    } catch (const std::exception& e) {
        lastError = e.what();
        if (p) release(p);
        for (int i = 0; i < in.numPairs; ++i) {
            memset(&results[i], 0, sizeof(results[i]));
            results[i].status = EDLIB_STATUS_ERROR;
            results[i].editDistance = -1;
        }
        return EDLIB_STATUS_ERROR;
    }
}

}  // namespace eb
