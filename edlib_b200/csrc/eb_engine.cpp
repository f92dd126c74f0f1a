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
#include "eb_engine_internal.h"

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
