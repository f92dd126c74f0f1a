// Leak check driver (LeakSanitizer at exit): repeated batches in every mode/task, staged API, handles, error paths.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>
#include "edlib.h"
#include "edlib_b200.h"
static std::mt19937 rng(7);
static std::string rs(int n, const char* a = "ACGT") { std::string s(n, 'A'); int L = (int)strlen(a); for (auto& c : s) c = a[rng() % L]; return s; }
static std::string mut(const std::string& s, double rate) {
    std::string o;
    for (char c : s) {
        double r = (rng() % 100000) / 100000.0;
        if (r < rate / 3) o += "ACGT"[rng() % 4];
        else if (r < 2 * rate / 3) { o += c; o += "ACGT"[rng() % 4]; }
        else if (r < rate) {}
        else o += c;
    }
    return o;
}
int main() {
    long long total = 0;
    for (int round = 0; round < 6; ++round) {
        for (int mode = 0; mode < 3; ++mode)
            for (int task = 0; task < 3; ++task) {
                std::string T = rs(round % 2 ? 30000 : 3000);
                int n = round % 3 == 0 ? 400 : 40;
                std::vector<std::string> q(n);
                for (auto& s : q) { int m = 20 + rng() % 400; int p = rng() % (T.size() - m); s = mode == 0 ? mut(T.substr(0, 2000), 0.02) : mut(T.substr(p, m), 0.05); }
                std::string T0 = mode == 0 ? T.substr(0, 2000) : T;
                std::vector<const char*> qp(n), tp(n); std::vector<int> ql(n), tl(n);
                for (int i = 0; i < n; ++i) { qp[i] = q[i].data(); ql[i] = (int)q[i].size(); tp[i] = T0.data(); tl[i] = (int)T0.size(); }
                EdlibEqualityPair eq[2] = {{'A', 'C'}, {'G', 'T'}};
                EdlibAlignConfig cfg = edlibNewAlignConfig(round % 2 ? -1 : 40, (EdlibAlignMode)mode, (EdlibAlignTask)task, round == 3 ? eq : NULL, round == 3 ? 2 : 0);
                std::vector<EdlibAlignResult> res(n);
                EdlibB200Target* h = (round == 4 && mode == 2) ? edlibB200TargetPrepare(T0.data(), (int)T0.size()) : NULL;
                if (edlibAlignBatch(qp.data(), ql.data(), tp.data(), tl.data(), n, cfg, res.data()) != 0) { printf("batch failed: %s\n", edlibB200LastError()); return 1; }
                if (task == 2) {
                    std::vector<char*> cg(n);
                    edlibB200AlignmentsToCigar(res.data(), n, EDLIB_CIGAR_EXTENDED, cg.data());
                    edlibB200FreeCigars(cg.data(), n);
                }
                if (round % 2) edlibB200FreeResults(res.data(), n); else for (auto& r : res) edlibFreeAlignResult(r);
                // staged
                EdlibB200Batch* b = edlibB200BatchPrepare(qp.data(), ql.data(), tp.data(), tl.data(), n, cfg);
                if (!b) { printf("prepare failed\n"); return 1; }
                EdlibB200Stats st;
                edlibB200BatchCompute(b, &st); edlibB200BatchCompute(b, &st);
                if (round % 2 == 0) { edlibB200BatchResults(b, res.data()); edlibB200FreeResults(res.data(), n); }
                edlibB200BatchFree(b);
                if (h) edlibB200TargetFree(h);
                // single calls + error paths
                EdlibAlignResult r = edlibAlign(q[0].data(), ql[0], T0.data(), tl[0], cfg); edlibFreeAlignResult(r);
                ql[1] = -5;
                int rc = edlibAlignBatch(qp.data(), ql.data(), tp.data(), tl.data(), n, cfg, res.data());
                if (rc == 0) { printf("negative length accepted\n"); return 1; }
                EdlibB200Batch* bad = edlibB200BatchPrepare(qp.data(), ql.data(), tp.data(), tl.data(), n, cfg);
                if (bad) { printf("bad prepare accepted\n"); return 1; }
                total += n;
            }
    }
    printf("leak driver done: %lld pairs\n", total);
    return 0;
}
