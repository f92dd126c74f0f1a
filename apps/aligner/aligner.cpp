// edlib-aligner (B200 build) -- the reference's command-line aligner re-expressed over the batched
// entry point.  Same options, same output lines as reference apps/aligner/aligner.cpp (flags: 49-61,
// per-query loop: 162-225, score table: 227-262, NICE printer: 331-377); the difference is HOW the
// work is issued: all queries go to the device in ONE edlibAlignBatch() call against the shared target
// instead of one edlibAlign() per query.
//
// -n N ("N best", reference 183-195 lowers k while it walks the queries): results are monotone in k
// (a bound only turns a result into -1), so the batch runs once and the reference's sequential
// k-lowering is replayed on the host over the true scores -- same lines, same order.  With -n the batch
// runs unbounded: the reference's k becomes top-1 = -1 ("no bound") once the N best are all exact
// matches, after which it reports scores above the user's -k as well; the replay reproduces that.
#include <getopt.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <queue>
#include <string>
#include <vector>

#include "edlib.h"

static bool read_fasta(const char* path, std::vector<std::string>* seqs) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    std::vector<char> buf(1 << 20);
    bool header = false, open = false;
    size_t got;
    while ((got = fread(buf.data(), 1, buf.size(), f)) > 0) {
        for (size_t i = 0; i < got; ++i) {
            const char c = buf[i];
            if (header) {
                if (c == '\n') header = false;
            } else if (c == '>') {
                header = true;
                open = false;
            } else if (c != '\n' && c != '\r') {
                if (!open) {
                    seqs->emplace_back();
                    open = true;
                }
                seqs->back().push_back(c);
            }
        }
    }
    fclose(f);
    return true;
}

// Three-line blocks of 50 columns: target, match bars, query (reference 331-377).
static void print_nice(const std::string& query, const std::string& target, const unsigned char* aln, int len, int endPos,
                       EdlibAlignMode mode) {
    int t = -1, q = -1;
    if (mode == EDLIB_MODE_HW) {
        t = endPos;
        for (int i = 0; i < len; ++i)
            if (aln[i] != EDLIB_EDOP_INSERT) --t;
    }
    for (int from = 0; from < len; from += 50) {
        const int to = std::min(len, from + 50);
        int first = -1;
        printf("T: ");
        for (int j = from; j < to; ++j) {
            if (aln[j] == EDLIB_EDOP_INSERT) putchar('-');
            else putchar(target[++t]);
            if (j == from) first = t;
        }
        printf(" (%d - %d)\n   ", std::max(first, 0), t);
        for (int j = from; j < to; ++j) putchar(aln[j] == EDLIB_EDOP_MATCH ? '|' : ' ');
        printf("\nQ: ");
        first = q;
        for (int j = from; j < to; ++j) {
            if (aln[j] == EDLIB_EDOP_DELETE) putchar('-');
            else putchar(query[++q]);
            if (j == from) first = q;
        }
        printf(" (%d - %d)\n\n", std::max(first, 0), q);
    }
}

int main(int argc, char* const argv[]) {
    bool silent = false, wantPath = false, wantLoc = false, bad = false;
    std::string mode = "NW", format = "NICE";
    int numBest = 0, kArg = -1, repeats = 1, opt;
    while ((opt = getopt(argc, argv, "m:n:k:f:r:spl")) >= 0) {
        switch (opt) {
            case 'm': mode = optarg; break;
            case 'n': numBest = atoi(optarg); break;
            case 'k': kArg = atoi(optarg); break;
            case 'f': format = optarg; break;
            case 'r': repeats = atoi(optarg); break;
            case 's': silent = true; break;
            case 'p': wantPath = true; break;
            case 'l': wantLoc = true; break;
            default: bad = true;
        }
    }
    if (bad || optind + 2 != argc) {
        fprintf(stderr,
                "\nUsage: %s [options...] <queries.fasta> <target.fasta>\nOptions:\n"
                "\t-s  No score or alignment output (silent mode).\n"
                "\t-m HW|NW|SHW  Alignment mode. [default: NW]\n"
                "\t-n N  Report only the N best (smallest score) sequences; 0 = all. [default: 0]\n"
                "\t-k K  Sequences with score > K are discarded; -1 = none. [default: -1]\n"
                "\t-p  Find and print the alignment path.\n"
                "\t-l  Find and print start locations.\n"
                "\t-f NICE|CIG_STD|CIG_EXT  Alignment path format (with -p). [default: NICE]\n"
                "\t-r N  Repeat the calculation N times (timing). [default: 1]\n",
                argv[0]);
        return 1;
    }
    if (format != "NICE" && format != "CIG_STD" && format != "CIG_EXT") {
        printf("Invalid alignment path format (-f)!\n");
        return 1;
    }
    EdlibAlignMode modeCode;
    if (mode == "SHW") modeCode = EDLIB_MODE_SHW;
    else if (mode == "HW") modeCode = EDLIB_MODE_HW;
    else if (mode == "NW") modeCode = EDLIB_MODE_NW;
    else {
        printf("Invalid mode (-m)!\n");
        return 1;
    }
    printf("Using %s alignment mode.\n", mode.c_str());
    const EdlibAlignTask task = wantPath ? EDLIB_TASK_PATH : wantLoc ? EDLIB_TASK_LOC : EDLIB_TASK_DISTANCE;

    std::vector<std::string> queries, targets;
    printf("Reading queries...\n");
    if (!read_fasta(argv[optind], &queries)) {
        printf("Error: There is no file with name %s\n", argv[optind]);
        return 1;
    }
    long long residues = 0;
    for (const auto& q : queries) residues += (long long)q.size();
    const int numQueries = (int)queries.size();
    printf("Read %d queries, %d residues total.\n", numQueries, (int)residues);
    printf("Reading target fasta file...\n");
    if (!read_fasta(argv[optind + 1], &targets) || targets.empty()) {
        printf("Error: There is no file with name %s\n", argv[optind + 1]);
        return 1;
    }
    const std::string& target = targets[0];
    printf("Read target, %d residues.\n", (int)target.size());

    printf("\nComparing queries to target...\n");
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<const char*> qp(numQueries), tp(numQueries, target.data());
    std::vector<int> ql(numQueries), tl(numQueries, (int)target.size());
    for (int i = 0; i < numQueries; ++i) {
        qp[i] = queries[i].data();
        ql[i] = (int)queries[i].size();
    }
    std::vector<EdlibAlignResult> res(numQueries);
    const EdlibAlignConfig cfg = edlibNewAlignConfig(numBest > 0 ? -1 : kArg, modeCode, task, NULL, 0);
    for (int rep = 0; rep < repeats; ++rep) {
        if (rep) for (auto& r : res) edlibFreeAlignResult(r);
        if (numQueries && edlibAlignBatch(qp.data(), ql.data(), tp.data(), tl.data(), numQueries, cfg, res.data()) != EDLIB_STATUS_OK) {
            fprintf(stderr, "Error: the device path failed (no usable GPU?)\n");
            return 2;
        }
    }

    // Replay of the reference's walk over the queries (its k tightens when -n is given).
    std::vector<int> scores(numQueries, -1);
    std::priority_queue<int> bestScores;
    int k = kArg;
    if (!wantPath || silent) printf("0/%d", numQueries);
    for (int i = 0; i < numQueries; ++i) {
        const EdlibAlignResult& r = res[i];
        const bool found = r.editDistance >= 0 && (k < 0 || r.editDistance <= k);
        scores[i] = found ? r.editDistance : -1;
        if (numBest > 0 && found) {
            bestScores.push(scores[i]);
            if ((int)bestScores.size() > numBest) bestScores.pop();
            if ((int)bestScores.size() == numBest) {
                k = bestScores.top() - 1;
                if (kArg >= 0 && kArg < k) k = kArg;
            }
        }
        if (!wantPath || silent) {
            if (i + 1 == numQueries) printf("\r%d/%d", i + 1, numQueries);
        } else if (found && r.alignment) {
            printf("\nQuery #%d (%d residues): score = %d\n", i, ql[i], scores[i]);
            if (format == "NICE") {
                print_nice(queries[i], target, r.alignment, r.alignmentLength, r.endLocations[0], modeCode);
            } else {
                printf("Cigar:\n");
                char* cigar = edlibAlignmentToCigar(r.alignment, r.alignmentLength,
                                                    format == "CIG_STD" ? EDLIB_CIGAR_STANDARD : EDLIB_CIGAR_EXTENDED);
                if (cigar) {
                    printf("%s\n", cigar);
                    free(cigar);
                } else {
                    printf("Error while printing cigar!\n");
                }
            }
        }
    }
    if (!silent && !wantPath) {
        int limit = -1;
        printf("\n");
        if (!bestScores.empty()) {
            printf("%d best scores:\n", (int)bestScores.size());
            limit = bestScores.top();
        } else {
            printf("Scores:\n");
        }
        printf("<query number>: <score>, <num_locations>, [(<start_location_in_target>, <end_location_in_target>)]\n");
        for (int i = 0; i < numQueries; ++i) {
            if (scores[i] < 0 || (limit != -1 && scores[i] > limit)) continue;
            const EdlibAlignResult& r = res[i];
            printf("#%d: %d  %d", i, scores[i], r.numLocations);
            if (r.numLocations > 0) {
                printf("  [");
                for (int j = 0; j < r.numLocations; ++j) {
                    if (r.startLocations) printf(" (%d, %d)", r.startLocations[j], r.endLocations[j]);
                    else printf(" (?, %d)", r.endLocations[j]);
                }
                printf(" ]");
            }
            printf("\n");
        }
    }
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("\nCpu time of searching: %lf\n", secs);
    for (auto& r : res) edlibFreeAlignResult(r);
    return 0;
}
