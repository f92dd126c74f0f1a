#!/usr/bin/env python
"""Small workloads that launch the kernels outside the headline step, for ncu captures and host-phase traces
(scripts/gpu_round.sh: stages ncu2 / longtrace):  band (config-3 shape, 20k pairs), path (200k reads, HW PATH:
lane / traceback / res kernels), long (the long reads of the E. coli fixture, HW LOC: warp kernel, seed planning).
usage: python scripts/ncu_targets.py band|path|long|long1"""
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402
from edlib_b200 import workloads  # noqa: E402
from edlib_b200._ffi import make_config  # noqa: E402


def main():
    what = sys.argv[1]
    E = bench.Engine(0)
    genome = workloads.ecoli_genome()
    t0 = time.time()
    if what == "band":
        qbuf, tbuf, tlens = workloads.long_pairs_packed(genome, 20000, 10_000, seed=43)
        n = 20000
        qptr = (qbuf.ctypes.data + np.arange(n, dtype=np.uint64) * np.uint64(qbuf.shape[1])).astype(np.uint64)
        tptr = (tbuf.ctypes.data + np.arange(n, dtype=np.uint64) * np.uint64(tbuf.shape[1])).astype(np.uint64)
        qlen = np.full(n, qbuf.shape[1], dtype=np.int32)
        cfg, _ = make_config(500, 0, 1)
        for _ in range(2):
            res = np.zeros(n, dtype=bench.RESULT_DTYPE)
            E.align_batch((qptr, qlen, tptr, tlens), n, cfg, res)
            E.free(res)
    elif what == "path":
        reads = workloads.reads_of(genome, 200_000, 150, seed=42)
        cfg, _ = make_config(-1, 2, 2)
        for _ in range(2):
            res = np.zeros(len(reads), dtype=bench.RESULT_DTYPE)
            E.align_batch(bench.pointer_arrays(reads, genome), len(reads), cfg, res)
            E.free(res)
    else:
        with open(os.path.join(REPO, "tests", "golden", "ecoli_reads.json")) as f:
            fx = json.load(f)["reads"]
        names = sorted(n for n in fx if len(fx[n]["seq"]) > 256)
        if what == "long1":
            names = [n for n in names if n.endswith("illumina_1x10000.fasta") or n.endswith("prefix10000.fasta")]
        gb = genome.tobytes()
        for n in names:
            t1 = time.time()
            r = E.lib.align(fx[n]["seq"].encode("ascii"), gb, -1, 2, 1)
            print("%-60s ed %5d  %.2f ms" % (n, r["editDistance"], 1000 * (time.time() - t1)), file=sys.stderr)
            assert r["editDistance"] == fx[n]["editDistance"]
    print(what, "done in %.1f s" % (time.time() - t0), E.kernel_report())


if __name__ == "__main__":
    main()
