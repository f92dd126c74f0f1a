#!/usr/bin/env python
"""Measures the non-headline BASELINE.json configs through the C ABI (host buffers, end to end) with a
bit-exact spot check against the reference build, and the single-call latency.  Not the bench line
(bench.py is); the numbers go into DESIGN.md.   Usage: python scripts/config_runs.py [--pairs3 N] [--reads4 N]"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
from edlib_b200 import workloads  # noqa: E402
from edlib_b200._ffi import AlignResult, EdlibLib, make_config, product_path, result_to_dict  # noqa: E402
import bench  # noqa: E402


def batch_call(lib, qarrs, tarrs, k, mode, task):
    n = len(qarrs)
    qptr = np.array([a.ctypes.data for a in qarrs], dtype=np.uint64)
    tptr = np.array([a.ctypes.data for a in tarrs], dtype=np.uint64)
    qlen = np.array([len(a) for a in qarrs], dtype=np.int32)
    tlen = np.array([len(a) for a in tarrs], dtype=np.int32)
    cfg, _ = make_config(k, mode, task)
    res = (AlignResult * n)()
    t0 = time.perf_counter()
    rc = lib.lib.edlibAlignBatch(bench.as_pp(qptr), bench.as_pi(qlen), bench.as_pp(tptr), bench.as_pi(tlen), n, cfg, res)
    dt = time.perf_counter() - t0
    assert rc == 0, lib.lib.edlibB200LastError()
    st = bench.Stats()
    lib.lib.edlibB200LastStats(C.byref(st))
    return res, dt, st


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs3", type=int, default=20000)
    ap.add_argument("--reads4", type=int, default=100000)
    args = ap.parse_args()
    lib = EdlibLib(product_path(), has_batch=True)
    L = lib.lib
    L.edlibB200LastError.restype = C.c_char_p
    L.edlibB200LastStats.argtypes = [C.POINTER(bench.Stats)]
    L.edlibB200FreeResults.argtypes = [C.c_void_p, C.c_int]
    assert L.edlibB200Available() == 1
    ref_so = os.path.join(REPO, "oracle", "_ref", "libedlib_ref.so")
    ref = EdlibLib(ref_so) if os.path.exists(ref_so) else EdlibLib(os.path.join(REPO, "oracle", "liboracle.so"), prefix="oracle")
    out = {}

    # ---- config 3: 10 kbp pairs, NW, k = 500, LOC -------------------------------------------------
    qs, ts = workloads.long_pairs(args.pairs3, 10_000, seed=43)
    batch_call(lib, qs[:64], ts[:64], 500, 0, 1)  # warm-up (context)
    res, dt_first, st = batch_call(lib, qs, ts, 500, 0, 1)  # first full-size call: pools and staging grow
    L.edlibB200FreeResults(res, len(qs))
    res, dt, st = batch_call(lib, qs, ts, 500, 0, 1)        # steady state
    cells = float(sum(len(q) * len(t) for q, t in zip(qs, ts)))
    bad = 0
    for i in range(0, len(qs), max(1, len(qs) // 200)):
        exp = ref.align(qs[i].tobytes(), ts[i].tobytes(), 500, 0, 1)
        bad += result_to_dict(res[i]) != exp
    eds = [res[i].editDistance for i in range(len(qs))]
    out["config3"] = {"pairs": len(qs), "e2e_first_call_s": dt_first, "e2e_s": dt, "pairs_per_s": len(qs) / dt, "gcups_e2e": cells / dt / 1e9,
                      "kernel_ms": st.kernelMs, "gcups_kernel": cells / (st.kernelMs / 1e3) / 1e9, "launches": st.launches,
                      "h2d": st.h2dBytes, "d2h": st.d2hBytes, "mean_ed": float(np.mean(eds)), "spot_mismatches": bad}
    L.edlibB200FreeResults(res, len(qs))
    print(json.dumps({"config3": out["config3"]}), flush=True)

    # ---- config 3 PATH (Hirschberg regime), smaller batch -----------------------------------------
    nq = max(64, args.pairs3 // 20)
    res, dt, st = batch_call(lib, qs[:nq], ts[:nq], 500, 0, 2)
    bad = 0
    for i in range(0, nq, max(1, nq // 50)):
        bad += result_to_dict(res[i]) != ref.align(qs[i].tobytes(), ts[i].tobytes(), 500, 0, 2)
    out["config3_path"] = {"pairs": nq, "e2e_s": dt, "pairs_per_s": nq / dt, "kernel_ms": st.kernelMs, "launches": st.launches,
                           "spot_mismatches": bad}
    L.edlibB200FreeResults(res, nq)
    print(json.dumps({"config3_path": out["config3_path"]}), flush=True)

    # ---- config 4: reads vs 5 Mbp target, HW, PATH (+ CIGAR on a sample) ---------------------------
    target, reads = workloads.reads_vs_target(args.reads4, 150, 5_000_000, seed=42)
    rl = [reads[i] for i in range(args.reads4)]
    res, dt_first, st = batch_call(lib, rl, [target] * args.reads4, -1, 2, 2)  # first call at this size
    L.edlibB200FreeResults(res, args.reads4)
    res, dt, st = batch_call(lib, rl, [target] * args.reads4, -1, 2, 2)        # steady state
    cells = float(args.reads4) * 150 * 5_000_000
    bad = 0
    tb = target.tobytes()
    for i in range(0, args.reads4, max(1, args.reads4 // 60)):
        bad += result_to_dict(res[i]) != ref.align(rl[i].tobytes(), tb, -1, 2, 2)
    alen = np.mean([res[i].alignmentLength for i in range(args.reads4)])
    t0 = time.perf_counter()
    ncig = min(args.reads4, 20000)
    for i in range(ncig):
        p = lib._cigar(res[i].alignment, res[i].alignmentLength, 1)
        lib._libc.free(p)
    cig_dt = time.perf_counter() - t0
    out["config4"] = {"reads": args.reads4, "e2e_first_call_s": dt_first, "e2e_s": dt, "aln_per_s": args.reads4 / dt, "gcups_e2e": cells / dt / 1e9,
                      "kernel_ms": st.kernelMs, "k1_ms": st.k1Ms, "launches": st.launches, "mean_alignment_len": float(alen),
                      "cigar_us_each": 1e6 * cig_dt / ncig, "spot_mismatches": bad}
    L.edlibB200FreeResults(res, args.reads4)
    print(json.dumps({"config4": out["config4"]}), flush=True)

    # ---- single-call latency (the reference's own calling pattern) ---------------------------------
    q = reads[0].tobytes()
    t10k = target[:10_000].tobytes()
    lib.align(q, t10k, -1, 2, 0)
    t0 = time.perf_counter()
    for _ in range(200):
        lib.align(q, t10k, -1, 2, 0)
    lat = (time.perf_counter() - t0) / 200
    t0 = time.perf_counter()
    for _ in range(20):
        ref.align(q, t10k, -1, 2, 0)
    lat_ref = (time.perf_counter() - t0) / 20
    out["single_call"] = {"shape": "150 x 10000 HW distance", "gpu_ms": lat * 1e3, "reference_cpu_ms": lat_ref * 1e3}
    print(json.dumps({"single_call": out["single_call"]}), flush=True)


if __name__ == "__main__":
    main()
