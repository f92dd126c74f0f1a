#!/usr/bin/env python
"""Long randomized parity run (not part of pytest): every case generator of tests/cases.py with many seeds
through the product library, bit-exact against the reference build.  Usage: python scripts/stress.py [minutes]"""
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import cases  # noqa: E402
import parity  # noqa: E402
from helpers import product  # noqa: E402

budget = float(sys.argv[1]) * 60 if len(sys.argv) > 1 else 120.0
if os.environ.get("EDLIB_STRESS_LIB"):  # e.g. the CPU emulation or one of its sanitizer builds (scripts/sanitize.sh)
    from edlib_b200._ffi import EdlibLib
    lib = EdlibLib(os.environ["EDLIB_STRESS_LIB"], has_batch=True)
else:
    lib = product()
t0 = time.time()
total = 0
seed = 1000
plan = [("single", cases.single_pair_cases, 400, parity.run_single), ("batch", cases.batch_cases, 12, parity.run_batches),
        ("pairwise", cases.pairwise_cases, 12, parity.run_batches), ("filter", cases.filter_cases, 10, parity.run_batches),
        ("long", cases.long_cases, 12, parity.run_single), ("path", cases.path_cases, 20, parity.run_single),
        ("stream", cases.stream_cases, 6, parity.run_batches), ("tied_ends", cases.tied_ends_cases, 6, parity.run_batches),
        ("band", cases.band_cases, 6, parity.run_batches), ("long_hw", cases.long_hw_cases, 3, parity.run_batches),
        ("equalities", cases.equality_read_cases, 6, parity.run_batches), ("small_k", cases.small_k_cases, 8, parity.run_batches),
        ("boundary_mix", cases.boundary_mix_cases, 40, parity.run_batches)]
counts = {}
while time.time() - t0 < budget:
    for name, gen, n, runner in plan:
        seed += 1
        got = runner(lib, seed, n, gen=gen)
        counts[name] = counts.get(name, 0) + got
        total += got
print("stress ok:", total, "alignments bit-exact in %.0f s" % (time.time() - t0), counts)
