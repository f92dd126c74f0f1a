"""Runs config-3-shaped batches (10 kbp pairs, NW, k=500, LOC) for profiling the warp kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edlib_b200 import workloads
import config_runs as cr
from edlib_b200._ffi import EdlibLib, product_path
import ctypes as C
import bench
lib = EdlibLib(product_path(), has_batch=True)
lib.lib.edlibB200LastError.restype = C.c_char_p
lib.lib.edlibB200LastStats.argtypes = [C.POINTER(bench.Stats)]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
qs, ts = workloads.long_pairs(n, 10_000, seed=43)
for _ in range(2):
    res, dt, st = cr.batch_call(lib, qs, ts, 500, 0, 1)
    print("pairs", n, "e2e_s", dt, "kernel_ms", st.kernelMs)
