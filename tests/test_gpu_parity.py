"""GPU parity tests (run with -m gpu on the B200 box): the product library, through its C ABI,
against the reference build / oracle on identical inputs -- bit-exact on every result field."""
import os

import pytest

import cases
import parity
from helpers import product

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    p = product()
    assert p.lib.edlibB200Available() == 1, "CUDA path unavailable: the product has no CPU fallback"
    return p


def test_known_and_golden(lib):
    parity.run_known(lib)
    assert parity.run_golden(lib) > 200


def test_single_pairs(lib):
    assert parity.run_single(lib, 21, 1500) == 1500


def test_batches_shared_targets(lib):
    assert parity.run_batches(lib, 22, 40) > 1000


def test_pairwise_batches_with_own_targets(lib):
    import cases
    assert parity.run_batches(lib, 17, 40, gen=cases.pairwise_cases) > 1500


def test_long_queries(lib):
    assert parity.run_single(lib, 23, 40, gen=cases.long_cases) == 40


def test_paths_beyond_the_stored_matrix_rule(lib):
    import cases
    assert parity.run_single(lib, 15, 60, gen=cases.path_cases) == 60


def test_candidate_filter_all_branches():
    """Same as the CPU test of that name, on the GPU: the filter forced on for small targets."""
    import subprocess
    from edlib_b200._ffi import REPO
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import parity, cases\n"
        "from helpers import product\n"
        "print(parity.run_batches(product(), 26, 40, gen=cases.filter_cases))\n"
    ) % (REPO, os.path.join(REPO, "tests"))
    for extra in ({}, {"EDLIB_B200_FILTER_K0": "4", "EDLIB_B200_FILTER_K1": "2", "EDLIB_B200_FILTER_SPREAD": "64",
                       "EDLIB_B200_FILTER_MAX_WINDOWS": "2", "EDLIB_B200_K1_MIN_CHUNK": "64"},
                  {"EDLIB_B200_FILTER_K1": "0", "EDLIB_B200_FILTER_SEED_K": "3", "EDLIB_B200_FILTER_SEED_BUCKET": "2"},
                  {"EDLIB_B200_FILTER_K0": "0", "EDLIB_B200_FILTER_K1": "12", "EDLIB_B200_FILTER_SEED_K": "0"},
                  {"EDLIB_B200_WINDOW_CHECK": "0"}, {"EDLIB_B200_WINDOW_CHECK": "-1", "EDLIB_B200_FILTER_SEED_LEVELS": "2"},
                  {"EDLIB_B200_FILTER_K0": "0", "EDLIB_B200_FILTER_K1": "0", "EDLIB_B200_FILTER_SEED_K": "40",
                   "EDLIB_B200_FILTER_SEED_LEVELS": "3", "EDLIB_B200_FILTER_SEED_SLACK": "100000"}):
        env = dict(os.environ, EDLIB_B200_FILTER_MIN_TARGET="128", EDLIB_B200_FILTER_MIN_LEVEL_READS="0", EDLIB_B200_K1_MIN_GROUP="4", **extra)
        out = subprocess.run(["python", "-c", code], env=env, check=True, capture_output=True, text=True)
        assert int(out.stdout.strip().splitlines()[-1]) > 500


def test_banded_nw_of_long_queries_on_the_band_kernel(lib):
    """k-banded NW sweeps of long queries on the thread-per-alignment band kernel, several window sizes per batch."""
    assert parity.run_batches(lib, 61, 24, gen=cases.band_cases) >= 80


def test_tightest_bounds_through_every_filter_path():
    """k = 0, 1, 2: exact and nearly exact reads through the seed levels (device- and host-driven), the prefix stages and
    the plain sweep, filter forced on for small targets (regression of the t = 0 early exit)."""
    import subprocess
    from edlib_b200._ffi import REPO
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import parity, cases\n"
        "from helpers import product\n"
        "print(parity.run_batches(product(), 101, 16, gen=cases.small_k_cases))\n"
    ) % (REPO, os.path.join(REPO, "tests"))
    for extra in ({}, {"EDLIB_B200_STREAM_MIN_PAIRS": "8"}, {"EDLIB_B200_DEVICE_STAGE": "0"}, {"EDLIB_B200_FILTER_SEED_K": "0"}):
        env = dict(os.environ, EDLIB_B200_FILTER_MIN_TARGET="128", EDLIB_B200_FILTER_MIN_LEVEL_READS="0", EDLIB_B200_K1_MIN_GROUP="4", **extra)
        out = subprocess.run(["python", "-c", code], env=env, check=True, capture_output=True, text=True)
        assert int(out.stdout.strip().splitlines()[-1]) > 500


def test_reads_that_tie_on_many_end_columns():
    """Homopolymer / tandem stretches (overflow list of the window sweeps), filter forced on for small targets: streamed
    device stage, host-driven seed levels, prefix stages alone."""
    import subprocess
    from edlib_b200._ffi import REPO
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import parity, cases\n"
        "from helpers import product\n"
        "print(parity.run_batches(product(), 51, 16, gen=cases.tied_ends_cases))\n"
    ) % (REPO, os.path.join(REPO, "tests"))
    for extra in ({"EDLIB_B200_STREAM_MIN_PAIRS": "8"}, {"EDLIB_B200_DEVICE_STAGE": "0", "EDLIB_B200_FILTER_SEED_BUCKET": "4096"},
                  {"EDLIB_B200_FILTER_SEED_K": "0", "EDLIB_B200_FILTER_K1": "12"}):
        env = dict(os.environ, EDLIB_B200_FILTER_MIN_TARGET="128", EDLIB_B200_FILTER_MIN_LEVEL_READS="0", EDLIB_B200_K1_MIN_GROUP="4", **extra)
        out = subprocess.run(["python", "-c", code], env=env, check=True, capture_output=True, text=True)
        assert int(out.stdout.strip().splitlines()[-1]) > 600


def test_read_sets_with_additional_equalities(lib):
    """Case-folding equalities (collapsed to one code per group: seed filter) and a wildcard on top (table path)."""
    assert parity.run_batches(lib, 91, 12, gen=cases.equality_read_cases) > 400


def test_queries_in_one_pinned_block_are_uploaded_directly():
    """Reads in ONE page-locked block (torch pin_memory): the device reads them from the caller's buffer, no staging
    copy -- streamed read sets and grouped batches; same results as ever."""
    import subprocess
    from edlib_b200._ffi import REPO
    from test_engine_emul import DIRECT_UPLOAD_CODE
    code = (DIRECT_UPLOAD_CODE % (REPO, os.path.join(REPO, "tests"))).replace("LOAD", "__import__('helpers').product()") \
        .replace("ALLOC", "(lambda shape: __import__('torch').zeros(shape, dtype=__import__('torch').uint8, pin_memory=True).numpy())")
    for extra in ({}, {"EDLIB_B200_DIRECT_UPLOAD": "0"}):
        env = dict(os.environ, EDLIB_B200_FILTER_MIN_TARGET="128", EDLIB_B200_FILTER_MIN_LEVEL_READS="0", EDLIB_B200_STREAM_MIN_PAIRS="8", EDLIB_B200_DIRECT_MIN_KB="1", **extra)
        out = subprocess.run(["python", "-c", code], env=env, check=True, capture_output=True, text=True)
        assert int(out.stdout.strip().splitlines()[-1]) == 1200


def test_target_handle():
    """edlibB200TargetPrepare on the GPU: same results with and without a resident target."""
    import subprocess
    from edlib_b200._ffi import REPO
    from test_engine_emul import TARGET_HANDLE_CODE
    code = (TARGET_HANDLE_CODE % (REPO, os.path.join(REPO, "tests"))).replace("LOAD", "__import__('helpers').product()")
    env = dict(os.environ, EDLIB_B200_FILTER_MIN_TARGET="128", EDLIB_B200_FILTER_MIN_LEVEL_READS="0", EDLIB_B200_STREAM_MIN_PAIRS="8")
    out = subprocess.run(["python", "-c", code], env=env, check=True, capture_output=True, text=True)
    assert int(out.stdout.strip().splitlines()[-1]) > 500


def test_large_alphabets_and_equalities(lib):
    """Protein-sized and full-byte alphabets over shared targets (per-thread Peq rows shrink the CTA or
    push the group to the warp kernel), with and without extra equalities."""
    import random
    from helpers import mutate, rand_seq
    chk = parity.checker()
    rng = random.Random(31)
    for asz in (20, 25, 64, 256):
        alpha = bytes(rng.sample(range(256), asz))
        t = rand_seq(rng, 4000, alpha)
        qs = []
        for _ in range(80):
            a = rng.randrange(0, len(t) - 300)
            qs.append(mutate(rng, t[a:a + rng.choice([40, 150, 250])], 0.08, alpha))
        eqs = [(bytes([alpha[0]]), bytes([alpha[1]])), (bytes([alpha[2]]), bytes([alpha[3]]))]
        for mode, task, e in ((2, 1, None), (1, 0, eqs), (0, 2, eqs), (2, 2, eqs)):
            st, res = lib.align_batch(qs, [t] * len(qs), -1, mode, task, e)
            assert st == 0
            for i in range(0, len(qs), 5):
                assert res[i] == chk.align(qs[i], t, -1, mode, task, e), (asz, mode, task, i)


def test_concurrent_callers(lib):
    """The ABI is re-entrant (called without the GIL by the reference binding): threads serialise safely."""
    import random
    from concurrent.futures import ThreadPoolExecutor
    from helpers import mutate, rand_seq
    chk = parity.checker()
    rng = random.Random(5)
    t = rand_seq(rng, 3000, b"ACGT")
    qs = [mutate(rng, t[a:a + 120], 0.05, b"ACGT") for a in range(0, 2400, 40)]
    exp = [chk.align(q, t, -1, 2, 1) for q in qs]
    with ThreadPoolExecutor(8) as ex:
        got = list(ex.map(lambda q: lib.align(q, t, -1, 2, 1), qs * 3))
    assert got == exp * 3
    # small calls run on side engines (own streams and locks) while a large batch holds the main engine
    big_q = [mutate(rng, t[a % 2800:a % 2800 + 150], 0.03, b"ACGT") for a in range(0, 40000, 7)]
    with ThreadPoolExecutor(9) as ex:
        fb = ex.submit(lambda: lib.align_batch(big_q, [t] * len(big_q), -1, 2, 0))
        got = list(ex.map(lambda q: lib.align(q, t, -1, 2, 2), qs * 4))
        st, res = fb.result()
    assert st == 0 and res[::97] == [chk.align(q, t, -1, 2, 0) for q in big_q[::97]]
    assert got == [chk.align(q, t, -1, 2, 2) for q in qs] * 4


def test_results_do_not_depend_on_the_plan():
    """Size-independent property on a config-2-shaped batch too large for the CPU checker: the same
    200k reads give identical results with every subset of the filter stages and under different chunkings."""
    import hashlib
    import subprocess
    from edlib_b200._ffi import REPO
    code = (
        "import sys, hashlib; sys.path.insert(0, %r)\n"
        "import numpy as np, ctypes as C\n"
        "import bench\n"
        "from edlib_b200 import workloads\n"
        "from edlib_b200._ffi import EdlibLib, AlignResult, make_config, product_path\n"
        "lib = EdlibLib(product_path(), has_batch=True)\n"
        "target, reads = workloads.reads_vs_target(200000, 150, 1000000, seed=11)\n"
        "qptr, qlen, tptr, tlen = bench.pointer_arrays(reads, target)\n"
        "cfg, _ = make_config(-1, 2, 1)\n"
        "res = np.zeros(len(reads), dtype=bench.RESULT_DTYPE)\n"
        "rc = lib.lib.edlibAlignBatch(bench.as_pp(qptr), bench.as_pi(qlen), bench.as_pp(tptr), bench.as_pi(tlen), len(reads), cfg,\n"
        "                             C.cast(res.ctypes.data, C.POINTER(AlignResult)))\n"
        "assert rc == 0\n"
        "h = hashlib.sha1()\n"
        "h.update(res['editDistance'].tobytes()); h.update(res['numLocations'].tobytes())\n"
        "for i in range(0, len(reads), 1):\n"
        "    n = int(res['numLocations'][i])\n"
        "    h.update(C.string_at(int(res['endLocations'][i]), 4 * n)); h.update(C.string_at(int(res['startLocations'][i]), 4 * n))\n"
        "print(h.hexdigest(), float(res['editDistance'].mean()))\n"
    ) % REPO
    digests = []
    off = {"EDLIB_B200_FILTER_SEED_K": "0", "EDLIB_B200_FILTER_K1": "0", "EDLIB_B200_FILTER_K0": "0"}
    for extra in ({}, {"EDLIB_B200_PACK_PARALLEL_KB": "1024"},                      # seed stages; threaded pack + upload
                  off,                                                              # no filter: plain sweeps
                  {"EDLIB_B200_FILTER_SEED_K": "0"},                                # prefix stages only
                  {"EDLIB_B200_FILTER_SEED_K": "0", "EDLIB_B200_FILTER_K1": "0", "EDLIB_B200_FILTER_K0": "5",
                   "EDLIB_B200_K1_MIN_CHUNK": "4096"},                              # tight 64-row stage, other chunking
                  {"EDLIB_B200_FILTER_SEED_K": "6", "EDLIB_B200_FILTER_K1": "0", "EDLIB_B200_FILTER_K0": "0"},  # seeds only, low t
                  dict(off, EDLIB_B200_K1_MIN_CHUNK="200000")):
        out = subprocess.run(["python", "-c", code], env=dict(os.environ, **extra), check=True, capture_output=True, text=True)
        digests.append(out.stdout.strip().split()[0])
    assert len(set(digests)) == 1, digests


def test_many_end_locations(lib):
    chk = parity.checker()
    for q, t, mode in [(b"A" * 64, b"B" * 70, 2), (b"A" * 10, b"A" * 300, 2), (b"AC" * 20, b"AC" * 200, 2),
                       (b"A" * 33, b"A" * 100, 1), (b"A" * 5, b"C" * 9, 2)]:
        for task in (0, 1, 2):
            assert lib.align(q, t, -1, mode, task) == chk.align(q, t, -1, mode, task)
    qs = [b"A" * 10] * 400
    t = b"A" * 5000
    st, res = lib.align_batch(qs, [t] * 400, -1, 2, 1)
    exp = chk.align(qs[0], t, -1, 2, 1)
    assert st == 0 and all(r == exp for r in res)


def test_config2_shape_sample(lib):
    """150 bp reads with 3 % sub/ins/del against one shared synthetic target, HW distance
    (BASELINE.json configs[1] shape, scaled to what the checker finishes in seconds)."""
    import numpy as np
    from edlib_b200 import workloads
    target, reads = workloads.reads_vs_target(num_reads=3000, read_len=150, target_len=400_000, seed=5)
    t = target.tobytes()
    qs = [r.tobytes() for r in reads]
    st, res = lib.align_batch(qs, [t] * len(qs), -1, 2, 0)
    assert st == 0
    chk = parity.checker()
    for i in range(0, len(qs), 7):
        assert res[i] == chk.align(qs[i], t, -1, 2, 0), i
    eds = np.array([r["editDistance"] for r in res])
    assert 2.0 < eds.mean() < 8.0


def test_config3_shape_every_field(lib):
    """BASELINE configs[2] shape: 10 kbp queries vs their 3 %-mutated copies, NW, k = 500, EDLIB_TASK_LOC -- 2,000 pairs,
    every result field of every pair against the reference build."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from edlib_b200 import workloads
    genome = workloads.random_dna(2_000_000, 9)
    qbuf, tbuf, tlens = workloads.long_pairs_packed(genome, 2000, 10_000, seed=43)
    qs = [qbuf[i].tobytes() for i in range(2000)]
    ts = [tbuf[i, :tlens[i]].tobytes() for i in range(2000)]
    st, res = lib.align_batch(qs, ts, 500, 0, 1)
    assert st == 0
    chk = parity.checker()
    with ThreadPoolExecutor(16) as ex:
        exp = list(ex.map(lambda i: chk.align(qs[i], ts[i], 500, 0, 1), range(2000)))
    assert res == exp
    eds = np.array([r["editDistance"] for r in res])
    assert (eds >= 0).mean() > 0.99 and 200 < eds[eds >= 0].mean() < 400


def test_config4_shape_every_field(lib):
    """BASELINE configs[3] shape: 150 bp reads (3 % sub/ins/del) vs one shared target, HW, EDLIB_TASK_PATH -- 50,000 reads
    through the streamed path, every field (alignment bytes included) of every 10th read against the reference build,
    plus the batched CIGAR helper against edlibAlignmentToCigar of the reference."""
    import ctypes as C
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    import bench
    from edlib_b200 import workloads
    from edlib_b200._ffi import AlignResult, make_config
    target, reads = workloads.reads_vs_target(num_reads=50_000, read_len=150, target_len=200_000, seed=8)
    qptr, qlen, tptr, tlen = bench.pointer_arrays(reads, target)
    cfg, _ = make_config(-1, 2, 2)
    res = np.zeros(len(reads), dtype=bench.RESULT_DTYPE)
    assert lib.lib.edlibAlignBatch(bench.as_pp(qptr), bench.as_pi(qlen), bench.as_pp(tptr), bench.as_pi(tlen), len(reads), cfg,
                                   C.cast(res.ctypes.data, C.POINTER(AlignResult))) == 0
    cg = (C.c_void_p * len(reads))()
    lib.lib.edlibB200AlignmentsToCigar.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    assert lib.lib.edlibB200AlignmentsToCigar(res.ctypes.data, len(reads), 1, cg) == 0
    chk = parity.checker()
    t = target.tobytes()
    idx = list(range(0, len(reads), 10))
    with ThreadPoolExecutor(16) as ex:
        exp = list(ex.map(lambda i: chk.align(reads[i].tobytes(), t, -1, 2, 2), idx))
    libc = C.CDLL(None)
    libc.free.argtypes = [C.c_void_p]
    for i, e in zip(idx, exp):
        assert bench.gpu_result_dict(res, i) == e, i
        assert C.string_at(cg[i]).decode() == chk.cigar(e["alignment"]), i
    for p in cg:
        if p:
            libc.free(p)
    lib.lib.edlibB200FreeResults.argtypes = [C.c_void_p, C.c_int]
    lib.lib.edlibB200FreeResults(res.ctypes.data, len(reads))
