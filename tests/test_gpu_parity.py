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
    for extra in ({}, {"EDLIB_B200_FILTER_K0": "4", "EDLIB_B200_FILTER_SPREAD": "64", "EDLIB_B200_K1_MIN_CHUNK": "64"}):
        env = dict(os.environ, EDLIB_B200_FILTER_MIN_TARGET="128", EDLIB_B200_K1_MIN_GROUP="4", **extra)
        out = subprocess.run(["python", "-c", code], env=env, check=True, capture_output=True, text=True)
        assert int(out.stdout.strip().splitlines()[-1]) > 500


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
