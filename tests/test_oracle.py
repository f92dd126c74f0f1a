"""CPU tests: the oracle restatement is pinned against the reference (oracle/_ref, built from
/root/reference when present) and against the committed golden vectors made from it."""
import json
import os

import pytest

import cases
from helpers import have_ref, oracle, ref
from edlib_b200._ffi import MODES, TASKS

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_vectors.json")


def test_known_answers_oracle():
    O = oracle()
    for q, t, mode, task, k, eqs, exp in cases.KNOWN:
        r = O.align(q, t, k, MODES[mode], TASKS[task], eqs)
        for key, val in exp.items():
            if key == "cigar":
                assert O.cigar(r["alignment"]) == val
            else:
                assert r[key] == val, (q, t, mode, key, r)


def test_cigar_formats():
    O = oracle()
    ops = bytes([0, 0, 1, 1, 1, 2, 1, 1, 3, 0, 0])  # runTests.cpp:506-533
    assert O.cigar(ops, 1) == "2=3I1D2I1X2="
    assert O.cigar(ops, 0) == "2M3I1D2I3M"
    assert O.cigar(b"", 1) == ""
    assert O.cigar(bytes([0, 4]), 1) is None


def test_golden_vectors_oracle():
    O = oracle()
    with open(GOLDEN) as f:
        gold = json.load(f)
    assert len(gold["cases"]) > 200
    for c in gold["cases"]:
        eqs = [(bytes.fromhex(a), bytes.fromhex(b)) for a, b in c["eqs"]] if c["eqs"] else None
        r = O.align(bytes.fromhex(c["q"]), bytes.fromhex(c["t"]), c["k"], c["mode"], c["task"], eqs)
        exp = dict(c["expect"])
        exp["alignment"] = bytes.fromhex(exp["alignment"]) if exp.get("alignment") is not None else None
        assert r == exp, c


@pytest.mark.skipif(not have_ref(), reason="reference build oracle/_ref not present")
def test_oracle_vs_reference_random():
    O, R = oracle(), ref()
    n = 0
    for c in cases.single_pair_cases(101, 4000):
        assert O.align(c["q"], c["t"], c["k"], c["mode"], c["task"], c["eqs"]) == \
            R.align(c["q"], c["t"], c["k"], c["mode"], c["task"], c["eqs"]), c
        n += 1
    assert n == 4000


@pytest.mark.skipif(not have_ref(), reason="reference build oracle/_ref not present")
def test_oracle_vs_reference_hirschberg_regime():
    """Paths beyond the reference's 1 MiB stored-matrix rule (edlib.cpp:1188-1211)."""
    import random
    from helpers import mutate, rand_seq
    O, R = oracle(), ref()
    rng = random.Random(7)
    for _ in range(40):
        alpha = bytes(rng.sample(range(256), rng.choice([2, 4, 10])))
        if rng.random() < 0.5:
            q = rand_seq(rng, rng.randrange(50, 350), alpha)
            t = rand_seq(rng, rng.randrange(9000, 14000), alpha)
        else:
            t = rand_seq(rng, rng.randrange(1500, 4000), alpha)
            q = mutate(rng, t, rng.choice([0.02, 0.2]), alpha)
        mode = rng.choice([0, 0, 1, 2])
        assert O.align(q, t, -1, mode, 2) == R.align(q, t, -1, mode, 2)
