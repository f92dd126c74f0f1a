"""Parity drivers shared by the CPU-emulation tests and the GPU tests: run seeded cases through
an implementation of the C ABI and compare EVERY field of EdlibAlignResult bit-exactly with a
checker (the reference build when oracle/_ref is present, else the oracle restatement)."""
import json
import os

import cases
from helpers import have_ref, oracle, ref

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_vectors.json")


def checker():
    return ref() if have_ref() else oracle()


def run_single(lib, seed, count, gen=cases.single_pair_cases):
    chk = checker()
    n = 0
    for c in gen(seed, count):
        got = lib.align(c["q"], c["t"], c["k"], c["mode"], c["task"], c["eqs"])
        exp = chk.align(c["q"], c["t"], c["k"], c["mode"], c["task"], c["eqs"])
        assert got == exp, dict(case={k: (v if not isinstance(v, bytes) else v[:60]) for k, v in c.items()},
                                got=str(got)[:400], exp=str(exp)[:400])
        n += 1
    return n


def run_batches(lib, seed, count, gen=cases.batch_cases):
    chk = checker()
    n = 0
    for c in gen(seed, count):
        st, res = lib.align_batch(c["qs"], c["ts"], c["k"], c["mode"], c["task"], c["eqs"])
        assert st == 0
        for i, (q, t) in enumerate(zip(c["qs"], c["ts"])):
            exp = chk.align(q, t, c["k"], c["mode"], c["task"], c["eqs"])
            assert res[i] == exp, dict(pair=i, k=c["k"], mode=c["mode"], task=c["task"], eqs=c["eqs"],
                                       m=len(q), n=len(t), got=str(res[i])[:400], exp=str(exp)[:400])
            n += 1
    return n


def run_golden(lib):
    with open(GOLDEN) as f:
        gold = json.load(f)
    for c in gold["cases"]:
        eqs = [(bytes.fromhex(a), bytes.fromhex(b)) for a, b in c["eqs"]] if c["eqs"] else None
        r = lib.align(bytes.fromhex(c["q"]), bytes.fromhex(c["t"]), c["k"], c["mode"], c["task"], eqs)
        exp = dict(c["expect"])
        exp["alignment"] = bytes.fromhex(exp["alignment"]) if exp.get("alignment") is not None else None
        assert r == exp, c
    return len(gold["cases"])


def run_known(lib):
    from edlib_b200._ffi import MODES, TASKS
    for q, t, mode, task, k, eqs, exp in cases.KNOWN:
        r = lib.align(q, t, k, MODES[mode], TASKS[task], eqs)
        for key, val in exp.items():
            if key == "cigar":
                assert lib.cigar(r["alignment"]) == val
            else:
                assert r[key] == val, (q, t, mode, key, r)
