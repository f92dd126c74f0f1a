"""Generates tests/golden/ref_vectors.json by running the UNMODIFIED reference build
(oracle/_ref/libedlib_ref.so, compiled from /root/reference by `make -C oracle ref`) on seeded
inputs.  Run here (the reference is not available on the GPU box); the JSON is committed.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import cases  # noqa: E402
from helpers import ref  # noqa: E402
from edlib_b200._ffi import MODES, TASKS  # noqa: E402


def main():
    R = ref()
    out = []

    def add(q, t, k, mode, task, eqs):
        r = R.align(q, t, k, mode, task, eqs)
        if r.get("alignment") is not None:
            r["alignment"] = r["alignment"].hex()
        out.append(dict(q=q.hex(), t=t.hex(), k=k, mode=mode, task=task,
                        eqs=[[a.hex(), b.hex()] for a, b in eqs] if eqs else None, expect=r))

    for q, t, mode, task, k, eqs, _ in cases.KNOWN:
        add(q, t, k, MODES[mode], TASKS[task], eqs)
    for c in cases.single_pair_cases(2024, 400):
        if len(c["q"]) + len(c["t"]) <= 700:
            add(c["q"], c["t"], c["k"], c["mode"], c["task"], c["eqs"])
    with open(os.path.join(HERE, "ref_vectors.json"), "w") as f:
        json.dump(dict(source="reference edlib v1.2.6 (commit 0ddc23e) built by oracle/Makefile", cases=out), f)
    print(len(out), "cases")


if __name__ == "__main__":
    main()
