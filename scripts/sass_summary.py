#!/usr/bin/env python
"""SASS evidence of the built library (no GPU needed): per-kernel instruction counts and the mnemonics that matter, plus
full listings of named kernels.  Usage: python scripts/sass_summary.py [substring-of-a-kernel-name[=file-name-part] ...]
Writes profiles/r02_sass_summary.txt and one profiles/r02_sass_<name>.txt per requested kernel."""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "edlib_b200", "lib", "libedlib_b200.so")
KEYS = ["UBLKCP", "SYNCS", "LOP3", "IADD3", "IMAD", "SHF", "LDS", "STS", "LDG", "STG", "SHFL", "VOTE", "POPC", "ATOMG", "ATOMS", "LDL", "STL"]


def kernels():
    text = subprocess.run(["cuobjdump", "-sass", LIB], check=True, capture_output=True, text=True).stdout
    name, body = None, []
    for line in text.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                yield name, body
            name, body = m.group(1), []
        elif name:
            body.append(line)
    if name:
        yield name, body


def main():
    wanted = sys.argv[1:]
    out = ["# cuobjdump -sass edlib_b200/lib/libedlib_b200.so (sm_100a): instructions per kernel and the mnemonics that matter",
           "# (UBLKCP = 1-D TMA bulk copy, SYNCS = mbarrier ops, LOP3/IADD3 = the bit-vector recurrences, LDS/STS = shared memory, LDL/STL = local memory)"]
    for name, body in kernels():
        ops = collections.Counter()
        n = 0
        for line in body:
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
            if not m:
                continue
            n += 1
            ops[m.group(1)] += 1
        keys = " ".join("%s=%d" % (k, ops[k]) for k in KEYS if ops[k])
        out.append("%-72s %5d  %s" % (name[:72], n, keys))
        for w in wanted:  # "substring" or "substring=file-name-part"
            sub, _, short = w.partition("=")
            if sub in name:
                short = short or re.sub(r"[^A-Za-z0-9]+", "_", sub).strip("_")
                with open(os.path.join(REPO, "profiles", "r02_sass_%s.txt" % short), "w") as f:
                    lines = []
                    for line in body:  # instruction text only (no encodings)
                        m = re.match(r"\s*(/\*[0-9a-f]{4,}\*/)\s+(.*?;)", line)
                        if m:
                            lines.append(m.group(1) + " " + m.group(2))
                        elif ".headerflags" in line or re.match(r"\s*\.L_", line):
                            lines.append(line.strip())
                    f.write("Function : %s\n" % name + "\n".join(lines) + "\n")
    with open(os.path.join(REPO, "profiles", "r02_sass_summary.txt"), "w") as f:
        f.write("\n".join(out) + "\n")
    print(len(out) - 2, "kernels")


if __name__ == "__main__":
    main()
