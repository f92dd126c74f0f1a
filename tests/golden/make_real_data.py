#!/usr/bin/env python
"""Fixtures from the reference's own test DATA (not sources), so that real-data parity and the bench run on the GPU
box, where /root/reference does not exist:

  edlib_b200/data/e_coli_DH1.2bit    the E. coli DH1 genome of test_data/E_coli_DH1/e_coli_DH1.fasta (4,630,707 bp,
                                     only ACGT), four bases per byte (A=0 C=1 G=2 T=3, first base in the low bits),
                                     preceded by its length as a little-endian uint32
  tests/golden/ecoli_reads.json      every read / mutated prefix shipped next to it (file name -> sequence) and the
                                     reference build's answers for them (HW, EDLIB_TASK_LOC, k = -1) -- generated
                                     with oracle/_ref/libedlib_ref.so

Run here (container with /root/reference): python tests/golden/make_real_data.py
"""
import glob
import json
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.dirname(HERE))
DATA = "/root/reference/test_data/E_coli_DH1"


def read_fasta(path):
    seqs, cur = [], []
    with open(path, "rb") as f:
        for line in f:
            if line.startswith(b">"):
                if cur:
                    seqs.append(b"".join(cur))
                cur = []
            else:
                cur.append(line.strip())
    if cur:
        seqs.append(b"".join(cur))
    return seqs


def main():
    from helpers import ref
    genome = read_fasta(os.path.join(DATA, "e_coli_DH1.fasta"))[0]
    g = np.frombuffer(genome, dtype=np.uint8)
    lut = np.full(256, 255, dtype=np.uint8)
    for i, ch in enumerate(b"ACGT"):
        lut[ch] = i
    codes = lut[g]
    assert codes.max() < 4
    pad = (-len(codes)) % 4
    c4 = np.concatenate([codes, np.zeros(pad, dtype=np.uint8)]).reshape(-1, 4)
    packed = (c4[:, 0] | (c4[:, 1] << 2) | (c4[:, 2] << 4) | (c4[:, 3] << 6)).astype(np.uint8)
    out = os.path.join(REPO, "edlib_b200", "data", "e_coli_DH1.2bit")
    with open(out, "wb") as f:
        f.write(struct.pack("<I", len(codes)))
        f.write(packed.tobytes())
    print("wrote", out, len(codes), "bp,", os.path.getsize(out), "bytes")

    files = sorted(glob.glob(os.path.join(DATA, "mason_illumina_reads", "*", "*.fasta")) +
                   glob.glob(os.path.join(DATA, "prefixes", "*", "*.fasta")))
    lib = ref()
    reads = {}
    for f in files:
        s = read_fasta(f)[0]
        r = lib.align(s, genome, -1, 2, 1)
        reads[os.path.relpath(f, DATA)] = {"seq": s.decode("ascii"), "editDistance": r["editDistance"],
                                           "endLocations": r["endLocations"], "startLocations": r["startLocations"],
                                           "alphabetLength": r["alphabetLength"]}
        print(os.path.relpath(f, DATA), len(s), r["editDistance"], r["endLocations"][:4])
    with open(os.path.join(HERE, "ecoli_reads.json"), "w") as f:
        json.dump({"genome_length": len(genome), "mode": "HW", "task": "locations", "k": -1, "reads": reads}, f, indent=0)


if __name__ == "__main__":
    main()
