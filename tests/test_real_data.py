"""Real-data parity on the reference's own test files (E. coli DH1 genome, 4.6 Mbp): the reads and mutated
prefixes shipped under /root/reference/test_data go through the engine (CPU emulation of the kernels) as ONE
HW batch over the shared genome -- the config-2 shape, with the candidate filter at its production settings --
and must reproduce SURVEY.md section 8c's values measured from the reference and the live reference build.
Skipped where the reference tree is absent (the GPU box)."""
import glob
import os

import pytest

import parity
from test_engine_emul import load_emul

DATA = "/root/reference/test_data/E_coli_DH1"

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(DATA, "e_coli_DH1.fasta")),
                                reason="reference test data not present")


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


# SURVEY.md 8c, measured from the reference: file (relative to DATA) -> (editDistance, end locations)
GOLDEN = {
    "mason_illumina_reads/50bp/e_coli_DH1_illumina_1x50.fasta": (0, [2646428]),
    "mason_illumina_reads/100bp/e_coli_DH1_illumina_1x100.fasta": (0, [2646478]),
    "mason_illumina_reads/250bp/e_coli_DH1_illumina_1x250.fasta": (0, [2646628]),
    "mason_illumina_reads/50bp/mutated_80_perc.fasta": (13, [2646427, 2646429, 2646430, 2646431]),
    "mason_illumina_reads/100bp/mutated_60_perc.fasta": (37, [2646472, 2646473, 3531981]),
}


def test_ecoli_reads_and_mutated_prefixes_hw_locations():
    genome = read_fasta(os.path.join(DATA, "e_coli_DH1.fasta"))[0]
    assert len(genome) == 4630707
    files = sorted(glob.glob(os.path.join(DATA, "mason_illumina_reads", "*", "*.fasta")) +
                   glob.glob(os.path.join(DATA, "prefixes", "*", "*.fasta")))
    names, reads = [], []
    for f in files:
        for s in read_fasta(f)[:1]:
            if 0 < len(s) <= 256:  # the lane-per-alignment path with the candidate filter
                names.append(os.path.relpath(f, DATA))
                reads.append(s)
    assert len(reads) >= 20
    lib = load_emul()
    st, res = lib.align_batch(reads, [genome] * len(reads), -1, 2, 1)
    assert st == 0
    got = dict(zip(names, res))
    for name, (ed, ends) in GOLDEN.items():
        assert name in got, name
        assert got[name]["editDistance"] == ed and got[name]["endLocations"] == ends, (name, got[name])
    chk = parity.checker()
    for name, read, r in zip(names, reads, res):
        assert r == chk.align(read, genome, -1, 2, 1), name


def test_phage_nw_distances_config1():
    """BASELINE configs[0] and its siblings: mutated phage genomes vs the phage genome (94 kbp), NW, k = -1
    (band doubling on the warp kernel): SURVEY.md 8c's measured reference values."""
    d = "/root/reference/test_data/Enterobacteria_Phage_1/"
    target = read_fasta(d + "Enterobacteria_phage_1.fasta")[0]
    assert len(target) == 94481
    lib = load_emul()
    for pc, ed in ((99, 990), (97, 2977), (94, 6042), (90, 9506)):
        q = read_fasta(d + "mutated_%d_perc.fasta" % pc)[0]
        r = lib.align(q, target, -1, 0, 0)
        assert r["editDistance"] == ed and r["endLocations"] == [94480] and r["startLocations"] is None, (pc, r)
