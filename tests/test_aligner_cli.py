"""The batched command-line aligner (apps/aligner/aligner.cpp over edlibAlignBatch) must print what the
reference's edlib-aligner prints (reference apps/aligner/aligner.cpp, built unmodified into
oracle/_ref/edlib-aligner_ref by `make -C oracle reftests`), for every option combination, except the
progress counter and the timing line."""
import os
import random
import subprocess

import pytest

from edlib_b200._ffi import REPO
from helpers import mutate, rand_seq

REFBIN = os.path.join(REPO, "oracle", "_ref")
OPTION_SETS = [["-m", "HW"], ["-m", "NW", "-l"], ["-m", "SHW", "-l", "-k", "30"], ["-m", "HW", "-p", "-f", "CIG_EXT"],
               ["-m", "HW", "-p"], ["-m", "NW", "-p", "-f", "CIG_STD"], ["-m", "HW", "-n", "5", "-l"],
               ["-m", "HW", "-n", "3", "-k", "8", "-p", "-f", "CIG_EXT"], ["-m", "SHW", "-n", "2"],
               [], ["-s"], ["-m", "HW", "-s", "-p"], ["-m", "HW", "-r", "2", "-l"], ["-m", "SHW", "-p", "-f", "CIG_STD", "-k", "40"],
               ["-m", "NW", "-p", "-f", "NICE", "-k", "0"], ["-m", "HW", "-k", "0", "-l"], ["-m", "NW", "-n", "1", "-p"]]
# wrong usage: same messages and exit codes (argv[0] differs in the usage text)
ERROR_SETS = [["-m", "XX"], ["-f", "SAM", "-p"], ["-m", "HW", "no-such-queries.fasta"], ["-m", "HW", "Q", "no-such-target.fasta"], ["-m", "HW", "onlyone"],
              ["-z"]]


def write_fasta(path, seqs):
    with open(path, "w") as f:
        for i, s in enumerate(seqs):
            f.write(">seq%d some description\n" % i)
            for a in range(0, len(s), 70):
                f.write(s[a:a + 70].decode() + "\n")


def make_inputs(tmp):
    rng = random.Random(77)
    t = rand_seq(rng, 20000, b"ACGT")
    qs = []
    for i in range(60):
        a = rng.randrange(0, len(t) - 400)
        qs.append(mutate(rng, t[a:a + rng.choice([60, 150, 300])], rng.choice([0.0, 0.03, 0.1]), b"ACGT"))
    qs.append(rand_seq(rng, 120, b"ACGT"))
    qf, tf = os.path.join(tmp, "q.fasta"), os.path.join(tmp, "t.fasta")
    write_fasta(qf, qs)
    write_fasta(tf, [t])
    return qf, tf


def normalised(exe, opts, qf, tf):
    out = subprocess.run([exe] + opts + [qf, tf], capture_output=True, text=True, check=True).stdout
    lines = out.replace("\r", "\n").split("\n")
    return [l for l in lines if not l.startswith("Cpu time") and not (l and l.replace("/", "").isdigit())]


def compare(mine, tmp_path):
    ref = os.path.join(REFBIN, "edlib-aligner_ref")
    if not (os.path.exists(ref) and os.path.exists(mine)):
        pytest.skip("aligner binaries not built")
    qf, tf = make_inputs(str(tmp_path))
    for opts in OPTION_SETS:
        assert normalised(mine, opts, qf, tf) == normalised(ref, opts, qf, tf), opts
    # FASTA reading: wrapped lines, blank lines, lower case and other letters, blanks inside a line, an empty record, CRLF line
    # ends, no header, an empty file, a header alone, no newline at the end, several records in the target file
    odd = {"odd": ">q1\nACGTACGTAC\nGTAC\n\n>q2 desc\nacgtnnACGT\n>q3\n\n>q4\nAC GT\tAC\n", "crlf": ">q1\r\nACGTAC\r\nGT\r\n>q2\r\nTTGACC\r\n",
           "nohdr": "ACGT\nACGT\n", "empty": "", "hdr": ">only header\n", "noeol": ">q1\nACGT"}
    small_t = os.path.join(str(tmp_path), "small_t.fasta")
    with open(small_t, "w") as f:
        f.write(">t\nACGTACGTACGTACGGTACCAGT\nACGTTTGACCA\n")
    for name, text in odd.items():
        path = os.path.join(str(tmp_path), name + ".fasta")
        with open(path, "w", newline="") as f:
            f.write(text)
        for opts in (["-m", "HW", "-l"], ["-m", "NW", "-p", "-f", "CIG_EXT"]):
            got, exp = [subprocess.run([exe] + opts + [path, small_t], capture_output=True, text=True) for exe in (mine, ref)]
            strip = lambda r: [l for l in r.stdout.replace("\r", "\n").split("\n")  # noqa: E731
                               if not l.startswith("Cpu time") and not (l and l.replace("/", "").isdigit())]
            assert (got.returncode, strip(got)) == (exp.returncode, strip(exp)), (name, opts)
    assert normalised(mine, ["-m", "HW"], small_t, os.path.join(str(tmp_path), "odd.fasta")) == \
        normalised(ref, ["-m", "HW"], small_t, os.path.join(str(tmp_path), "odd.fasta"))  # first record of a multi-record target file
    for opts in ERROR_SETS:
        runs = []
        for exe in (mine, ref):
            args = [qf if a == "Q" else a for a in opts]
            if not any(a.endswith(".fasta") or a == "onlyone" for a in opts):
                args += [qf, tf]
            elif opts[-1] == "no-such-queries.fasta":
                args += [tf]
            r = subprocess.run([exe] + args, capture_output=True, text=True)
            usage = [l for l in r.stderr.replace(exe, "PROG").splitlines() if l.startswith("Usage:")]  # the option help is worded differently
            runs.append((r.returncode, r.stdout, usage))
        assert runs[0] == runs[1], opts


def test_aligner_matches_reference_on_emulated_kernels(tmp_path):
    if os.path.isdir("/root/reference"):
        subprocess.run(["make", "-s", "-C", os.path.join(REPO, "tests", "emul")], check=True)
        subprocess.run(["make", "-s", "-C", os.path.join(REPO, "oracle"), "reftests"], check=True)
    compare(os.path.join(REFBIN, "edlib-aligner_emul"), tmp_path)


@pytest.mark.gpu
def test_aligner_matches_reference_on_gpu(tmp_path):
    compare(os.path.join(REPO, "edlib_b200", "lib", "edlib-aligner"), tmp_path)
