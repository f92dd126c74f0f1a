"""Python host mirror (edlib_b200.align / align_batch / getNiceAlignment) against the known answers of the
reference binding's own test (bindings/python/test.py:6-80)."""
import pytest

import edlib_b200


def test_nice_alignment_formatting_only():
    # pure formatting (no device): the reference's expected strings for its NW/HW/SHW example
    res = {"locations": [(0, 14)], "cigar": "1I4=2I4=1X3=1D2="}
    nice = edlib_b200.getNiceAlignment(res, "TAAGGATGGTCCCATTC", "AAGGGGTCTCATATC")
    assert nice["query_aligned"] == "TAAGGATGGTCCCAT-TC"
    assert nice["matched_aligned"] == "-||||--||||.|||-||"
    assert nice["target_aligned"] == "-AAGG--GGTCTCATATC"
    with pytest.raises(Exception):
        edlib_b200.getNiceAlignment({"locations": [(0, 1)], "cigar": None}, "A", "A")


def check_known_answers():
    a = edlib_b200.align
    assert a("telephone", "elephant")["editDistance"] == 3
    assert a(b"telephone", b"elephant")["editDistance"] == 3
    assert a("ACTG", "CACTRT", mode="HW", task="path", additionalEqualities=[("R", "A"), ("R", "G")])["editDistance"] == 0
    for mode in ("NW", "HW", "SHW"):
        r = a(query="TAAGGATGGTCCCATTC", target="AAGGGGTCTCATATC", mode=mode, task="path")
        nice = edlib_b200.getNiceAlignment(r, "TAAGGATGGTCCCATTC", "AAGGGGTCTCATATC")
        assert nice == {"query_aligned": "TAAGGATGGTCCCAT-TC", "matched_aligned": "-||||--||||.|||-||",
                        "target_aligned": "-AAGG--GGTCTCATATC"}
    assert a("TAAGGATGGTCCCATTC", "AAGGGGTCTCATATC", mode="NW", task="distance")["cigar"] is None
    assert a("", "elephant")["editDistance"] == 8 and a("telephone", "")["editDistance"] == 9
    assert a("", "elephant", mode="HW")["editDistance"] == 0 and a("telephone", "", mode="HW")["editDistance"] == 9
    assert a("", "elephant", mode="SHW")["editDistance"] == 0 and a("telephone", "", mode="SHW")["editDistance"] == 9
    r = a("ты милая", "ты гений")
    assert r["editDistance"] == 5 and r["alphabetLength"] == 12
    alpha = "".join(chr(i) for i in range(1, 257))
    assert a(alpha * 3, alpha + alpha[::-1] + alpha)["editDistance"] == 256
    assert a("telephone", "elephant", task="path") == {"editDistance": 3, "alphabetLength": 8, "locations": [(0, 7)],
                                                        "cigar": "1I5=1X1=1X"}


def check_batch_matches_align():
    t = "ACGTTGCAATGCCGTAAGGCTTAACGGATCCA" * 20
    qs = ["TTGCAATGC", "GGGGGGGG", "AAGGCTTAACGG", ""]
    got = edlib_b200.align_batch(qs, t, mode="HW", task="path")
    assert got == [edlib_b200.align(q, t, mode="HW", task="path") for q in qs]
    # one target per query, sequences of arbitrary hashables (recoded over the joint alphabet), equalities
    ts = [t, "GGGGAGGG", t[::-1], "ACGT"]
    got = edlib_b200.align_batch(qs, ts, mode="NW", task="path")
    assert got == [edlib_b200.align(q, x, mode="NW", task="path") for q, x in zip(qs, ts)]
    words = [("tele", "phone", "x"), ("ele", "phant")]
    r = edlib_b200.align(words[0], words[1], task="path", additionalEqualities=[("tele", "ele")])
    assert r["editDistance"] == 2 and r["alphabetLength"] == 5 and r["cigar"] == "1=1X1I"


@pytest.mark.gpu
def test_reference_binding_known_answers():
    check_known_answers()


@pytest.mark.gpu
def test_align_batch_matches_align():
    check_batch_matches_align()


def test_mirror_host_logic_over_the_emulated_kernels(monkeypatch):
    """The same checks with the package's library handle pointed at the CPU emulation build (the marshalling, recoding
    and result shaping of edlib_b200/__init__.py are host code; the product library itself needs the GPU)."""
    import ctypes as C
    from test_engine_emul import load_emul
    lib = load_emul()
    lib.lib.edlibB200LastError.restype = C.c_char_p
    monkeypatch.setattr(edlib_b200, "_lib", lib)
    check_known_answers()
    check_batch_matches_align()
    # an unknown mode name leaves the default (NW), as in edlib.pyx:102-104; more than 256 distinct values cannot be recoded
    assert edlib_b200.align("AAAA", "AACA", mode="no-such-mode") == edlib_b200.align("AAAA", "AACA", mode="NW")
    with pytest.raises(ValueError):
        edlib_b200.align(list(range(200)), list(range(100, 400)))
