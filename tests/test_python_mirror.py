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


@pytest.mark.gpu
def test_reference_binding_known_answers():
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


@pytest.mark.gpu
def test_align_batch_matches_align():
    t = "ACGTTGCAATGCCGTAAGGCTTAACGGATCCA" * 20
    qs = ["TTGCAATGC", "GGGGGGGG", "AAGGCTTAACGG", ""]
    got = edlib_b200.align_batch(qs, t, mode="HW", task="path")
    assert got == [edlib_b200.align(q, t, mode="HW", task="path") for q in qs]
