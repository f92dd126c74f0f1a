"""edlib_b200 -- B200-native batched edit-distance engine behind the edlib C ABI.

Python host mirror of the reference's binding (bindings/python/edlib.pyx): `align()` takes the same
arguments and returns the same dict (edlib.pyx:56-155), `getNiceAlignment()` is edlib.pyx:157-238,
`align_batch()` is the batched form over `edlibAlignBatch`.  All computation happens in the CUDA library
edlib_b200/lib/libedlib_b200.so; importing this package without it (or without a GPU) works, calling
`align*` then raises.
"""
import ctypes as _C
import re as _re

from ._ffi import (EDLIB_CIGAR_EXTENDED, EDLIB_CIGAR_STANDARD, EDLIB_STATUS_OK, MODES, TASKS, EdlibLib,
                   product_path)

__all__ = ["align", "align_batch", "align_many", "getNiceAlignment", "library"]

_lib = None


def library():
    """The loaded product library (raises OSError if it has not been built)."""
    global _lib
    if _lib is None:
        _lib = EdlibLib(product_path(), prefix="edlib", has_batch=True)
        _lib.lib.edlibB200LastError.restype = _C.c_char_p
    return _lib


def _is_plain(s):
    """bytes, or str whose UTF-8 form has one byte per character (edlib.pyx:11-19)."""
    return isinstance(s, (bytes, bytearray)) or (isinstance(s, str) and len(s.encode("utf-8")) == len(s))


def _plain_bytes(s):
    return bytes(s) if isinstance(s, (bytes, bytearray)) else s.encode("utf-8")


def _map_to_bytes(seqs, additional_equalities):
    """Sequences of hashables -> byte strings (edlib.pyx:22-53): ASCII str / bytes pass through,
    anything else is recoded over the joint alphabet (at most 256 distinct values)."""
    if all(_is_plain(s) for s in seqs):
        eqs = None
        if additional_equalities is not None:
            eqs = [(_plain_bytes(a)[:1], _plain_bytes(b)[:1]) for a, b in additional_equalities]
        return [_plain_bytes(s) for s in seqs], eqs
    alphabet = set()
    for s in seqs:
        alphabet.update(s)
    if len(alphabet) > 256:
        raise ValueError("query and target combined have more than 256 unique values, this is not supported.")
    code = {c: bytes([i]) for i, c in enumerate(alphabet)}
    eqs = None
    if additional_equalities is not None:
        eqs = [(code[a], code[b]) for a, b in additional_equalities if a in code and b in code]
    return [b"".join(code[c] for c in s) for s in seqs], eqs


def _result(d, want_cigar):
    if d["status"] != EDLIB_STATUS_OK:
        raise Exception("There was an error. (" + library().lib.edlibB200LastError().decode() + ")")
    locations = []
    if d["endLocations"] is not None:
        starts = d["startLocations"] or [None] * d["numLocations"]
        locations = list(zip(starts, d["endLocations"]))
    cigar = library().cigar(d["alignment"], EDLIB_CIGAR_EXTENDED) if (want_cigar and d["alignment"] is not None) else None
    return {"editDistance": d["editDistance"], "alphabetLength": d["alphabetLength"],
            "locations": locations, "cigar": cigar}


def align(query, target, mode="NW", task="distance", k=-1, additionalEqualities=None):
    """Same contract as the reference's `edlib.align` (edlib.pyx:56-155)."""
    (q, t), eqs = _map_to_bytes([query, target], additionalEqualities)
    d = library().align(q, t, -1 if k is None else k, MODES.get(mode, 0), TASKS.get(task, 0), eqs)
    return _result(d, True)


def align_batch(queries, targets, mode="NW", task="distance", k=-1, additionalEqualities=None):
    """Batched `align`: `targets` is one sequence shared by all queries, or one per query (repeat the
    same object to share its upload).  Returns one dict per query."""
    queries = list(queries)
    shared = isinstance(targets, (bytes, bytearray, str))
    tlist = [targets] if shared else list(targets)
    distinct, index = [], {}
    for t in tlist:
        if id(t) not in index:
            index[id(t)] = len(distinct)
            distinct.append(t)
    mapped, eqs = _map_to_bytes(queries + distinct, additionalEqualities)
    qs, ts = mapped[:len(queries)], mapped[len(queries):]
    per_query = [ts[0]] * len(qs) if shared else [ts[index[id(t)]] for t in tlist]
    st, res = library().align_batch(qs, per_query, -1 if k is None else k, MODES.get(mode, 0), TASKS.get(task, 0), eqs)
    if st != EDLIB_STATUS_OK:
        raise Exception("There was an error. (" + library().lib.edlibB200LastError().decode() + ")")
    return [_result(d, True) for d in res]


align_many = align_batch  # the name SURVEY.md 8f proposes for the batched binding entry


def getNiceAlignment(alignResult, query, target, gapSymbol="-"):
    """Human-readable three-line view of an `align(..., task='path')` result (edlib.pyx:157-238):
    dict with 'query_aligned', 'matched_aligned' ('|' match, '.' mismatch, gap symbol for indels) and
    'target_aligned'."""
    if not isinstance(alignResult, dict):
        raise Exception("The object alignResult is expected to be a python dictionary. Please check the input alignResult.")
    if "locations" not in alignResult:
        raise Exception("The object alignResult is expected to contain a field 'locations'. Please check the input alignResult.")
    if "cigar" not in alignResult:
        raise Exception("The object alignResult is expected to contain a CIGAR string. Please check the input alignResult.")
    cigar = alignResult["cigar"]
    if not cigar:
        raise Exception("The object alignResult contains an empty CIGAR string. Users must run align() with task='path'. "
                        "Please check the input alignResult.")
    tpos = alignResult["locations"][0][0] or 0
    qpos = 0
    t_aln, m_aln, q_aln = [], [], []
    for count, op in _re.findall(r"(\d+)(\D)", cigar):
        n = int(count)
        if op in "=X":
            t_aln.append(target[tpos:tpos + n])
            q_aln.append(query[qpos:qpos + n])
            m_aln.append(("|" if op == "=" else ".") * n)
            tpos += n
            qpos += n
        elif op == "D":
            t_aln.append(target[tpos:tpos + n])
            q_aln.append(gapSymbol * n)
            m_aln.append(gapSymbol * n)
            tpos += n
        elif op == "I":
            t_aln.append(gapSymbol * n)
            q_aln.append(query[qpos:qpos + n])
            m_aln.append(gapSymbol * n)
            qpos += n
        else:
            raise Exception("The CIGAR string from alignResult contains a symbol not '=', 'X', 'D', 'I'. "
                            "Please check the validity of alignResult and alignResult.cigar")
    return {"query_aligned": "".join(q_aln), "matched_aligned": "".join(m_aln), "target_aligned": "".join(t_aln)}
