"""edlib_b200 -- B200-native batched edit-distance engine behind the edlib C ABI.

Python host mirror of the reference's binding (bindings/python/edlib.pyx:56-155): `align()`
takes the same keyword arguments and returns the same dict; `align_batch()` is the batched
form over `edlibAlignBatch`.  All computation happens in the CUDA library
edlib_b200/lib/libedlib_b200.so; importing this package without it (or without a GPU) works,
calling into it raises.
"""
from ._ffi import (EDLIB_CIGAR_EXTENDED, EDLIB_CIGAR_STANDARD, EDLIB_STATUS_OK, MODES, TASKS, EdlibLib,
                   product_path)

_lib = None


def library():
    """The loaded product library (raises OSError if it has not been built)."""
    global _lib
    if _lib is None:
        _lib = EdlibLib(product_path(), prefix="edlib", has_batch=True)
    return _lib


def _to_bytes(x):
    return x.encode("utf-8") if isinstance(x, str) else bytes(x)


def _result(d, want_cigar):
    if d["status"] != EDLIB_STATUS_OK:
        raise RuntimeError("edlib_b200: device path failed: " + library().lib.edlibB200LastError().decode()
                           if hasattr(library().lib, "edlibB200LastError") else "edlib_b200: device path failed")
    locations = None
    if d["endLocations"] is not None:
        starts = d["startLocations"] or [None] * d["numLocations"]
        locations = list(zip(starts, d["endLocations"]))
    cigar = library().cigar(d["alignment"], EDLIB_CIGAR_EXTENDED) if (want_cigar and d["alignment"] is not None) else None
    return {"editDistance": d["editDistance"], "alphabetLength": d["alphabetLength"],
            "locations": locations, "cigar": cigar}


def align(query, target, mode="NW", task="distance", k=-1, additionalEqualities=None):
    """Same contract as the reference's `edlib.align` for byte/str inputs (edlib.pyx:56-155)."""
    lib = library()
    lib.lib.edlibB200LastError.restype = __import__("ctypes").c_char_p
    eqs = [(_to_bytes(a), _to_bytes(b)) for a, b in additionalEqualities] if additionalEqualities else None
    d = lib.align(_to_bytes(query), _to_bytes(target), k, MODES[mode], TASKS[task], eqs)
    return _result(d, task == "path")


def align_batch(queries, targets, mode="NW", task="distance", k=-1, additionalEqualities=None):
    """Batched `align`: `targets` may be one sequence (shared by all queries) or one per query."""
    lib = library()
    lib.lib.edlibB200LastError.restype = __import__("ctypes").c_char_p
    qs = [_to_bytes(q) for q in queries]
    if isinstance(targets, (bytes, str, bytearray)):
        t = _to_bytes(targets)
        ts = [t] * len(qs)
    else:
        cache = {}
        ts = [cache.setdefault(id(t), _to_bytes(t)) for t in targets]
    eqs = [(_to_bytes(a), _to_bytes(b)) for a, b in additionalEqualities] if additionalEqualities else None
    st, res = lib.align_batch(qs, ts, k, MODES[mode], TASKS[task], eqs)
    if st != EDLIB_STATUS_OK:
        raise RuntimeError("edlib_b200: device path failed: " + lib.lib.edlibB200LastError().decode())
    return [_result(d, task == "path") for d in res]
