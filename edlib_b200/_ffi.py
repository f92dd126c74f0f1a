"""ctypes view of the edlib C ABI (include/edlib.h).

`EdlibLib` binds any shared object that exports the ABI under a symbol prefix; the package
itself only ever loads the product library edlib_b200/lib/libedlib_b200.so (prefix "edlib").
The test-suite re-uses the class to bind its checkers; that wiring lives in tests/, not here.
"""
import ctypes as C
import os

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

EDLIB_STATUS_OK, EDLIB_STATUS_ERROR = 0, 1
EDLIB_MODE_NW, EDLIB_MODE_SHW, EDLIB_MODE_HW = 0, 1, 2
EDLIB_TASK_DISTANCE, EDLIB_TASK_LOC, EDLIB_TASK_PATH = 0, 1, 2
EDLIB_CIGAR_STANDARD, EDLIB_CIGAR_EXTENDED = 0, 1
MODES = {"NW": 0, "SHW": 1, "HW": 2}
TASKS = {"distance": 0, "locations": 1, "path": 2}


class EqualityPair(C.Structure):          # include/edlib.h EdlibEqualityPair (2 bytes)
    _fields_ = [("first", C.c_char), ("second", C.c_char)]


class AlignConfig(C.Structure):           # include/edlib.h EdlibAlignConfig (32 bytes)
    _fields_ = [("k", C.c_int), ("mode", C.c_int), ("task", C.c_int),
                ("additionalEqualities", C.POINTER(EqualityPair)),
                ("additionalEqualitiesLength", C.c_int)]


class AlignResult(C.Structure):           # include/edlib.h EdlibAlignResult (48 bytes)
    _fields_ = [("status", C.c_int), ("editDistance", C.c_int),
                ("endLocations", C.POINTER(C.c_int)), ("startLocations", C.POINTER(C.c_int)),
                ("numLocations", C.c_int), ("alignment", C.POINTER(C.c_ubyte)),
                ("alignmentLength", C.c_int), ("alphabetLength", C.c_int)]


assert C.sizeof(EqualityPair) == 2 and C.sizeof(AlignConfig) == 32 and C.sizeof(AlignResult) == 48


def make_config(k=-1, mode=EDLIB_MODE_NW, task=EDLIB_TASK_DISTANCE, equalities=None):
    """Returns (config, keepalive) -- keepalive owns the equality array."""
    cfg = AlignConfig()
    cfg.k, cfg.mode, cfg.task = int(k), int(mode), int(task)
    keep = None
    if equalities:
        keep = (EqualityPair * len(equalities))()
        for i, (a, b) in enumerate(equalities):
            keep[i].first = a if isinstance(a, bytes) else bytes([a])
            keep[i].second = b if isinstance(b, bytes) else bytes([b])
        cfg.additionalEqualities = C.cast(keep, C.POINTER(EqualityPair))
        cfg.additionalEqualitiesLength = len(equalities)
    else:
        cfg.additionalEqualities = None
        cfg.additionalEqualitiesLength = 0
    return cfg, keep


def result_to_dict(r):
    """Copies every field of an AlignResult into plain Python (None for NULL arrays)."""
    d = {"status": r.status, "editDistance": r.editDistance, "numLocations": r.numLocations,
         "alignmentLength": r.alignmentLength, "alphabetLength": r.alphabetLength}
    if r.status != EDLIB_STATUS_OK:
        return {"status": r.status}
    d["endLocations"] = [r.endLocations[i] for i in range(r.numLocations)] if r.endLocations else None
    d["startLocations"] = [r.startLocations[i] for i in range(r.numLocations)] if r.startLocations else None
    d["alignment"] = bytes(bytearray(r.alignment[i] for i in range(r.alignmentLength))) if r.alignment else None
    return d


class EdlibLib:
    """One loaded implementation of the edlib ABI; `prefix` is the exported symbol prefix."""

    def __init__(self, path, prefix="edlib", has_batch=False):
        self.path = path
        self.lib = C.CDLL(path)
        self._align = getattr(self.lib, prefix + "Align")
        self._align.restype = AlignResult
        self._align.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, AlignConfig]
        self._free = getattr(self.lib, prefix + "FreeAlignResult")
        self._free.restype = None
        self._free.argtypes = [AlignResult]
        self._cigar = getattr(self.lib, prefix + "AlignmentToCigar")
        self._cigar.restype = C.c_void_p
        self._cigar.argtypes = [C.POINTER(C.c_ubyte), C.c_int, C.c_int]
        self._libc = C.CDLL(None)
        self._libc.free.argtypes = [C.c_void_p]
        self._batch = None
        if has_batch:
            self._batch = self.lib.edlibAlignBatch
            self._batch.restype = C.c_int
            self._batch.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_int),
                                    C.POINTER(C.c_char_p), C.POINTER(C.c_int),
                                    C.c_int, AlignConfig, C.POINTER(AlignResult)]

    def align_raw(self, q, t, cfg):
        return self._align(q, len(q), t, len(t), cfg)

    def free(self, r):
        self._free(r)

    def align(self, q, t, k=-1, mode=EDLIB_MODE_NW, task=EDLIB_TASK_DISTANCE, equalities=None):
        cfg, keep = make_config(k, mode, task, equalities)
        r = self.align_raw(q, t, cfg)
        d = result_to_dict(r)
        if r.status == EDLIB_STATUS_OK:
            self.free(r)
        del keep
        return d

    def cigar(self, alignment, fmt=EDLIB_CIGAR_EXTENDED):
        n = len(alignment)
        buf = (C.c_ubyte * max(n, 1))(*alignment)
        p = self._cigar(buf, n, fmt)
        if not p:
            return None
        s = C.string_at(p).decode("ascii")
        self._libc.free(p)
        return s

    def align_batch(self, queries, targets, k=-1, mode=EDLIB_MODE_NW, task=EDLIB_TASK_DISTANCE,
                    equalities=None):
        """queries/targets: lists of bytes (targets may repeat the SAME bytes object to share it).
        Returns (status, [dict])."""
        assert self._batch is not None
        n = len(queries)
        cfg, keep = make_config(k, mode, task, equalities)
        qptr = (C.c_char_p * n)(*queries)
        qlen = (C.c_int * n)(*[len(q) for q in queries])
        # identical bytes objects must map to identical pointers: build one buffer per distinct object
        bufs = {}
        tptr_vals, tlen_vals = [], []
        for t in targets:
            key = id(t)
            if key not in bufs:
                bufs[key] = C.create_string_buffer(t, len(t)) if len(t) else C.create_string_buffer(1)
            tptr_vals.append(C.cast(bufs[key], C.c_char_p))
            tlen_vals.append(len(t))
        tptr = (C.c_char_p * n)(*tptr_vals)
        tlen = (C.c_int * n)(*tlen_vals)
        res = (AlignResult * n)()
        st = self._batch(qptr, qlen, tptr, tlen, n, cfg, res)
        out = []
        for i in range(n):
            out.append(result_to_dict(res[i]))
            if res[i].status == EDLIB_STATUS_OK:
                self.free(res[i])
        del keep, bufs
        return st, out


def product_path():
    return os.path.join(REPO, "edlib_b200", "lib", "libedlib_b200.so")
