"""ctypes front of oracle/pileup_oracle.c (TEST INFRASTRUCTURE ONLY — see oracle/__init__.py).

Works on the same flat arrays as the product (any object with the ``ReadBatch`` attributes),
but shares no code with it: the C file is a streaming, column-by-column restatement of
htslib's pileup engine + pysam's string formatting + the reference's classifier
(TrueConsense/indexing.py:102-132).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_SRC = os.path.join(_HERE, "pileup_oracle.c")


class OReads(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("n_seq_words", C.c_int64), ("n_cigar_ops", C.c_int64)] + [
        (n, C.c_void_p) for n in ("pos", "flag", "mapq", "l_seq", "seq_off", "cigar_off", "seq4", "qual", "cigar",
                                  "qname_hash", "mpos", "isize")]


class OParams(C.Structure):
    _fields_ = [("flag_filter", C.c_uint32), ("min_mapq", C.c_int32), ("min_base_quality", C.c_int32),
                ("ignore_orphans", C.c_int32), ("max_depth", C.c_int64), ("kernel", C.c_int32), ("reserved", C.c_int32)]


# pysam arguments at the reference's two call sites
BUILDINDEX = dict(flag_filter=0x4, min_mapq=0, min_base_quality=0, ignore_orphans=0, max_depth=10_000_000)   # indexing.py:100
EXTRACTINSERTS = dict(flag_filter=0x4 | 0x100 | 0x200 | 0x400, min_mapq=0, min_base_quality=13, ignore_orphans=1,
                      max_depth=8000)                                                                         # Events.py:66

_lib = None


def build(force: bool = False) -> str:
    os.makedirs(os.path.dirname(_LIB_PATH), exist_ok=True)
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(_SRC):
        subprocess.run(["gcc", "-O2", "-g", "-fPIC", "-shared", "-fopenmp", "-Wall", "-o", _LIB_PATH, _SRC, "-lm"],
                       check=True)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        l = C.CDLL(_LIB_PATH)
        l.oracle_pileup_counts.argtypes = [C.POINTER(OReads), C.c_int32, C.POINTER(OParams), C.c_void_p, C.c_int]
        l.oracle_pileup_counts.restype = C.c_int
        l.oracle_pileup_strings.argtypes = [C.POINTER(OReads), C.POINTER(OParams), C.c_int32, C.c_int32,
                                            C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.POINTER(C.c_void_p),
                                            C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
        l.oracle_pileup_strings.restype = C.c_int
        l.oracle_free.argtypes = [C.c_void_p]
        l.oracle_free.restype = None
        _lib = l
    return _lib


def _reads_struct(b) -> OReads:
    s = OReads()
    s.n_reads = len(b.pos)
    s.n_seq_words = len(b.seq4)
    s.n_cigar_ops = len(b.cigar)
    keep = []
    for n in ("pos", "flag", "mapq", "l_seq", "seq_off", "cigar_off", "seq4", "qual", "cigar", "qname_hash", "mpos", "isize"):
        a = getattr(b, n, None)
        if n == "mpos" and a is not None and getattr(b, "tid", None) is not None and getattr(b, "mtid", None) is not None:
            # the overlap table never pairs a read whose mate maps to another reference (htslib sam.c overlap_push: mtid != tid)
            other = (b.mtid >= 0) & (b.mtid != b.tid)
            if other.any():
                a = np.where(other, np.int32(-2), a).astype(np.int32)
                keep.append(a)
        setattr(s, n, None if a is None else a.ctypes.data)
    s._keep = keep
    return s


def _params(**kw) -> OParams:
    p = OParams()
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def pileup_counts(batch, ref_len: int, threads: int = 1, **params) -> np.ndarray:
    """int32[8][ref_len] count table (rows coverage,A,T,C,G,X,I,pad) as BuildIndex would build it."""
    kw = dict(BUILDINDEX)
    kw.update(params)
    out = np.zeros((8, ref_len), dtype=np.int32)
    rs = _reads_struct(batch)
    rc = lib().oracle_pileup_counts(C.byref(rs), ref_len, C.byref(_params(**kw)), out.ctypes.data, threads)
    if rc == -3:
        raise ValueError("Unsorted input. Pileup aborts")
    if rc != 0:
        raise RuntimeError(f"oracle_pileup_counts failed ({rc})")
    return out


def pileup_columns(batch, region: tuple[int, int] | None = None, **params):
    """List of ``(pos0, strings)`` for every yielded column; ``strings`` is what pysam's
    ``get_query_sequences(add_indels=True)`` returns: a list of str, or ``""`` when every entry
    failed the base-quality test."""
    kw = dict(BUILDINDEX)
    kw.update(params)
    rs = _reads_struct(batch)
    buf = C.c_void_p(); blen = C.c_int64(); cpos = C.c_void_p(); coff = C.c_void_p(); cn = C.c_void_p(); ncol = C.c_int64()
    s, e = (-1, -1) if region is None else region
    rc = lib().oracle_pileup_strings(C.byref(rs), C.byref(_params(**kw)), s, e, C.byref(buf), C.byref(blen),
                                     C.byref(cpos), C.byref(coff), C.byref(cn), C.byref(ncol))
    try:
        if rc == -3:
            raise ValueError("Unsorted input. Pileup aborts")
        if rc != 0:
            raise RuntimeError(f"oracle_pileup_strings failed ({rc})")
        n = ncol.value
        raw = C.string_at(buf, blen.value).decode("ascii")
        pos = np.ctypeslib.as_array(C.cast(cpos, C.POINTER(C.c_int32)), shape=(max(n, 1),))[:n].copy()
        off = np.ctypeslib.as_array(C.cast(coff, C.POINTER(C.c_int64)), shape=(n + 1,)).copy()
        nent = np.ctypeslib.as_array(C.cast(cn, C.POINTER(C.c_int32)), shape=(max(n, 1),))[:n].copy()
        cols = []
        for j in range(n):
            cols.append((int(pos[j]), raw[off[j]:off[j + 1]].split(":") if nent[j] > 0 else ""))
        return cols
    finally:
        for p in (buf, cpos, coff, cn):
            lib().oracle_free(p)
