"""Flat read arrays — the host-side image of ``tc_reads_t`` (include/trueconsense_b200.h).

A :class:`ReadBatch` is what the BAM decoder (``bamio.read_bam``) and the synthetic generator
(``synth``) produce and what the GPU entry points consume.  Arrays are numpy views; they may sit
on C-owned memory (kept alive by ``_owner``), plain numpy memory, or pinned torch memory
(after :meth:`ReadBatch.pin`).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Any

import numpy as np

# CIGAR op codes (BAM): MIDNSHP=X
CIGAR_OPS = "MIDNSHP=X"
REF_CONSUMING = (0, 2, 3, 7, 8)
SEQ_CODES = "=ACMGRSVTWYHKDBN"


class TcReads(C.Structure):
    """ctypes mirror of ``tc_reads_t``."""

    _fields_ = [
        ("n_reads", C.c_int64),
        ("n_seq_words", C.c_int64),
        ("n_cigar_ops", C.c_int64),
        ("pos", C.c_void_p),
        ("flag", C.c_void_p),
        ("mapq", C.c_void_p),
        ("l_seq", C.c_void_p),
        ("seq_off", C.c_void_p),
        ("cigar_off", C.c_void_p),
        ("seq4", C.c_void_p),
        ("qual", C.c_void_p),
        ("cigar", C.c_void_p),
        ("qname_hash", C.c_void_p),
        ("mpos", C.c_void_p),
        ("isize", C.c_void_p),
        ("max_ref_span", C.c_int32),
        ("reserved", C.c_int32),
        ("cigar16", C.c_void_p),
        ("seq2", C.c_void_p),
        ("seq_exc_idx", C.c_void_p),
        ("seq_exc_val", C.c_void_p),
        ("n_seq_exc", C.c_int64),
    ]


_ARRAYS = ("pos", "flag", "mapq", "l_seq", "seq_off", "cigar_off", "seq4", "qual", "cigar",
           "qname_hash", "mpos", "isize")
_DTYPES = {
    "pos": np.int32, "flag": np.uint16, "mapq": np.uint8, "l_seq": np.int32, "seq_off": np.uint32,
    "cigar_off": np.uint32, "seq4": np.uint32, "qual": np.uint8, "cigar": np.uint32,
    "qname_hash": np.uint64, "mpos": np.int32, "isize": np.int32, "tid": np.int32, "mtid": np.int32,
}


@dataclass
class ReadBatch:
    pos: np.ndarray
    flag: np.ndarray
    mapq: np.ndarray
    l_seq: np.ndarray
    seq_off: np.ndarray      # [n+1] uint32, 32-bit words
    cigar_off: np.ndarray    # [n+1] uint32
    seq4: np.ndarray         # uint32 words holding BAM's nibble byte stream
    qual: np.ndarray         # uint8, 8 bytes per seq word
    cigar: np.ndarray        # uint32 len<<4|op
    qname_hash: np.ndarray | None = None
    mpos: np.ndarray | None = None
    isize: np.ndarray | None = None
    tid: np.ndarray | None = None
    mtid: np.ndarray | None = None
    ref_names: list[str] = field(default_factory=lambda: ["ref"])
    ref_lens: list[int] = field(default_factory=list)
    aligned_bases: int = -1
    max_ref_span: int = -1
    sorted: bool = True
    info: dict = field(default_factory=dict)
    _owner: Any = None       # keeps backing memory alive
    cigar16: np.ndarray | None = None    # the same operations in 16 bits (tc_reads_t.cigar16), see with_cigar16()
    seq2: np.ndarray | None = None       # two bits per base, one uint16 per seq4 word (tc_reads_t.seq2), see with_seq2()
    seq_exc_idx: np.ndarray | None = None    # ... and the words that hold anything else than A C G T
    seq_exc_val: np.ndarray | None = None

    # ------------------------------------------------------------------ basics
    @property
    def n_reads(self) -> int:
        return int(self.pos.shape[0])

    def __len__(self) -> int:
        return self.n_reads

    def validate(self) -> None:
        n = self.n_reads
        for name in _ARRAYS:
            a = getattr(self, name)
            if a is None:
                continue
            if a.dtype != _DTYPES[name]:
                raise TypeError(f"{name}: dtype {a.dtype}, expected {_DTYPES[name]}")
            if not a.flags["C_CONTIGUOUS"]:
                raise ValueError(f"{name} must be contiguous")
        if self.seq_off.shape[0] != n + 1 or self.cigar_off.shape[0] != n + 1:
            raise ValueError("offset arrays must have n_reads+1 entries")
        if int(self.seq_off[-1]) != self.seq4.shape[0] or int(self.cigar_off[-1]) != self.cigar.shape[0]:
            raise ValueError("offsets do not end at the buffer sizes")
        if self.qual.shape[0] != 8 * self.seq4.shape[0]:
            raise ValueError("qual must hold 8 bytes per seq word")

    def c_struct(self) -> TcReads:
        """``tc_reads_t`` pointing at this batch's host memory (keep ``self`` alive while used)."""
        s = TcReads()
        s.n_reads = self.n_reads
        s.n_seq_words = int(self.seq4.shape[0])
        s.n_cigar_ops = int(self.cigar.shape[0])
        for name in _ARRAYS:
            a = getattr(self, name)
            if name == "mpos" and a is not None:
                a = self._mpos_for_abi()
            setattr(s, name, None if a is None else a.ctypes.data)
        s.max_ref_span = max(int(self.max_ref_span), 0)      # 0 = unknown
        s.cigar16 = None if self.cigar16 is None else self.cigar16.ctypes.data
        if self.seq2 is not None:
            s.seq2 = self.seq2.ctypes.data
            s.n_seq_exc = int(self.seq_exc_idx.shape[0])
            s.seq_exc_idx = self.seq_exc_idx.ctypes.data if s.n_seq_exc else None
            s.seq_exc_val = self.seq_exc_val.ctypes.data if s.n_seq_exc else None
        return s

    def with_seq2(self, max_exception_fraction: float = 0.125) -> "ReadBatch":
        """This batch with the compact SEQ transport attached (``tc_reads_t.seq2``: two bits per base plus the list of words that
        hold anything else than A C G T — half the SEQ bytes over PCIe); the batch itself when more than
        ``max_exception_fraction`` of the words would be exceptions."""
        import copy

        from . import bamio

        if self.seq2 is not None or self.seq4.size == 0:
            return self
        lib = bamio.host_lib()
        seq2 = np.empty(self.seq4.shape[0], dtype=np.uint16)
        idx, val, n = C.POINTER(C.c_uint32)(), C.POINTER(C.c_uint32)(), C.c_int64(0)
        seq4 = np.ascontiguousarray(self.seq4)
        rc = lib.tc_seq2_pack(seq4.ctypes.data, seq4.shape[0], np.ascontiguousarray(self.seq_off).ctypes.data,
                              np.ascontiguousarray(self.l_seq).ctypes.data, self.n_reads, seq2.ctypes.data, C.byref(idx), C.byref(val),
                              C.byref(n), 0)
        if rc != 0:
            raise ValueError(f"tc_seq2_pack failed ({rc})")
        try:
            if n.value > max_exception_fraction * seq4.shape[0]:
                return self
            out = copy.copy(self)
            out.seq2 = seq2
            out.seq_exc_idx = np.ctypeslib.as_array(idx, shape=(n.value,)).copy() if n.value else np.empty(0, np.uint32)
            out.seq_exc_val = np.ctypeslib.as_array(val, shape=(n.value,)).copy() if n.value else np.empty(0, np.uint32)
            return out
        finally:
            lib.tc_host_free(idx)
            lib.tc_host_free(val)

    def with_cigar16(self) -> "ReadBatch":
        """This batch with the compact CIGAR transport array attached (``tc_reads_t.cigar16``: half the bytes over PCIe) when
        every operation is shorter than 4096; otherwise the batch itself."""
        import copy

        if self.cigar16 is not None or self.cigar.size == 0 or int(self.cigar.max()) >= (4096 << 4):
            return self
        out = copy.copy(self)
        out.cigar16 = self.cigar.astype(np.uint16)
        return out

    def _mpos_for_abi(self) -> np.ndarray:
        """tc_reads_t.mpos: PNEXT, -1 if unavailable, -2 when the mate maps to another reference (htslib's overlap
        handling never pairs such reads).  The batch itself keeps the BAM's PNEXT."""
        if self.tid is None or self.mtid is None:
            return self.mpos
        cached = getattr(self, "_mpos_abi", None)           # (computed once per batch: c_struct() is on the per-sample path)
        if cached is None or cached[0] is not self.mpos:
            other = (self.mtid >= 0) & (self.mtid != self.tid)
            arr = np.where(other, np.int32(-2), self.mpos).astype(np.int32) if other.any() else self.mpos
            cached = (self.mpos, arr)
            object.__setattr__(self, "_mpos_abi", cached)
        return cached[1]

    # ------------------------------------------------------------------ derived quantities
    def ref_spans(self) -> np.ndarray:
        """Reference span (sum of M,=,X,D,N lengths) of every read."""
        op = self.cigar & 15
        ln = (self.cigar >> 4).astype(np.int64)
        consuming = np.isin(op, REF_CONSUMING)
        contrib = np.where(consuming, ln, 0)
        csum = np.concatenate(([0], np.cumsum(contrib)))
        return (csum[self.cigar_off[1:].astype(np.int64)] - csum[self.cigar_off[:-1].astype(np.int64)]).astype(np.int64)

    def count_aligned_bases(self, flag_filter: int = 0x4) -> int:
        """Pileup entries these reads produce under a flag filter: the benchmark's unit of work
        (one iteration of TrueConsense/indexing.py:116)."""
        spans = self.ref_spans()
        keep = (self.flag & flag_filter) == 0
        return int(spans[keep].sum())

    def algorithmic_bytes(self, ref_len: int, with_qual: bool = False) -> int:
        """per read ceil(l_seq/2) + 4*n_cigar + 16 (+ l_seq QUAL bytes when a
        base-quality filter is applied); per sample 8*4*L count table."""
        l = self.l_seq.astype(np.int64)
        ncig = np.diff(self.cigar_off.astype(np.int64))
        b = int(((l + 1) // 2).sum() + 4 * ncig.sum() + 16 * self.n_reads)
        if with_qual:
            b += int(l.sum())
        return b + 8 * 4 * int(ref_len)

    # ------------------------------------------------------------------ re-shaping
    def slice(self, r0: int, r1: int) -> "ReadBatch":
        """Reads [r0, r1) as a batch of their own (big arrays are views, offsets rebased)."""
        r0 = max(0, int(r0)); r1 = min(self.n_reads, int(r1))
        if r1 < r0:
            r1 = r0
        s0, s1 = int(self.seq_off[r0]), int(self.seq_off[r1])
        c0, c1 = int(self.cigar_off[r0]), int(self.cigar_off[r1])

        def cut(a):
            return None if a is None else a[r0:r1]

        return ReadBatch(
            pos=self.pos[r0:r1], flag=self.flag[r0:r1], mapq=self.mapq[r0:r1], l_seq=self.l_seq[r0:r1],
            seq_off=(self.seq_off[r0:r1 + 1] - np.uint32(s0)).astype(np.uint32),
            cigar_off=(self.cigar_off[r0:r1 + 1] - np.uint32(c0)).astype(np.uint32),
            seq4=self.seq4[s0:s1], qual=self.qual[8 * s0:8 * s1], cigar=self.cigar[c0:c1],
            qname_hash=cut(self.qname_hash), mpos=cut(self.mpos), isize=cut(self.isize),
            tid=cut(self.tid), mtid=cut(self.mtid),
            ref_names=list(self.ref_names), ref_lens=list(self.ref_lens), sorted=self.sorted,
            max_ref_span=self.max_ref_span, _owner=self,
        )

    def pin(self) -> "ReadBatch":
        """Copy every array into pinned (page-locked) host memory so H2D copies run at full
        PCIe rate and asynchronously."""
        import torch

        keep = []

        def pinned(a):
            if a is None:
                return None
            nbytes = max(int(a.nbytes), 1)
            t = torch.empty(nbytes, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
            keep.append(t)
            v = t.numpy()[: a.nbytes].view(a.dtype)
            v[...] = a.reshape(-1)
            return v

        out = ReadBatch(
            **{name: pinned(getattr(self, name)) for name in _ARRAYS},
            tid=self.tid, mtid=self.mtid, ref_names=list(self.ref_names), ref_lens=list(self.ref_lens),
            aligned_bases=self.aligned_bases, max_ref_span=self.max_ref_span, sorted=self.sorted,
            info=dict(self.info),
        )
        out.cigar16 = pinned(self.cigar16)
        out.seq2, out.seq_exc_idx, out.seq_exc_val = pinned(self.seq2), pinned(self.seq_exc_idx), pinned(self.seq_exc_val)
        out._owner = keep
        return out

    # ------------------------------------------------------------------ construction from python objects
    @staticmethod
    def from_records(records: list[dict]) -> "ReadBatch":
        """Build a batch from a list of ``{"pos", "cigar": "10M2I5M" | [(op,len)], "seq": "ACGT..",
        "qual": [..] | int, "flag", "mapq", "qname", "mpos", "isize"}`` (tests and small fixtures)."""
        import re

        n = len(records)
        pos = np.zeros(n, np.int32); flag = np.zeros(n, np.uint16); mapq = np.zeros(n, np.uint8)
        l_seq = np.zeros(n, np.int32); seq_off = np.zeros(n + 1, np.uint32); cig_off = np.zeros(n + 1, np.uint32)
        qh = np.zeros(n, np.uint64); mpos = np.full(n, -1, np.int32); isize = np.zeros(n, np.int32)
        seq_bytes = bytearray(); quals = bytearray(); cig: list[int] = []
        code = {c: i for i, c in enumerate(SEQ_CODES)}
        for i, r in enumerate(records):
            pos[i] = r["pos"]; flag[i] = r.get("flag", 0); mapq[i] = r.get("mapq", 60)
            c = r.get("cigar", "")
            if isinstance(c, str):
                ops = [(CIGAR_OPS.index(o), int(l)) for l, o in re.findall(r"(\d+)([MIDNSHP=X])", c)] if c != "*" else []
            else:
                ops = [(CIGAR_OPS.index(o) if isinstance(o, str) else int(o), int(l)) for o, l in c]
            for o, l in ops:
                cig.append((l << 4) | o)
            cig_off[i + 1] = len(cig)
            s = r.get("seq", "")
            if s == "*":
                s = ""
            l_seq[i] = len(s)
            nb = bytearray((len(s) + 1) // 2)
            for q, ch in enumerate(s):
                v = code[ch.upper()]
                nb[q >> 1] |= v << 4 if (q & 1) == 0 else v
            words = (len(s) + 7) // 8
            seq_bytes += nb + bytes(4 * words - len(nb))
            q = r.get("qual", 30)
            qb = bytearray([int(q)] * len(s)) if isinstance(q, int) else bytearray(int(x) for x in q)
            if len(qb) != len(s):
                raise ValueError("qual length != seq length")
            quals += qb + bytes(8 * words - len(qb))
            seq_off[i + 1] = seq_off[i] + words
            name = r.get("qname", f"r{i}")
            qh[i] = _qname_hash(name)
            mpos[i] = r.get("mpos", -1); isize[i] = r.get("isize", 0)
        b = ReadBatch(
            pos=pos, flag=flag, mapq=mapq, l_seq=l_seq, seq_off=seq_off, cigar_off=cig_off,
            seq4=np.frombuffer(bytes(seq_bytes), dtype=np.uint32).copy() if seq_bytes else np.zeros(0, np.uint32),
            qual=np.frombuffer(bytes(quals), dtype=np.uint8).copy() if quals else np.zeros(0, np.uint8),
            cigar=np.array(cig, dtype=np.uint32), qname_hash=qh, mpos=mpos, isize=isize,
            tid=np.zeros(n, np.int32), mtid=np.where((flag & 1) != 0, 0, -1).astype(np.int32),
        )
        b.sorted = bool(np.all(np.diff(pos.astype(np.int64)) >= 0))
        b.validate()
        return b

    def seq_string(self, i: int) -> str:
        """Decode read i's SEQ (debugging / tests)."""
        n = int(self.l_seq[i])
        raw = self.seq4[int(self.seq_off[i]):int(self.seq_off[i + 1])].view(np.uint8)
        out = []
        for q in range(n):
            b = int(raw[q >> 1])
            out.append(SEQ_CODES[(b >> 4) if (q & 1) == 0 else (b & 15)])
        return "".join(out)

    def cigar_string(self, i: int) -> str:
        ops = self.cigar[int(self.cigar_off[i]):int(self.cigar_off[i + 1])]
        return "".join(f"{int(c) >> 4}{CIGAR_OPS[int(c) & 15]}" for c in ops) or "*"


def _qname_hash(name: str) -> int:
    """low 32 bits: khash X31 string hash (what htslib keys its overlap table with);
    high 32 bits: folded FNV-1a — same as csrc/host/bamio.c."""
    bs = name.encode()
    h = bs[0] if bs else 0
    for ch in bs[1:]:
        h = ((h << 5) - h + ch) & 0xFFFFFFFF
    f = 1469598103934665603
    for ch in bs:
        f ^= ch
        f = (f * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    hi = ((f >> 32) ^ f) & 0xFFFFFFFF
    return (hi << 32) | h
