"""Pileup index building with the reference's API (TrueConsense/indexing.py).

``BuildIndex(bamfile, ref)`` returns the same DataFrame (seven int64 columns
``coverage,A,T,C,G,X,I``, int64 index 1..len(ref[0]), ``index.name is None``) the reference
builds from a pysam pileup plus a per-string Python classifier (indexing.py:75-154).  Here the
BAM is inflated on the host (csrc/host/bamio.c), its records are parsed on the GPU into flat arrays
(``tc_bam_records_to_reads``) and the whole pileup is one pass of the CUDA kernels behind ``tc_pileup_counts``.
"""
from __future__ import annotations

import threading

import numpy as np
import pandas as pd

from . import bamio, gpu
from .reads import ReadBatch

COLUMNS = ["coverage", "A", "T", "C", "G", "X", "I"]


class BamHandle:
    """What ``Readbam`` returns: a BAM on its way to the device plus the attributes of ``pysam.AlignmentFile`` the reference
    touches (``references``; Events.py:63).  The file travels to the device as it is: BGZF members inflated, record offsets
    found and records parsed on the GPU (``gpu.Context.bam_file_to_device``) straight into the context's read buffers; ``reads`` (a host-side decode into numpy arrays) is
    only made when somebody asks for it."""

    def __init__(self, filename: str, reads: ReadBatch | None = None):
        self.filename = filename
        self._reads = reads
        self._payload = None
        self._hdr = None
        self._dev = None
        self._lock = threading.Lock()
        self._insert_cache: dict = {}

    @property
    def reads(self) -> ReadBatch:
        with self._lock:
            if self._reads is None:
                self._reads = bamio.read_bam(self.filename)
            return self._reads

    def _header(self):
        """(reference names, reference lengths) from whichever decode exists."""
        with self._lock:
            if self._reads is not None:
                return self._reads.ref_names, self._reads.ref_lens
            if self._payload is not None:
                return self._payload.ref_names, self._payload.ref_lens
            if self._hdr is None:
                self._hdr = bamio.read_bam_header(self.filename)       # only the members holding the header are inflated
            return self._hdr

    @property
    def references(self):
        return tuple(self._header()[0])

    @property
    def lengths(self):
        return tuple(self._header()[1])

    @property
    def ref_len(self) -> int:
        return int(self._header()[1][0])

    def contig0(self) -> ReadBatch:
        """Host-side reads placed on the first reference (the only one the reference implementation is
        meaningful for: indexing.py:98,139 and Events.py:63 use lengths[0] / references[0])."""
        b = self.reads
        if b.tid is not None and b.n_reads and np.any(b.tid != 0):
            raise ValueError("multi-contig BAM: TrueConsense indexes positions of a single reference "
                             "(its DataFrame index would hold duplicate positions)")
        return b

    def device_reads(self, with_host_qual: bool = False):
        """The reads in device memory, made once per handle and reused by later passes while the context has not staged
        anything else.  From a file: inflated on the host, parsed on the GPU (QUAL and the mate fields included).  From a
        host batch handed to the constructor: uploaded without QUAL / mate fields, which ExtractInserts then stages for the
        reads over its candidate columns only."""
        ctx = gpu.default_context()
        with self._lock:
            d = self._dev
            if d is None or d.ctx is not ctx or d.generation != ctx._generation:
                if self._reads is not None:
                    b = self._reads
                    if b.tid is not None and b.n_reads and np.any(b.tid != 0):
                        raise ValueError("multi-contig BAM: TrueConsense indexes positions of a single reference")
                    if not b.sorted:
                        raise ValueError("Unsorted input. Pileup aborts")
                    d = ctx.upload(b, with_qual=False)
                else:
                    try:
                        d = ctx.bam_file_to_device(self.filename)      # inflate, record index and parse on the GPU
                        self._hdr = (d.ref_names, d.ref_lens)
                    except gpu.TcError:
                        # a file the device path refuses: the host reader names the broken member / record (or, should it
                        # read the file after all, feeds the device parse)
                        self._payload = bamio.read_bam_payload(self.filename)
                        d = ctx.bam_to_device(self._payload)
                        self._payload.release()     # the device holds the arrays now; the header stays
                    if d.stats.multi_contig:
                        raise ValueError("multi-contig BAM: TrueConsense indexes positions of a single reference "
                                         "(its DataFrame index would hold duplicate positions)")
                    if d.stats.unsorted:
                        raise ValueError("Unsorted input. Pileup aborts")
                self._dev = d
            return d.with_host_qual() if with_host_qual else d

    def pileup(self, *args, **kwargs):
        raise NotImplementedError("column iteration is not part of this implementation; "
                                  "use Events.ExtractInserts / indexing.BuildIndex")

    def close(self):
        pass


def Readbam(f):
    """indexing.py:6-19.  A handle passed in is returned as is, so one decoded copy of the BAM can
    serve BuildIndex, ListInserts and WriteOutputs."""
    if isinstance(f, BamHandle):
        return f
    return BamHandle(f)


class _GffHeader:
    def __init__(self, raw_text: str):
        self.raw_text = raw_text


class GffIndex:
    """Stand-in for AminoExtract's GFFDataFrame (not installed here): ``.df`` with the nine GFF3
    columns (start/end int64) and ``.header.raw_text`` — all the reference reads
    (TrueConsense.py:238-241, Outputs.py:66)."""

    GFF_COLUMNS = ["seqid", "source", "type", "start", "end", "score", "strand", "phase", "attributes"]

    def __init__(self, file: str):
        header, rows = [], []
        with open(file) as fh:
            for line in fh:
                if line.startswith("#"):
                    if not rows:
                        header.append(line)
                    continue
                if not line.strip():
                    continue
                f = line.rstrip("\n").split("\t")
                f += [""] * (9 - len(f))
                rows.append(f[:9])
        self.header = _GffHeader("".join(header))
        df = pd.DataFrame(rows, columns=self.GFF_COLUMNS)
        df["start"] = df["start"].astype("int64")
        df["end"] = df["end"].astype("int64")
        self.df = df


def Gffindex(file: str) -> GffIndex:
    """indexing.py:22-36."""
    return GffIndex(file)


def read_override_index(f):
    """indexing.py:39-52."""
    return pd.read_csv(f, sep=",", compression="gzip", index_col=0)


def Override_index_positions(index, override_data):
    """indexing.py:55-72."""
    index.loc[override_data.index, :] = override_data[:]
    return index


def frame_from_counts(counts: np.ndarray) -> pd.DataFrame:
    """int32[8][L] count table -> the reference's index frame."""
    L = counts.shape[1]
    df = pd.DataFrame({c: counts[r].astype(np.int64) for r, c in enumerate(COLUMNS)},
                      index=pd.Index(np.arange(1, L + 1, dtype=np.int64)))
    df.index.name = None
    return df


def BuildIndex(bamfile, ref):
    """indexing.py:75-154."""
    handle = bamfile if isinstance(bamfile, BamHandle) else BamHandle(bamfile)
    _, lens = bamio.read_fasta_lengths(ref)
    if not lens:
        raise ValueError(f"no sequences in {ref}")
    ref_length = int(lens[0])
    counts = gpu.default_context().pileup_counts(handle.device_reads(), ref_length)
    return frame_from_counts(counts)
