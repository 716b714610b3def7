"""Stand-in ``pysam`` / ``AminoExtract`` / ``Bio`` modules so the UNMODIFIED reference package at
/root/reference can be imported and run in this container (TEST INFRASTRUCTURE ONLY — see
oracle/__init__.py).  Used by oracle/make_golden.py to generate tests/golden/*, and by tests that
run only where /root/reference exists (never on the GPU box).

pysam 0.23.3, aminoextract 0.4.1 and biopython 1.85 (the reference's pins, pyproject.toml:26-32)
are not installed in this image.  The ``pysam`` stand-in answers the reference's calls
(indexing.py:19,96-100,139; Events.py:63-67) from oracle/pileup_oracle.c — the restated htslib
engine — so that the reference's own classifier, mode/regex logic, consensus walk, GFF correction
and writers all run as written on top of it.
"""
from __future__ import annotations

import os
import sys
import types

import pandas as pd

from . import bam_py, pileup

REFERENCE_ROOT = "/root/reference"


# ------------------------------------------------------------------------------ pysam
class _PileupColumn:
    def __init__(self, pos0: int, strings):
        self.pos = pos0
        self.reference_pos = pos0
        self._strings = strings
        self.nsegments = len(strings) if strings != "" else 0

    def get_query_sequences(self, mark_matches=False, mark_ends=False, add_indels=False):
        if not add_indels or mark_matches or mark_ends:
            raise NotImplementedError("the stand-in only implements add_indels=True")
        return self._strings


class AlignmentFile:
    """pysam.AlignmentFile look-alike over an oracle-decoded BAM."""

    def __init__(self, path, mode="rb", **kw):
        self.filename = path
        self._b = bam_py.read_bam(path)
        self.references = tuple(self._b.ref_names)
        self.lengths = tuple(self._b.ref_lens)
        self.nreferences = len(self.references)
        self.pileup_calls: list[tuple] = []

    def pileup(self, contig=None, start=None, stop=None, region=None, reference=None, end=None, truncate=False,
               stepper="samtools", max_depth=8000, min_base_quality=13, ignore_overlaps=True,
               flag_filter=0x4 | 0x100 | 0x200 | 0x400, ignore_orphans=True, min_mapping_quality=0, **kw):
        contig = contig if contig is not None else reference
        stop = stop if stop is not None else end
        self.pileup_calls.append((contig, start, stop, truncate))
        if stepper == "nofilter":
            params = dict(flag_filter=0x4, min_mapq=0, ignore_orphans=0)
        elif stepper == "samtools":
            params = dict(flag_filter=int(flag_filter) | 0x4, min_mapq=int(min_mapping_quality),
                          ignore_orphans=int(bool(ignore_orphans)))
        else:
            raise NotImplementedError(stepper)
        params.update(min_base_quality=int(min_base_quality), max_depth=int(max_depth))
        if any(t != 0 for t in self._b.tid):
            raise NotImplementedError("stand-in handles single-contig BAMs")
        reg = None
        if contig is not None:
            if contig != self.references[0]:
                raise KeyError(contig)
            if not truncate:
                raise NotImplementedError("the reference only issues truncate=True region pileups (Events.py:66)")
            reg = (0 if start is None else int(start), self.lengths[0] if stop is None else int(stop))
        cols = pileup.pileup_columns(self._b, region=reg, **params)
        for pos0, strings in cols:
            yield _PileupColumn(pos0, strings)

    def close(self):
        pass


class FastaFile:
    def __init__(self, path):
        recs = bam_py.fasta_records(path)
        self.references = tuple(n for n, _ in recs)
        self.lengths = tuple(len(s) for _, s in recs)
        self.nreferences = len(recs)

    def close(self):
        pass


# ------------------------------------------------------------------------------ AminoExtract
GFF_COLUMNS = ["seqid", "source", "type", "start", "end", "score", "strand", "phase", "attributes"]


class GFFColumns:
    @classmethod
    def get_names(cls):
        return list(GFF_COLUMNS)


class _GFFHeader:
    def __init__(self, raw_text: str):
        self.raw_text = raw_text


class GFFDataFrame:
    def __init__(self, file=None, logger=None, **kw):
        header_lines, rows = [], []
        with open(file) as fh:
            for line in fh:
                if line.startswith("#"):
                    if not rows:
                        header_lines.append(line)
                    continue
                if not line.strip():
                    continue
                f = line.rstrip("\n").split("\t")
                f += [""] * (9 - len(f))
                rows.append(f[:9])
        self.header = _GFFHeader("".join(header_lines))
        df = pd.DataFrame(rows, columns=GFF_COLUMNS)
        df["start"] = df["start"].astype("int64")
        df["end"] = df["end"].astype("int64")
        self.df = df


class SequenceReader:
    def __init__(self, logger=None, verbose=False):
        self.logger = logger

    def read_gff(self, file):
        return GFFDataFrame(file=file, logger=self.logger)


# ------------------------------------------------------------------------------ Bio
class _SeqRecord:
    def __init__(self, name, seq):
        self.id = name
        self.name = name
        self.seq = seq


def _seqio_parse(path, fmt):
    if fmt != "fasta":
        raise NotImplementedError(fmt)
    for n, s in bam_py.fasta_records(path):
        yield _SeqRecord(n, s)


# ------------------------------------------------------------------------------ installation
def install_stubs() -> None:
    if "pysam" not in sys.modules:
        m = types.ModuleType("pysam")
        m.AlignmentFile = AlignmentFile
        m.FastaFile = FastaFile
        m.__stand_in__ = True
        sys.modules["pysam"] = m
    if "AminoExtract" not in sys.modules:
        m = types.ModuleType("AminoExtract")
        m.SequenceReader = SequenceReader
        m.GFFDataFrame = GFFDataFrame
        g = types.ModuleType("AminoExtract.gff_data")
        g.GFFColumns = GFFColumns
        m.gff_data = g
        m.__stand_in__ = True
        sys.modules["AminoExtract"] = m
        sys.modules["AminoExtract.gff_data"] = g
    if "Bio" not in sys.modules:
        m = types.ModuleType("Bio")
        s = types.ModuleType("Bio.SeqIO")
        s.parse = _seqio_parse
        m.SeqIO = s
        m.__stand_in__ = True
        sys.modules["Bio"] = m
        sys.modules["Bio.SeqIO"] = s


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "TrueConsense"))


def load_reference():
    """Import the unmodified reference package (read-only tree: no bytecode is written)."""
    if not reference_available():
        raise FileNotFoundError(f"{REFERENCE_ROOT} is not present (it never is on the GPU box)")
    sys.dont_write_bytecode = True
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib

    pkg = importlib.import_module("TrueConsense")
    mods = {}
    for name in ("Ambig", "Coverage", "Events", "ORFs", "Sequences", "indexing", "Outputs", "TrueConsense", "func"):
        mods[name] = importlib.import_module(f"TrueConsense.{name}")
    return types.SimpleNamespace(pkg=pkg, **mods)


class FakeBam:
    """Duck-typed ``bam`` for Events.ExtractInserts fed by hand-written column strings
    (the SURVEY.md §4.3 recipe): ``columns`` maps 0-based position -> list of str (or "")."""

    def __init__(self, columns: dict | None = None, name: str = "ref"):
        self.references = [name]
        self.columns = columns or {}
        self.calls: list[tuple] = []

    def pileup(self, rname, start, end, truncate=False, **kw):
        self.calls.append((rname, start, end, truncate))
        for p in range(start, end):
            if p in self.columns:
                yield _PileupColumn(p, self.columns[p])
