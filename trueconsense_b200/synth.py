"""Seeded synthetic inputs shaped like BASELINE.json's five configs.

Genomes and GFFs are made here (numpy); reads come from the native generator
(csrc/host/synth.c) straight into flat arrays, optionally round-tripped through a real BAM by
``bamio.write_bam`` / ``bamio.read_bam``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import bamio
from .reads import ReadBatch

BASES = "ACGT"
CODE_OF = {"A": 1, "C": 2, "G": 4, "T": 8, "N": 15}
STOPS = ("TAA", "TAG", "TGA")

VAR_SUB, VAR_INS, VAR_DEL = 0, 1, 2

# (name, start, end) 1-based inclusive, laid out like SARS-CoV-2 (NC_045512.2) CDS features
SARS2_LAYOUT = [
    ("ORF1ab", 266, 21555), ("S", 21563, 25384), ("ORF3a", 25393, 26220), ("E", 26245, 26472),
    ("M", 26523, 27191), ("ORF6", 27202, 27387), ("ORF7a", 27394, 27759), ("ORF8", 27894, 28259),
    ("N", 28274, 29533), ("ORF10", 29558, 29674),
]


class _Variant(C.Structure):
    _fields_ = [("pos", C.c_int32), ("kind", C.c_int32), ("len", C.c_int32), ("alt", C.c_int32), ("frac", C.c_double)]


class _SynthParams(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("n_reads", C.c_int64), ("ref_len", C.c_int32), ("read_len", C.c_int32),
        ("read_len_jitter", C.c_int32), ("paired", C.c_int32), ("insert_mean", C.c_int32), ("insert_sd", C.c_int32),
        ("n_amplicons", C.c_int32), ("amplicon_jitter", C.c_int32), ("indel_maxlen", C.c_int32),
        ("softclip_max", C.c_int32), ("qual_min", C.c_int32), ("qual_max", C.c_int32), ("n_variants", C.c_int32),
        ("sub_rate", C.c_double), ("indel_rate", C.c_double), ("softclip_rate", C.c_double), ("n_rate", C.c_double),
        ("iupac_rate", C.c_double), ("refskip_rate", C.c_double), ("special_flag_rate", C.c_double),
        ("lowmapq_rate", C.c_double), ("variants", C.POINTER(_Variant)),
    ]


@dataclass
class Variant:
    pos: int            # 0-based column
    kind: int           # VAR_SUB / VAR_INS / VAR_DEL
    len: int = 1
    alt: int = 0        # SUB: 4-bit code; INS: seed of the inserted bases
    frac: float = 1.0


@dataclass
class SynthParams:
    seed: int = 1
    n_reads: int = 1000
    ref_len: int = 0
    read_len: int = 150
    read_len_jitter: int = 0
    paired: bool = False
    insert_mean: int = 300
    insert_sd: int = 50
    n_amplicons: int = 0
    amplicon_jitter: int = 0
    indel_maxlen: int = 1
    softclip_max: int = 0
    qual_min: int = 2
    qual_max: int = 40
    sub_rate: float = 0.005
    indel_rate: float = 0.0
    softclip_rate: float = 0.0
    n_rate: float = 0.0
    iupac_rate: float = 0.0
    refskip_rate: float = 0.0
    special_flag_rate: float = 0.0
    lowmapq_rate: float = 0.01
    variants: list[Variant] = field(default_factory=list)


def encode_ref(ref: str) -> np.ndarray:
    lut = np.zeros(256, np.uint8)
    for ch, c in CODE_OF.items():
        lut[ord(ch)] = c
        lut[ord(ch.lower())] = c
    return lut[np.frombuffer(ref.encode(), dtype=np.uint8)]


def generate_reads(p: SynthParams, ref: str | np.ndarray, threads: int = 0, read_range: tuple[int, int] | None = None) -> ReadBatch:
    """The start-sorted synthetic read set — or, with ``read_range`` = (r0, r1), only its reads [r0, r1) (every read draws from
    its own stream: a shard equals the slice of the whole)."""
    lib = bamio.host_lib()
    if not hasattr(lib, "_synth_ready"):
        lib.tc_synth_reads_range.argtypes = [C.POINTER(_SynthParams), C.c_void_p, C.c_int, C.c_int64, C.c_int64,
                                             C.POINTER(bamio.TcHostReads), C.c_char_p, C.c_int]
        lib.tc_synth_reads_range.restype = C.c_int
        lib._synth_ready = True
    codes = encode_ref(ref) if isinstance(ref, str) else np.ascontiguousarray(ref, dtype=np.uint8)
    vs = sorted(p.variants, key=lambda v: v.pos)
    varr = (_Variant * max(len(vs), 1))()
    for i, v in enumerate(vs):
        varr[i] = _Variant(v.pos, v.kind, v.len, v.alt, v.frac)
    sp = _SynthParams(
        seed=p.seed, n_reads=p.n_reads, ref_len=p.ref_len or len(codes), read_len=p.read_len,
        read_len_jitter=p.read_len_jitter, paired=int(p.paired), insert_mean=p.insert_mean, insert_sd=p.insert_sd,
        n_amplicons=p.n_amplicons, amplicon_jitter=p.amplicon_jitter, indel_maxlen=p.indel_maxlen,
        softclip_max=p.softclip_max, qual_min=p.qual_min, qual_max=p.qual_max, n_variants=len(vs),
        sub_rate=p.sub_rate, indel_rate=p.indel_rate, softclip_rate=p.softclip_rate, n_rate=p.n_rate,
        iupac_rate=p.iupac_rate, refskip_rate=p.refskip_rate, special_flag_rate=p.special_flag_rate,
        lowmapq_rate=p.lowmapq_rate, variants=C.cast(varr, C.POINTER(_Variant)),
    )
    hr = bamio.TcHostReads()
    err = C.create_string_buffer(512)
    r0, r1 = (0, -1) if read_range is None else (int(read_range[0]), int(read_range[1]))
    rc = lib.tc_synth_reads_range(C.byref(sp), codes.ctypes.data, threads, r0, r1, C.byref(hr), err, len(err))
    if rc != 0:
        raise RuntimeError(f"tc_synth_reads failed ({rc}): {err.value.decode(errors='replace')}")
    b = bamio.batch_from_hostreads(hr, lib)
    b.ref_names = ["ref"]
    b.ref_lens = [sp.ref_len]
    return b


# ------------------------------------------------------------------------------ genomes / GFF
def _stop_free_orf(rng: np.random.Generator, seq: np.ndarray, s0: int, e0: int) -> None:
    """Make seq[s0:e0] (0-based half-open, length % 3 == 0) ATG ... sense codons ... TAA."""
    seq[s0:s0 + 3] = [ord(c) for c in "ATG"]
    for c in range(s0 + 3, e0 - 3, 3):
        cod = bytes(seq[c:c + 3]).decode()
        if cod in STOPS:
            seq[c] = ord("C")
    seq[e0 - 3:e0] = [ord(c) for c in "TAA"]


def make_genome(ref_len: int, seed: int, layout: str = "sars2"):
    """Random ACGT genome with engineered '+' CDS features.  Returns (sequence, features) where
    features are dicts with GFF columns (start/end 1-based inclusive; end = last base of the stop)."""
    rng = np.random.default_rng(seed)
    seq = np.frombuffer("ACGT".encode(), dtype=np.uint8)[rng.integers(0, 4, ref_len)].copy()
    feats = []
    if layout == "sars2":
        scale = ref_len / 29903.0
        for name, s, e in SARS2_LAYOUT:
            s1 = max(1, int(round(s * scale))); e1 = min(ref_len, int(round(e * scale)))
            n = (e1 - s1 + 1) // 3 * 3
            if n < 9:
                continue
            e1 = s1 + n - 1
            feats.append({"name": name, "start": s1, "end": e1, "strand": "+"})
    elif layout == "mpox":
        p = 200
        i = 0
        while p + 400 < ref_len - 200:
            n = int(rng.integers(100, 900)) * 3
            if p + n >= ref_len - 100:
                break
            strand = "+" if rng.random() < 0.6 else "-"
            feats.append({"name": f"OPG{i:03d}", "start": p + 1, "end": p + n, "strand": strand})
            p += n + int(rng.integers(40, 260))
            i += 1
    elif layout == "none":
        pass
    else:
        raise ValueError(layout)
    for f in feats:
        if f["strand"] == "+":
            _stop_free_orf(rng, seq, f["start"] - 1, f["end"])
    return bytes(seq).decode(), feats


def write_fasta(path: str, name: str, seq: str, width: int = 70) -> None:
    with open(path, "w") as fh:
        fh.write(f">{name}\n")
        for i in range(0, len(seq), width):
            fh.write(seq[i:i + width] + "\n")


def write_gff(path: str, seqid: str, ref_len: int, feats: list[dict]) -> None:
    with open(path, "w") as fh:
        fh.write("##gff-version 3\n")
        fh.write(f"##sequence-region {seqid} 1 {ref_len}\n")
        for f in feats:
            attrs = f"ID=cds-{f['name']};Name={f['name']};gbkey=CDS"
            fh.write("\t".join([seqid, "synthetic", f.get("type", "CDS"), str(f["start"]), str(f["end"]), ".",
                                f["strand"], "0", attrs]) + "\n")


def gff_dict(feats: list[dict], seqid: str = "ref") -> dict:
    """The ``GffDF.to_dict("index")`` shape the reference hands to BuildConsensus
    (TrueConsense/TrueConsense.py:238-241)."""
    out = {}
    for i, f in enumerate(feats):
        out[i] = {
            "seqid": seqid, "source": "synthetic", "type": f.get("type", "CDS"), "start": int(f["start"]),
            "end": int(f["end"]), "score": ".", "strand": f["strand"], "phase": "0",
            "attributes": f"ID=cds-{f['name']};Name={f['name']};gbkey=CDS",
        }
    return out


# ------------------------------------------------------------------------------ the five configs
@dataclass
class Workload:
    name: str
    ref: str
    feats: list[dict]
    params: SynthParams
    mincov: int = 30
    n_samples: int = 1
    note: str = ""


def _orf_interior(feats, k, off):
    f = [x for x in feats if x["strand"] == "+"][k]
    return f["start"] - 1 + 3 * off      # 0-based, codon aligned


def config(idx: int, scale: float = 1.0, seed: int | None = None, sample: int = 0) -> Workload:
    """BASELINE.json ``configs[idx]`` (idx 0..4).  ``scale`` multiplies the read count (tests use
    small scales; the benchmark uses 1.0).  ``sample`` selects the per-sample seed of config 2."""
    seed = (20260101 + idx) if seed is None else seed
    if idx in (0, 1, 2, 3):
        ref, feats = make_genome(29903, 20260101, "sars2")
    else:
        ref, feats = make_genome(197209, 20260105, "mpox")
    L = len(ref)
    if idx == 0:
        # 50k Illumina 2x150 (25k pairs, ~250x), mates overlap, a few hundred odd flags, soft clips
        vs = [
            Variant(_orf_interior(feats, 1, 100), VAR_SUB, 1, CODE_OF["T"], 0.97),
            Variant(_orf_interior(feats, 1, 200), VAR_SUB, 1, CODE_OF["G"], 0.50),   # ambiguity
            Variant(_orf_interior(feats, 0, 1500) + 2, VAR_INS, 3, 11, 0.90),        # in-frame insertion
            Variant(_orf_interior(feats, 0, 3000), VAR_DEL, 3, 0, 0.85),             # in-frame deletion
            Variant(_orf_interior(feats, 8, 50), VAR_DEL, 1, 0, 0.30),               # minority frameshift del
        ]
        p = SynthParams(seed=seed, n_reads=max(2, int(50_000 * scale)), ref_len=L, read_len=150, paired=True,
                        insert_mean=300, insert_sd=50, sub_rate=0.005, softclip_rate=0.02, softclip_max=10,
                        special_flag_rate=0.008, lowmapq_rate=0.01, n_rate=0.0005, variants=vs)
        return Workload("cfg1_illumina_50k_2x150", ref, feats, p, 30)
    if idx == 1:
        # 2M ONT 400-bp amplicon reads (~25,000x): 98 tiled amplicons, homopolymer-like indels ~1/30 bp,
        # 3 true insertions > 55 %, 2 at the threshold edge, 3-bp and 1-bp deletions at 15-90 %
        vs = [
            Variant(_orf_interior(feats, 0, 400) + 2, VAR_INS, 3, 21, 0.92),
            Variant(_orf_interior(feats, 1, 300) + 2, VAR_INS, 6, 22, 0.75),
            Variant(_orf_interior(feats, 8, 100) + 2, VAR_INS, 12, 23, 0.64),
            Variant(_orf_interior(feats, 0, 2500) + 2, VAR_INS, 3, 24, 0.565),
            Variant(_orf_interior(feats, 0, 4200) + 2, VAR_INS, 1, 25, 0.585),
            Variant(_orf_interior(feats, 0, 800), VAR_DEL, 3, 0, 0.90),
            Variant(_orf_interior(feats, 1, 600), VAR_DEL, 3, 0, 0.55),
            Variant(_orf_interior(feats, 1, 900), VAR_DEL, 1, 0, 0.16),
            Variant(_orf_interior(feats, 4, 40), VAR_DEL, 1, 0, 0.88),
            Variant(_orf_interior(feats, 8, 300), VAR_DEL, 2, 0, 0.45),
            Variant(_orf_interior(feats, 1, 1000), VAR_SUB, 1, CODE_OF["A"], 0.52),
        ]
        p = SynthParams(seed=seed, n_reads=max(1, int(2_000_000 * scale)), ref_len=L, read_len=400, read_len_jitter=20,
                        n_amplicons=98, amplicon_jitter=3, sub_rate=0.02, indel_rate=1.0 / 30.0, indel_maxlen=1,
                        softclip_rate=0.05, softclip_max=20, n_rate=0.0002, lowmapq_rate=0.01, qual_min=2, qual_max=40,
                        variants=vs)
        return Workload("cfg2_ont_2M_400bp_amplicon", ref, feats, p, 30)
    if idx == 2:
        # 96-sample plate, 1M x 150 bp per sample, distinct variants per sample
        rng = np.random.default_rng(977 + sample)
        vs = [Variant(int(rng.integers(300, L - 300)), VAR_SUB, 1, int(rng.choice([1, 2, 4, 8])), float(rng.uniform(0.4, 1.0)))
              for _ in range(12)]
        vs.append(Variant(_orf_interior(feats, 1, 50 + sample) + 2, VAR_INS, 3, 100 + sample, 0.9))
        vs.append(Variant(_orf_interior(feats, 0, 700 + 5 * sample), VAR_DEL, 3, 0, 0.8))
        p = SynthParams(seed=seed * 1000 + sample, n_reads=max(1, int(1_000_000 * scale)), ref_len=L, read_len=150,
                        sub_rate=0.005, softclip_rate=0.02, softclip_max=8, variants=vs)
        return Workload(f"cfg3_plate96_sample{sample:02d}_1M_150bp", ref, feats, p, 30, n_samples=96)
    if idx == 3:
        vs = [Variant(_orf_interior(feats, 1, 100), VAR_SUB, 1, CODE_OF["T"], 0.97),
              Variant(_orf_interior(feats, 0, 1500) + 2, VAR_INS, 3, 11, 0.90),
              Variant(_orf_interior(feats, 0, 3000), VAR_DEL, 3, 0, 0.85)]
        p = SynthParams(seed=seed, n_reads=max(1, int(50_000_000 * scale)), ref_len=L, read_len=150, sub_rate=0.005,
                        softclip_rate=0.02, softclip_max=8, variants=vs)
        return Workload("cfg4_ultradeep_50M_150bp", ref, feats, p, 30)
    if idx == 4:
        plus = [f for f in feats if f["strand"] == "+"]
        vs = []
        for k, dl, fr in ((3, 1, 0.9), (9, 2, 0.9), (15, 3, 0.9), (21, 4, 0.9), (27, 1, 0.2), (33, 2, 0.18)):
            if k < len(plus):
                vs.append(Variant(plus[k]["start"] - 1 + 60, VAR_DEL, dl, 0, fr))
        # minority-deletion completion: 1-bp 20 % deletion followed by a 2-bp 90 % deletion
        if len(plus) > 40:
            a = plus[40]["start"] - 1 + 90
            vs += [Variant(a, VAR_DEL, 1, 0, 0.2), Variant(a + 1, VAR_DEL, 2, 0, 0.9)]
        vs.append(Variant(plus[5]["start"] - 1 + 152, VAR_INS, 4, 77, 0.8))
        p = SynthParams(seed=seed, n_reads=max(1, int(98_500 * scale)), ref_len=L, read_len=10_000, read_len_jitter=1500,
                        sub_rate=0.02, indel_rate=1.0 / 40.0, indel_maxlen=2, softclip_rate=0.1, softclip_max=50,
                        lowmapq_rate=0.01, variants=vs)
        return Workload("cfg5_mpox_197kb_10kb_reads", ref, feats, p, 30)
    raise ValueError(f"config index {idx} out of range 0..4")
