"""Seeded inputs shared by oracle/make_golden.py and tests/ (TEST INFRASTRUCTURE ONLY).

* ``quirk_batch()``  — a hand-written read set hitting every CIGAR adjacency rule of the restated
  pileup and the reference classifier's corner cases.
* ``mini_workload(name)`` — BASELINE.json's configs shrunk to a 3-6 kb genome so the reference's
  O(ORF^2) walk finishes in seconds.
* ``random_walk_case(rng)`` — random count tables + GFF + insertion strings for the consensus walk.
"""
from __future__ import annotations

import numpy as np

from trueconsense_b200 import synth
from trueconsense_b200.reads import ReadBatch

QUIRK_REF_LEN = 120


def quirk_records() -> list[dict]:
    S = "ACGTACGTTGCAAGCTTAGGCCATATGCGCGATATCGGCTAACGTTAGCAT"
    recs = [
        dict(pos=0, cigar="10M2I5M", seq=S[:17]),
        dict(pos=0, cigar="5S10M3D5M2S", seq=S[:22], flag=16),
        dict(pos=1, cigar="3I10M", seq=S[:13]),
        dict(pos=2, cigar="10M3I", seq=S[:13]),
        dict(pos=3, cigar="10M3D2I5M", seq=S[:17]),                 # D followed by I: "*+2.." with the offset quirk
        dict(pos=3, cigar="10M3D2I", seq=S[:12], flag=16),          # ... running past the end of SEQ
        dict(pos=4, cigar="10M2I3D5M", seq=S[:17]),
        dict(pos=5, cigar="5M1D2D5M", seq=S[:10]),                  # 1D2D: second D carries no indel
        dict(pos=6, cigar="5M2P3I5M", seq=S[:13]),                  # pad then insertion
        dict(pos=6, cigar="5M3I2P2I5M", seq=S[:15], flag=16),
        dict(pos=7, cigar="5M10N5M", seq=S[:10]),                   # ref-skip forward  '>'
        dict(pos=7, cigar="5M10N5M", seq=S[:10], flag=16),          # ref-skip reverse  '<'
        dict(pos=8, cigar="5H5M", seq=S[:5]),
        dict(pos=8, cigar="5=3X2M", seq=S[:10]),
        dict(pos=9, cigar="10M", seq="*"),                          # SEQ '*': every base reads as N
        dict(pos=9, cigar="12M", seq="ACMGRSVTWYHK"),               # IUPAC codes count towards coverage only
        dict(pos=9, cigar="4M", seq="DBN=", flag=16),
        dict(pos=10, cigar="10M", seq=S[:10], flag=4),              # UNMAP with coordinates: never piled up
        dict(pos=10, cigar="10M", seq=S[:10], flag=0x100),          # secondary: counted by BuildIndex
        dict(pos=10, cigar="10M", seq=S[:10], flag=0x200),
        dict(pos=10, cigar="10M", seq=S[:10], flag=0x400),
        dict(pos=10, cigar="10M", seq=S[:10], flag=0x800),
        dict(pos=10, cigar="10M", seq=S[:10], flag=1),              # paired, not proper: an orphan
        dict(pos=11, cigar="2D5M", seq=S[:5]),                      # leading deletion
        dict(pos=11, cigar="1S2D5M", seq=S[:6]),
        dict(pos=12, cigar="5S", seq=S[:5]),                        # no reference span at all
        dict(pos=12, cigar="*", seq=S[:5]),
        dict(pos=12, cigar="4I", seq=S[:4]),
        dict(pos=13, cigar="5M2D3N4M", seq=S[:9]),                  # D directly followed by N
        dict(pos=13, cigar="5M3N2D4M", seq=S[:9], flag=16),         # N followed by D: indel reported on '>'
        dict(pos=14, cigar="5M3N2I4M", seq=S[:11]),                 # N followed by I
        dict(pos=15, cigar="6M1I1M1I1M1D1M", seq=S[:11]),
        dict(pos=20, cigar="30M", seq=S[:30], qual=[5] * 15 + [40] * 15),
        dict(pos=20, cigar="15M4I15M", seq=S[:34], qual=12),        # all below the ExtractInserts threshold
        dict(pos=20, cigar="15M4I15M", seq=S[:34], qual=13),
        dict(pos=60, cigar="30M", seq=S[:30]),
        dict(pos=119, cigar="1M", seq="A"),                         # last column of the reference
    ]
    for i, r in enumerate(recs):
        r.setdefault("qual", 20 + (i % 20))
        r.setdefault("qname", f"quirk{i}")
    return recs


def quirk_batch() -> ReadBatch:
    b = ReadBatch.from_records(quirk_records())
    b.ref_names = ["ref"]
    b.ref_lens = [QUIRK_REF_LEN]
    return b


def mini_workload(name: str):
    """Small-genome versions of the BASELINE configs: returns (Workload, ReadBatch)."""
    V = synth.Variant
    if name == "illumina":
        ref, feats = synth.make_genome(3000, 11, "sars2")
        a = feats[0]["start"] - 1
        vs = [V(a + 30, synth.VAR_SUB, 1, 8, 0.97), V(a + 60, synth.VAR_SUB, 1, 4, 0.5),
              V(a + 92, synth.VAR_INS, 3, 5, 0.9), V(a + 150, synth.VAR_DEL, 3, 0, 0.85),
              V(a + 210, synth.VAR_DEL, 1, 0, 0.3), V(a + 300, synth.VAR_DEL, 1, 0, 0.9),
              V(a + 360, synth.VAR_DEL, 2, 0, 0.9), V(a + 420, synth.VAR_INS, 12, 6, 0.8),
              V(a + 480, synth.VAR_INS, 1, 7, 0.56)]
        p = synth.SynthParams(seed=101, n_reads=4000, ref_len=3000, read_len=150, paired=True, insert_mean=300,
                              insert_sd=50, sub_rate=0.005, softclip_rate=0.03, softclip_max=10,
                              special_flag_rate=0.02, lowmapq_rate=0.02, n_rate=0.001, variants=vs)
        w = synth.Workload("mini_illumina", ref, feats, p, 30)
    elif name == "ont":
        ref, feats = synth.make_genome(3000, 12, "sars2")
        a = feats[0]["start"] - 1
        vs = [V(a + 44, synth.VAR_INS, 3, 21, 0.92), V(a + 104, synth.VAR_INS, 6, 22, 0.7),
              V(a + 200, synth.VAR_INS, 2, 23, 0.56), V(a + 300, synth.VAR_DEL, 3, 0, 0.9),
              V(a + 390, synth.VAR_DEL, 1, 0, 0.16), V(a + 450, synth.VAR_DEL, 1, 0, 0.88),
              V(a + 600, synth.VAR_SUB, 1, 1, 0.52), V(a + 700, synth.VAR_DEL, 2, 0, 0.5)]
        p = synth.SynthParams(seed=102, n_reads=3000, ref_len=3000, read_len=400, read_len_jitter=20, n_amplicons=9,
                              amplicon_jitter=3, sub_rate=0.02, indel_rate=1 / 30, indel_maxlen=1, softclip_rate=0.05,
                              softclip_max=20, n_rate=0.0005, iupac_rate=0.0002, refskip_rate=0.002, lowmapq_rate=0.01,
                              variants=vs)
        w = synth.Workload("mini_ont", ref, feats, p, 30)
    elif name == "long":
        ref, feats = synth.make_genome(6000, 13, "mpox")
        plus = [f for f in feats if f["strand"] == "+"]
        vs = []
        for k, (dl, fr) in enumerate(((1, 0.9), (2, 0.9), (3, 0.9), (4, 0.9), (1, 0.2))):
            if k < len(plus):
                vs.append(V(plus[k]["start"] - 1 + 60, synth.VAR_DEL, dl, 0, fr))
        if plus:
            a = plus[-1]["start"] - 1 + 90
            vs += [V(a, synth.VAR_DEL, 1, 0, 0.2), V(a + 1, synth.VAR_DEL, 2, 0, 0.9), V(a + 31, synth.VAR_INS, 4, 77, 0.8)]
        p = synth.SynthParams(seed=103, n_reads=300, ref_len=6000, read_len=2500, read_len_jitter=400, sub_rate=0.02,
                              indel_rate=1 / 40, indel_maxlen=2, softclip_rate=0.1, softclip_max=50, variants=vs)
        w = synth.Workload("mini_long", ref, feats, p, 30)
    else:
        raise KeyError(name)
    return w, synth.generate_reads(w.params, w.ref)


MINI_NAMES = ("illumina", "ont", "long")


def random_counts(rng: np.random.Generator, L: int, depth: int) -> np.ndarray:
    """Random count table [8][L] with the structures the walk branches on."""
    c = np.zeros((8, L), dtype=np.int64)
    for j in range(L):
        kind = rng.integers(0, 12)
        cov = int(rng.integers(0, depth + 1)) if rng.random() < 0.15 else depth
        if kind == 0:
            cov = 0
        row = np.zeros(5, dtype=np.int64)          # A T C G X
        if cov > 0:
            main = int(rng.integers(0, 4))
            if kind in (1, 2):                       # deletion dominated
                x = int(cov * rng.uniform(0.5, 1.0)); row[4] = x; row[main] = cov - x
            elif kind == 3:                          # minority deletion
                x = int(cov * rng.uniform(0.1, 0.3)); row[4] = x; row[main] = cov - x
            elif kind == 4:                          # two-way ambiguity
                o = (main + 1 + int(rng.integers(0, 3))) % 4
                a = cov // 2 + int(rng.integers(-cov // 8 - 1, cov // 8 + 2)); a = min(max(a, 0), cov)
                row[main] = a; row[o] = cov - a
            elif kind == 5:                          # three / four way
                parts = rng.multinomial(cov, [0.3, 0.3, 0.3, 0.1] if rng.random() < 0.5 else [0.25] * 4)
                row[:4] = parts
            elif kind == 6:                          # coverage exceeds the letters (N-rich)
                row[main] = int(cov * rng.uniform(0.0, 0.6))
            else:
                e = int(cov * rng.uniform(0, 0.08)); row[main] = cov - e; row[(main + 1) % 4] = e
        c[0, j] = cov; c[1:6, j] = row
        if cov > 0 and rng.random() < 0.12:
            c[6, j] = int(cov * rng.uniform(0.3, 1.0))
    return c


def random_walk_case(rng: np.random.Generator) -> dict:
    L = int(rng.integers(6, 61))
    mincov = int(rng.choice([0, 1, 5, 10, 30]))
    depth = int(rng.choice([10, 40, 100]))
    counts = random_counts(rng, L, depth)
    feats = {}
    for k in range(int(rng.integers(0, 4))):
        s = int(rng.integers(1, L)); e = int(rng.integers(s, L + 1))
        feats[k] = {"seqid": "s", "source": "x", "type": "CDS", "start": s, "end": e, "score": ".",
                    "strand": "+" if rng.random() < 0.75 else "-", "phase": "0", "attributes": f"ID=f{k};Name=f{k}"}
    cols = {}
    for j in range(L):
        if counts[6, j] > 0 and rng.random() < 0.9:
            n = int(rng.choice([1, 2, 3, 4, 12]))
            ins = "".join(rng.choice(list("ACGT"), n))
            base = "ACGT"[int(rng.integers(0, 4))]
            kind = rng.random()
            if kind < 0.7:
                strings = [f"{base}+{n}{ins}"] * 6 + [base] * 3 + [f"{base.lower()}+{n}{ins.lower()}"] * 2
            elif kind < 0.8:
                strings = [f"{base}-{n}{'N' * n}"] * 5 + [f"{base}+1T"] * 4
            elif kind < 0.9:
                strings = [base] * 5 + [f"{base}+{n}{ins}"] * 5
            else:
                strings = ""
            cols[j] = strings
    return dict(L=L, mincov=mincov, counts=counts[:7].tolist(), gff=feats, columns=cols,
                include_ambig=bool(rng.random() < 0.6))


# ------------------------------------------------------------------ mate overlaps (htslib tweak_overlap_quality)
OVERLAP_REF_LEN = 200
OVERLAP_COL = 100           # 0-based column every read below covers; most carry a 2-base insertion behind it


def x31_wang_bit(name: str) -> int:
    """khash's X31 string hash of a QNAME through __ac_Wang_hash, lowest bit: 1 = the first-arrived mate keeps its
    qualities (htslib >= 1.13 sam.c tweak_overlap_quality).  Written out here, independently of the C oracle."""
    bs = name.encode()
    h = bs[0]
    for ch in bs[1:]:
        h = (h * 31 + ch) & 0xFFFFFFFF
    k = h
    k = (k + (~(k << 15) & 0xFFFFFFFF)) & 0xFFFFFFFF
    k ^= k >> 10
    k = (k + (k << 3)) & 0xFFFFFFFF
    k ^= k >> 6
    k = (k + (~(k << 11) & 0xFFFFFFFF)) & 0xFFFFFFFF
    k ^= k >> 16
    return k & 1


def overlap_pair(name, base_a, qa, base_b, qb, cigar_a="41M2I19M", cigar_b="21M2I39M", pos_a=60, pos_b=80, proper=True, ins="TT"):
    """Two mates that both cover OVERLAP_COL (a: pos_a + 40, b: pos_b + 20 with the default CIGARs) with the given
    base / quality there; every other base is 'G' with quality 30."""
    import re

    def mk(pos, cigar, col_base, col_q, flag, mpos, isize):
        ops = [(int(l), o) for l, o in re.findall(r"(\d+)([MIDNSHP=X])", cigar)]
        seq, qual, x = [], [], pos
        for l, o in ops:
            if o in "M=X":
                for j in range(l):
                    hit = (x + j == OVERLAP_COL)
                    seq.append(col_base if hit else "G"); qual.append(col_q if hit else 30)
                x += l
            elif o in "IS":
                seq += list((ins * l)[:l]); qual += [30] * l
            elif o in "DN":
                x += l
        return dict(pos=pos, cigar=cigar, seq="".join(seq), qual=qual, flag=flag, qname=name, mpos=mpos, isize=isize)

    fa = 1 | (2 if proper else 0) | 0x20 | 0x40
    fb = 1 | (2 if proper else 0) | 0x10 | 0x80
    return [mk(pos_a, cigar_a, base_a, qa, fa, pos_b, 100), mk(pos_b, cigar_b, base_b, qb, fb, pos_a, -100)]


def overlap_kat_records():
    """(records sorted by position, expected strings at OVERLAP_COL under ExtractInserts' filters) — the expectations
    are derived by hand from htslib's rule, not from the C oracle:
      bases agree      -> the kept mate reads min(200, qa + qb), the other 0
      bases disagree   -> the better base keeps int(0.8 q), the other 0; equal qualities: the kept mate int(0.8 q)
      kept mate        -> the first-arrived one iff Wang(X31(QNAME)) is odd
    min_base_quality = 13 then decides which entries exist."""
    recs, exp = [], []
    for i in range(12):         # the unpaired majority: makes the column an insertion candidate
        recs.append(dict(pos=70, cigar="31M2I19M", seq="G" * 30 + "A" + "TT" + "G" * 19, qual=30, qname=f"solo{i}"))
        exp.append("A+2TT")

    def keeper(name):
        return 0 if x31_wang_bit(name) else 1       # index of the mate that keeps its qualities

    cases = [("agree_low", "A", 8, "A", 8), ("agree_high", "A", 30, "A", 30), ("agree_sum_below", "A", 6, "A", 6),
             ("dis_a_better", "A", 30, "C", 20), ("dis_b_better", "A", 14, "C", 35), ("dis_lost", "A", 15, "C", 10),
             ("dis_tie", "A", 20, "C", 20), ("dis_tie_lost", "A", 16, "C", 16)]
    for name, ba, qa, bb, qb in cases:
        recs += overlap_pair(name, ba, qa, bb, qb)
        k = keeper(name)
        strs = {0: f"{ba}+2TT", 1: f"{bb.lower()}+2tt"}          # mate b is on the reverse strand
        if ba == bb:
            if min(200, qa + qb) >= 13:
                exp.append(strs[k])
        elif qa > qb:
            if int(0.8 * qa) >= 13:
                exp.append(strs[0])
        elif qb > qa:
            if int(0.8 * qb) >= 13:
                exp.append(strs[1])
        elif int(0.8 * qa) >= 13:
            exp.append(strs[k])
    # not a proper pair: orphans are skipped by the samtools stepper altogether (ignore_orphans)
    recs += overlap_pair("orphans", "A", 30, "A", 30, proper=False)
    # the mate does not cover the column: it is never fetched, nothing is rewritten
    far = overlap_pair("mate_far", "A", 30, "A", 30, cigar_b="20M", pos_b=120)
    recs += far
    exp.append("A+2TT")
    recs.sort(key=lambda r: r["pos"])
    return recs, exp


def overlap_workload():
    """A 200-bp reference where the insertion behind OVERLAP_COL reads "TT" in 9 unpaired reads and "CC" in both mates of
    6 overlapping proper pairs: counted per entry "CC" wins 12 : 9, under htslib's mate-overlap rewriting (one entry per
    pair) "TT" wins 9 : 6 — ListInserts and the consensus differ.  Returns (name, refseq, feats, mincov, ReadBatch)."""
    rng = np.random.default_rng(77)
    refseq = "".join("ACGT"[i] for i in rng.integers(0, 4, OVERLAP_REF_LEN))
    recs = []
    for i in range(9):
        recs.append(dict(pos=70, cigar="31M2I19M", seq="G" * 30 + "A" + "TT" + "G" * 19, qual=30, qname=f"solo{i}"))
    for i in range(6):
        recs += overlap_pair(f"pair{i}", "A", 30, "A", 30, ins="CC")
    recs.sort(key=lambda r: r["pos"])
    b = ReadBatch.from_records(recs)
    b.ref_names = ["ref"]; b.ref_lens = [OVERLAP_REF_LEN]
    feats = [{"name": "o", "start": 10, "end": 189, "strand": "+"}]
    return "mini_overlap", refseq, feats, 5, b
