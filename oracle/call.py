"""CPU restatement of the per-position call and of the sequential consensus walk
(TEST INFRASTRUCTURE ONLY — see oracle/__init__.py).

Pure-Python, one position at a time, written from the behaviour of
  TrueConsense/Sequences.py:119-165  GetNucleotide / GetDistribution
  TrueConsense/Ambig.py:18-228       IsAmbiguous and helpers
  TrueConsense/Events.py:5-106       ListInserts / ExtractInserts / MinorityDel
  TrueConsense/Sequences.py:9-53,168-322  WalkForward / BuildConsensus
  TrueConsense/ORFs.py:1-192         in_orf / SolveTripletLength / CorrectStartPositions / CorrectGFF
It is pinned against the live reference (tests/test_oracle_vs_reference.py, runs where
/root/reference exists) and against tests/golden/*.json generated from the reference by
oracle/make_golden.py.

The index is handled as ``counts``: an int array [8][L] with rows coverage,A,T,C,G,X,I,pad
(row order of TrueConsense/indexing.py:134); position p (1-based) is column p-1.
"""
from __future__ import annotations

import copy
import re
from collections import Counter

import numpy as np

ROW = {"coverage": 0, "A": 1, "T": 2, "C": 3, "G": 4, "X": 5, "I": 6}
LETTERS = ("A", "T", "C", "G", "X")

# flag bits of tc_call_table_t.flags (include/trueconsense_b200.h)
CF_LOWCOV, CF_PRIMARY_X, CF_MINORITY_DEL, CF_INS_CANDIDATE = 0x01, 0x02, 0x04, 0x08
CF_COV_GT_MINCOV, CF_XRUN_OFF_END, CF_ZERO_COV, CF_AMBIG = 0x10, 0x20, 0x40, 0x80


def counts_from_index(index: dict, L: int | None = None) -> np.ndarray:
    """{pos: {"coverage","A","T","C","G","X","I"}} -> int64[8][L]."""
    L = len(index) if L is None else L
    out = np.zeros((8, L), dtype=np.int64)
    for p, row in index.items():
        for k, r in ROW.items():
            out[r, p - 1] = row[k]
    return out


def index_from_counts(counts: np.ndarray) -> dict:
    """The reference's ``IndexDF.to_dict("index")`` (TrueConsense.py:237)."""
    L = counts.shape[1]
    return {p: {k: int(counts[r, p - 1]) for k, r in ROW.items()} for p in range(1, L + 1)}


# ---------------------------------------------------------------------------- ranking
def ranked(counts: np.ndarray, p: int):
    """All five (letter, count) by decreasing (count, letter): Sequences.py:137-140 sorts
    (value, key) tuples ascending and indexes from the back, so ties go to the larger letter
    (X > T > G > C > A)."""
    col = [(int(counts[ROW[l], p - 1]), l) for l in LETTERS]
    col.sort()
    return [(l, c) for c, l in reversed(col)]


def get_nucleotide(counts: np.ndarray, p: int, k: int):
    if p < 1 or p > counts.shape[1]:
        raise KeyError(p)
    return ranked(counts, p)[k - 1]


# ---------------------------------------------------------------------------- ambiguity
_PAIR = {frozenset("AC"): "M", frozenset("AG"): "R", frozenset("AT"): "W", frozenset("CG"): "S",
         frozenset("CT"): "Y", frozenset("GT"): "K"}
_TRIPLE = {frozenset("ACG"): "V", frozenset("ACT"): "H", frozenset("AGT"): "D", frozenset("CGT"): "B"}


def is_ambiguous(one, two, three, four, cov, maxdist=10):
    """Ambig.py:179-228.  Percentages are (count / cov) * 100 in IEEE double, in that order of
    operations (Ambig.py:123-126); distances are compared with <= maxdist (Ambig.py:156-171)."""
    if cov == 0:
        return False, None
    (n1, c1), (n2, c2), (n3, c3), (n4, c4) = one, two, three, four
    if n1 == "X" or n2 == "X":
        return False, None
    p1, p2, p3, p4 = ((c / cov) * 100 for c in (c1, c2, c3, c4))
    if not abs(p1 - p2) <= maxdist:
        return False, None
    if abs(p1 - p3) <= maxdist and abs(p2 - p3) <= maxdist:
        if abs(p1 - p4) <= maxdist and abs(p2 - p4) <= maxdist and abs(p3 - p4) <= maxdist:
            return True, "N"
        if "X" in (n1, n2, n3):
            return True, "N"
        return True, _TRIPLE[frozenset((n1, n2, n3))]
    return True, _PAIR[frozenset((n1, n2))]


def minority_del(counts: np.ndarray, p: int, pct=15) -> bool:
    """Events.py:85-106."""
    if p < 1 or p > counts.shape[1]:
        raise KeyError(p)
    cov = int(counts[0, p - 1])
    dels = int(counts[5, p - 1])
    return (dels / cov) * 100 >= pct          # ZeroDivisionError when cov == 0, like the reference


def walk_forward(counts: np.ndarray, p: int) -> list[int]:
    """Sequences.py:44-52 (the "expand" mode): positions after p whose rank-1 letter is X."""
    track = []
    p += 1
    while True:
        n, _ = get_nucleotide(counts, p, 1)   # KeyError(L+1) when the run reaches the end
        if n != "X":
            return track
        track.append(p)
        p += 1


# ---------------------------------------------------------------------------- insertions
def insert_candidates(counts: np.ndarray, mincov: int, pct=55) -> list[int]:
    """Positions passing the ListInserts test, Events.py:25-36."""
    out = []
    for p in range(1, counts.shape[1] + 1):
        cov = int(counts[0, p - 1]); ins = int(counts[6, p - 1])
        if cov < mincov or cov == 0 or ins == 0:
            continue
        if (ins / cov) * 100 > pct:
            out.append(p)
    return out


def extract_insert(strings):
    """Events.py:67-81 applied to one column's strings (what get_query_sequences returned)."""
    if not strings:
        return None, None
    found = [s.upper() for s in strings]
    top = next(iter(dict(Counter(found).most_common())))
    m = re.search(r"(\d)([a-zA-Z]+)", top)
    if m:
        return m.group(2), m.group(1)
    return None, None


def list_inserts(counts: np.ndarray, mincov: int, column_strings):
    """Events.py:5-44.  ``column_strings(pos1)`` returns the strings of the column the reference
    would pile up for 1-based position pos1 (or None when pysam yields no column)."""
    positions = {}
    for p in insert_candidates(counts, mincov):
        strings = column_strings(p)
        bases, size = (None, None) if strings is None else extract_insert(strings)
        if bases is None or size is None:
            continue
        positions[p] = {size: bases}
    if not positions:
        return False, None
    return True, positions


# ---------------------------------------------------------------------------- call table
def call_table(counts: np.ndarray, mincov: int, include_ambig: bool, maxdist=10, mdel_pct=15, ins_pct=55) -> dict:
    """Everything tc_call emits, one position at a time (see tc_call_table_t)."""
    L = counts.shape[1]
    call_char = np.zeros(L, np.uint8); flags = np.zeros(L, np.uint8); xrun = np.zeros(L, np.int32)
    rank_letter = np.zeros((4, L), np.uint8); rank_count = np.zeros((4, L), np.int32); ambig = np.zeros(L, np.uint8)
    prim_x = np.zeros(L + 2, bool)
    for p in range(1, L + 1):
        r = ranked(counts, p)
        cov = int(counts[0, p - 1]); ins = int(counts[6, p - 1]); dels = int(counts[5, p - 1])
        for k in range(4):
            rank_letter[k, p - 1] = ord(r[k][0]); rank_count[k, p - 1] = r[k][1]
        amb, ch = is_ambiguous(r[0], r[1], r[2], r[3], cov, maxdist)
        f = 0
        if cov < mincov: f |= CF_LOWCOV
        if cov > mincov: f |= CF_COV_GT_MINCOV
        if cov == 0: f |= CF_ZERO_COV
        if r[0][0] == "X":
            f |= CF_PRIMARY_X; prim_x[p] = True
        if cov != 0 and (dels / cov) * 100 >= mdel_pct: f |= CF_MINORITY_DEL
        if not (cov < mincov) and cov != 0 and ins != 0 and (ins / cov) * 100 > ins_pct: f |= CF_INS_CANDIDATE
        if amb:
            f |= CF_AMBIG; ambig[p - 1] = ord(ch)
        if r[0][0] != "X":
            c = ch if (include_ambig and amb) else (r[0][0].lower() if r[0][1] < mincov else r[0][0].upper())
        else:
            c = r[1][0].lower() if r[1][1] < mincov else r[1][0].upper()
        call_char[p - 1] = ord(c)
        flags[p - 1] = f
    run = 0
    for p in range(L, 0, -1):
        # run = number of consecutive X-primary positions starting at p+1; WalkForward(p) runs off
        # the end (KeyError L+1) exactly when that run reaches position L
        xrun[p - 1] = run
        if run == L - p:
            flags[p - 1] |= CF_XRUN_OFF_END
        run = run + 1 if prim_x[p] else 0
    return dict(call_char=call_char, flags=flags, xrun=xrun, rank_letter=rank_letter, rank_count=rank_count,
                ambig_char=ambig)


# ---------------------------------------------------------------------------- GFF helpers
def in_orf(loc: int, gffd: dict) -> bool:
    """ORFs.py:1-26 — end exclusive."""
    return any(g["start"] <= loc < g["end"] for g in gffd.values())


def solve_triplet_length(uds, mds) -> bool:
    """ORFs.py:45-77."""
    if len(uds) % 3 == 0:
        return len(mds) % 3 == 0
    return (len(mds) + len(uds)) % 3 == 0


def correct_gff(oldgff: dict, newgff: dict, cons: list, p: int, inserts, mincov: int, cov: int) -> dict:
    """ORFs.py:111-192."""
    if inserts is not None and p in inserts and cov > mincov:
        shift_by = int(list(inserts[p].keys())[0])
        for g in newgff.values():                       # CorrectStartPositions, ORFs.py:80-108
            if g["start"] > p:
                g["start"] = int(g["start"]) + shift_by
    for k, g in newgff.items():
        start, end = g["start"], g["end"]
        if not (start <= p < end) or g.get("strand") != "+":
            continue
        rseq = "".join(cons)[start - 1:]
        shift = rseq.count("-")
        seq = rseq.replace("-", "")
        if cons[-1] == "-":
            g["end"] = oldgff[k]["end"]
            continue
        it = 0
        achieved = False
        for c0 in range(0, len(seq), 3):
            it += 1
            cod = seq[c0:c0 + 3]
            if "TAG" in cod or "TAA" in cod or "TGA" in cod:
                achieved = True
                break
        newend = start + it * 3 + shift - 1
        if not achieved:
            newend += 1
        if p == newend:
            g["end"] = newend
    return newgff


# ---------------------------------------------------------------------------- consensus walk
def build_consensus(mincov: int, counts: np.ndarray, gffdict: dict, include_ambig: bool, inserts, include_ins: bool):
    """Sequences.py:168-322.  ``inserts`` is ListInserts' return value ``(has, positions)``.
    Returns (consensus string, corrected gff dict); raises what the reference raises."""
    has_ins, ins_pos = inserts
    L = counts.shape[1]
    cons: list[str] = []
    dskips: set[int] = set()
    newgff = copy.deepcopy(gffdict)

    def default_char(r, amb, ch):
        if include_ambig and amb:
            return ch
        return r[0][0].lower() if r[0][1] < mincov else r[0][0].upper()

    for b in range(1, L + 1):
        cov = int(counts[0, b - 1])
        within = in_orf(b, newgff)
        if b in dskips:
            cons.append("-")
            newgff = correct_gff(gffdict, newgff, cons, b, ins_pos, mincov, cov)
            continue
        if cov < mincov:
            cons.append("N")
            newgff = correct_gff(gffdict, newgff, cons, b, ins_pos, mincov, cov)
            continue
        r = ranked(counts, b)
        amb, ch = is_ambiguous(r[0], r[1], r[2], r[3], cov)
        if r[0][0] != "X":
            out = default_char(r, amb, ch)
            if minority_del(counts, b):
                run = walk_forward(counts, b)
                if run:
                    if solve_triplet_length(run, [b]):
                        out = "-"
                        dskips.update([b] + run)
                elif minority_del(counts, b + 1):
                    run2 = walk_forward(counts, b + 1)
                    if run2 and solve_triplet_length(run2, [b, b + 1]):
                        out = "-"
                        dskips.update([b, b + 1] + run2)
            cons.append(out)
        else:
            if within:
                run = walk_forward(counts, b)
                if len(run) >= 2:
                    cons.append("-")
                    dskips.update([b] + run)
                else:
                    if include_ambig and amb:      # unreachable: IsAmbiguous is False when rank 1 is X
                        cons.append(ch)
                    else:
                        cons.append(r[1][0].lower() if r[1][1] < mincov else r[1][0].upper())
            else:
                cons.append("-")
        if include_ins and cov > mincov and has_ins is True:
            if b in ins_pos:
                for size in ins_pos[b]:
                    cons.append(str(ins_pos[b][size]))
        newgff = correct_gff(gffdict, newgff, cons, b, ins_pos, mincov, cov)
    return "".join(cons), newgff
