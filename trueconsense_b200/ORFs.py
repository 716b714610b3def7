"""Host-side ORF helpers with the reference's names and results (TrueConsense/ORFs.py).

These stay on the host by design (north_star: the sequential ORF / frameshift correction consumes
the GPU's per-position candidate table).  ``CorrectGFF`` is kept call-compatible with the
reference; the consensus walk itself uses :class:`GffTracker`, an incremental restatement that
does the same stop-codon bookkeeping in O(1) amortised per position instead of re-joining and
re-scanning the growing consensus at every position (ORFs.py:156-188 is O(ORF length) per call).
"""
from __future__ import annotations

STOP_CODONS = ("TAG", "TAA", "TGA")


def in_orf(loc, gffd):
    """True if ``loc`` lies in [start, end) of any feature (ORFs.py:1-26)."""
    for g in gffd.values():
        if g.get("start") <= loc < g.get("end"):
            return True
    return False


def split_to_codons(seq):
    return [seq[i:i + 3] for i in range(0, len(seq), 3)]


def SolveTripletLength(uds, mds):
    """ORFs.py:45-77: may the minority-deletion group ``mds`` join the upcoming deletion run ``uds``?"""
    if len(uds) % 3 == 0:
        return len(mds) % 3 == 0
    return (len(mds) + len(uds)) % 3 == 0


def CorrectStartPositions(gffd, shifts, p):
    """ORFs.py:80-108: push the start of every feature beginning after ``p`` by ``shifts``."""
    for g in gffd.values():
        start = g.get("start")
        if start > p:
            g.update({"start": int(start) + int(shifts)})
    return gffd


def CorrectGFF(oldgffdict, newgffdict, cons, p, inserts, mincov, cov):
    """ORFs.py:111-192, same arguments and result.  Plain restatement (re-scans the consensus); the
    walk in Sequences.BuildConsensus uses GffTracker instead."""
    if inserts is not None and p in inserts and cov > mincov:
        newgffdict = CorrectStartPositions(newgffdict, list(inserts[p].keys())[0], p)
    joined = None
    for k, g in newgffdict.items():
        start, end = g.get("start"), g.get("end")
        if not (start <= p < end) or g.get("strand") != "+":
            continue
        if joined is None:
            joined = "".join(cons)
        rseq = joined[start - 1:]
        shift = rseq.count("-")
        if cons[-1] == "-":
            g.update({"end": oldgffdict[k].get("end")})
            continue
        seq = rseq.replace("-", "")
        it = 0
        achieved = False
        for i in range(0, len(seq), 3):
            it += 1
            if seq[i:i + 3] in STOP_CODONS:
                achieved = True
                break
        newend = start + it * 3 + shift - 1
        if not achieved:
            newend += 1
        if p == newend:
            g.update({"end": newend})
    return newgffdict


class _FeatState:
    __slots__ = ("next_idx", "shift", "nondash", "stop_it", "c0", "c1")

    def __init__(self, next_idx):
        self.next_idx = next_idx    # next index of the joined consensus this feature has not consumed yet
        self.shift = 0              # '-' characters seen from start-1 on
        self.nondash = 0            # other characters seen
        self.stop_it = 0            # 1-based codon index of the first stop codon (0: none yet)
        self.c0 = ""                # pending codon characters
        self.c1 = ""


class GffTracker:
    """Incremental CorrectGFF: same updates of ``start`` / ``end`` as ORFs.py:111-192 applied after
    every position, but each '+' feature carries the running state of its own scan (dash count,
    codon phase, first stop) over the characters appended since its last evaluation."""

    def __init__(self, oldgffdict, newgffdict):
        self.old = oldgffdict
        self.new = newgffdict
        self.keys = list(newgffdict.keys())
        self.plus = [k for k in self.keys if newgffdict[k].get("strand") == "+"]
        self.state: dict = {}
        self.chars: list[str] = []     # the joined consensus, one character per element

    def append(self, s: str) -> None:
        self.chars.extend(s)

    def in_orf(self, loc: int) -> bool:
        new = self.new
        for k in self.keys:
            g = new[k]
            if g["start"] <= loc < g["end"]:
                return True
        return False

    def correct(self, p: int, last_element: str, inserts, mincov: int, cov: int) -> None:
        new = self.new
        if inserts is not None and p in inserts and cov > mincov:
            shift_by = int(list(inserts[p].keys())[0])
            for k in self.keys:
                g = new[k]
                if g["start"] > p:
                    g["start"] = int(g["start"]) + shift_by
        chars = self.chars
        n = len(chars)
        for k in self.plus:
            g = new[k]
            start = g["start"]
            if not (start <= p < g["end"]):
                continue
            st = self.state.get(k)
            if st is None:
                st = self.state[k] = _FeatState(start - 1)
            i = st.next_idx
            if i < n:
                shift, nondash, stop_it, c0, c1 = st.shift, st.nondash, st.stop_it, st.c0, st.c1
                while i < n:
                    ch = chars[i]
                    i += 1
                    if ch == "-":
                        shift += 1
                        continue
                    nondash += 1
                    if stop_it:
                        continue
                    if not c0:
                        c0 = ch
                    elif not c1:
                        c1 = ch
                    else:
                        if c0 == "T" and ((c1 == "A" and (ch == "G" or ch == "A")) or (c1 == "G" and ch == "A")):
                            stop_it = nondash // 3
                        c0 = c1 = ""
                st.next_idx, st.shift, st.nondash, st.stop_it, st.c0, st.c1 = i, shift, nondash, stop_it, c0, c1
            if last_element == "-":
                g["end"] = self.old[k].get("end")
                continue
            if st.stop_it:
                newend = start + st.stop_it * 3 + st.shift - 1
            else:
                newend = start + ((st.nondash + 2) // 3) * 3 + st.shift
            if p == newend:
                g["end"] = newend
