"""Independent pure-Python BAM reader (TEST INFRASTRUCTURE ONLY — see oracle/__init__.py).

BGZF is a series of gzip members, so ``gzip.open`` decompresses it; records are unpacked with
``struct``.  Shares no code with trueconsense_b200/csrc/host/bamio.c and is used to check it,
and to back the stand-in ``pysam`` module of oracle/ref_stubs.py.
"""
from __future__ import annotations

import gzip
import struct
from types import SimpleNamespace

import numpy as np


def _x31(name: bytes) -> int:
    h = name[0] if name else 0
    for ch in name[1:]:
        h = ((h << 5) - h + ch) & 0xFFFFFFFF
    return h


def _fnv_hi(name: bytes) -> int:
    f = 1469598103934665603
    for ch in name:
        f ^= ch
        f = (f * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return ((f >> 32) ^ f) & 0xFFFFFFFF


def read_bam(path: str) -> SimpleNamespace:
    """Flat arrays in the layout of tc_reads_t (placed records only, file order)."""
    with gzip.open(path, "rb") as fh:
        data = fh.read()
    if data[:4] != b"BAM\1":
        raise ValueError("not a BAM file")
    p = 4
    (l_text,) = struct.unpack_from("<i", data, p); p += 4
    text = data[p:p + l_text]; p += l_text
    (n_ref,) = struct.unpack_from("<i", data, p); p += 4
    names, lens = [], []
    for _ in range(n_ref):
        (l_name,) = struct.unpack_from("<i", data, p); p += 4
        names.append(data[p:p + l_name].rstrip(b"\0").decode()); p += l_name
        (l_ref,) = struct.unpack_from("<i", data, p); p += 4
        lens.append(l_ref)
    pos, flag, mapq, l_seq, tid, mtid, mpos, isize, qh = [], [], [], [], [], [], [], [], []
    seq_off, cig_off = [0], [0]
    seq_bytes = bytearray(); qual = bytearray(); cigar = []
    qnames = []
    while p + 4 <= len(data):
        (bs,) = struct.unpack_from("<i", data, p); p += 4
        rec = data[p:p + bs]; p += bs
        refid, rpos, l_name, mq, _bin, n_cig, fl, ls, nref, npos, tlen = struct.unpack_from("<iiBBHHHIiii", rec, 0)
        if refid < 0:
            continue
        name = rec[32:32 + l_name].rstrip(b"\0")
        q = 32 + l_name
        cg = struct.unpack_from(f"<{n_cig}I", rec, q); q += 4 * n_cig
        nb = (ls + 1) // 2
        sq = rec[q:q + nb]; q += nb
        ql = rec[q:q + ls]
        words = (ls + 7) // 8
        seq_bytes += sq + bytes(4 * words - nb)
        qual += ql + bytes(8 * words - ls)
        cigar.extend(cg)
        pos.append(rpos); flag.append(fl); mapq.append(mq); l_seq.append(ls); tid.append(refid); mtid.append(nref)
        mpos.append(npos); isize.append(tlen); qh.append((_fnv_hi(name) << 32) | _x31(name)); qnames.append(name.decode())
        seq_off.append(seq_off[-1] + words); cig_off.append(len(cigar))
    return SimpleNamespace(
        pos=np.array(pos, np.int32), flag=np.array(flag, np.uint16), mapq=np.array(mapq, np.uint8),
        l_seq=np.array(l_seq, np.int32), seq_off=np.array(seq_off, np.uint32), cigar_off=np.array(cig_off, np.uint32),
        seq4=np.frombuffer(bytes(seq_bytes), dtype=np.uint32).copy() if seq_bytes else np.zeros(0, np.uint32),
        qual=np.frombuffer(bytes(qual), dtype=np.uint8).copy() if qual else np.zeros(0, np.uint8),
        cigar=np.array(cigar, np.uint32), qname_hash=np.array(qh, np.uint64), mpos=np.array(mpos, np.int32),
        isize=np.array(isize, np.int32), tid=np.array(tid, np.int32), mtid=np.array(mtid, np.int32),
        ref_names=names, ref_lens=lens, header_text=text.decode(errors="replace"), qnames=qnames,
    )


def fasta_records(path: str) -> list[tuple[str, str]]:
    out: list[tuple[str, list[str]]] = []
    with open(path) as fh:
        for line in fh:
            if line.startswith(">"):
                f = line[1:].split()
                out.append((f[0] if f else "", []))
            elif out:
                out[-1][1].append(line.strip())
    return [(n, "".join(s)) for n, s in out]
