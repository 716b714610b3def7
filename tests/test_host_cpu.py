"""CPU-only tests (no GPU): host logic, the BAM reader/writer, repository layout, and that the
C-ABI library loads and exports every symbol include/*.h declares (no compute calls)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLD, ROOT, load_golden_json


# ---------------------------------------------------------------------------- C-ABI surface
def _declared_symbols(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tc_[a-z0-9_]+)\s*\(", text)))


def test_cuda_library_exports_every_declared_symbol():
    from trueconsense_b200 import build

    path = build.build_cuda()           # nvcc cross-compiles sm_100a without a GPU
    lib = ctypes.CDLL(path)
    declared = _declared_symbols("trueconsense_b200.h")
    assert "tc_pileup_counts" in declared and "tc_extract_inserts" in declared and len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/trueconsense_b200.h but not exported"
    lib.tc_abi_version.restype = ctypes.c_int
    assert lib.tc_abi_version() == 3


def test_cuda_library_is_sm100a_only():
    from trueconsense_b200 import build

    out = subprocess.run(["cuobjdump", "--list-elf", build.build_cuda()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_host_library_exports_every_declared_symbol():
    from trueconsense_b200 import build

    lib = ctypes.CDLL(build.build_host())
    for name in _declared_symbols("tc_host.h"):
        assert hasattr(lib, name), name


def test_no_device_fails_loudly():
    """Without a CUDA device the product raises; it never computes on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from trueconsense_b200 import gpu

    with pytest.raises(gpu.TcError) as ei:
        gpu.Context(0)
    assert ei.value.code == -6


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "trueconsense_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text and "pileup_oracle" not in text, f
    code = ("import sys; sys.path.insert(0, %r); import trueconsense_b200.indexing, trueconsense_b200.Sequences, "
            "trueconsense_b200.Events, trueconsense_b200.Ambig, trueconsense_b200.Coverage, trueconsense_b200.TrueConsense; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'" % ROOT)
    subprocess.run([sys.executable, "-c", code], check=True)


# ---------------------------------------------------------------------------- BAM io
def test_bam_roundtrip_and_independent_reader(tmp_path, host_libs):
    from oracle import bam_py
    from trueconsense_b200 import bamio, synth

    w = synth.config(0, scale=0.01)
    b = synth.generate_reads(w.params, w.ref)
    assert b.sorted and np.all(np.diff(b.pos.astype(np.int64)) >= 0)
    path = str(tmp_path / "x.bam")
    bamio.write_bam(path, b, "ref", len(w.ref), level=6)
    c = bamio.read_bam(path)
    p = bam_py.read_bam(path)
    assert c.ref_names == ["ref"] and c.ref_lens == [len(w.ref)] and c.sorted
    for name in ("pos", "flag", "mapq", "l_seq", "seq_off", "cigar_off", "seq4", "qual", "cigar", "mpos", "isize"):
        assert np.array_equal(getattr(b, name), getattr(c, name)), name
        assert np.array_equal(getattr(c, name), getattr(p, name)), name
    assert np.array_equal(c.qname_hash, p.qname_hash)
    assert c.aligned_bases == b.aligned_bases == b.count_aligned_bases(0)
    # mates share their name hash
    paired = (c.flag & 1) != 0
    assert paired.any()
    _, counts = np.unique(c.qname_hash[paired], return_counts=True)
    assert set(counts.tolist()) <= {1, 2} and (counts == 2).sum() > 10


def test_bam_reader_golden_fixtures(host_libs):
    from oracle import bam_py
    from trueconsense_b200 import bamio

    for name in ("quirk", "mini_illumina", "mini_ont", "mini_long"):
        c = bamio.read_bam(f"{GOLD}/{name}.bam")
        p = bam_py.read_bam(f"{GOLD}/{name}.bam")
        for arr in ("pos", "flag", "mapq", "l_seq", "seq_off", "cigar_off", "seq4", "qual", "cigar"):
            assert np.array_equal(getattr(c, arr), getattr(p, arr)), (name, arr)
        meta = load_golden_json(f"{name}.json")
        assert c.n_reads == meta["n_reads"]


def test_bam_reader_errors(tmp_path, host_libs):
    from trueconsense_b200 import bamio

    bad = tmp_path / "bad.bam"
    bad.write_bytes(b"this is not a bam file, not even gzip" * 4)
    with pytest.raises(OSError):
        bamio.read_bam(str(bad))
    with pytest.raises(OSError):
        bamio.read_bam(str(tmp_path / "missing.bam"))


def _bgzf_block(payload: bytes, isize=None, bsize_delta=0, xlen_extra=b"") -> bytes:
    import struct
    import zlib

    comp = zlib.compressobj(6, zlib.DEFLATED, -15)
    data = comp.compress(payload) + comp.flush()
    extra = b"BC" + struct.pack("<HH", 2, 0) + xlen_extra          # BSIZE patched below
    bsize = 12 + len(extra) + len(data) + 8
    extra = b"BC" + struct.pack("<HH", 2, bsize - 1 + bsize_delta) + xlen_extra
    hdr = bytes([31, 139, 8, 4, 0, 0, 0, 0, 0, 255]) + struct.pack("<H", len(extra)) + extra
    return hdr + data + struct.pack("<Ii", zlib.crc32(payload) & 0xFFFFFFFF, len(payload) if isize is None else isize)


def test_bam_reader_rejects_crafted_sizes(tmp_path, host_libs):
    """Every size the reader takes from the file is checked before it is used: a crafted ISIZE pair that
    under-allocated the inflate buffer, a BSIZE shorter than the block's own header, subfields past XLEN, negative
    header lengths, records shorter than their fields, truncation and random corruption all end in an error, never in
    a crash."""
    import struct

    from trueconsense_b200 import bamio

    def header(n_ref=1, l_text=0, l_name=4):
        return b"BAM\1" + struct.pack("<i", l_text) + struct.pack("<i", n_ref) + struct.pack("<i", l_name) + b"ref\0" + struct.pack("<i", 1000)

    def record(l_seq=4, n_cig=1, bs_delta=0, l_name=3, name=b"r1\0"):
        body = struct.pack("<iiBBHHHIiii", 0, 10, l_name, 60, 0, n_cig, 0, l_seq, -1, -1, 0) + name
        body += struct.pack("<I", (4 << 4) | 0) * 1 + bytes((4 + 1) // 2) + bytes(4)
        return struct.pack("<i", len(body) + bs_delta) + body

    eof = _bgzf_block(b"")
    cases = {
        "isize_pair": _bgzf_block(b"x" * 65000, isize=65000) + _bgzf_block(b"y" * 1000, isize=-64000) + eof,
        "isize_huge": _bgzf_block(header() + record(), isize=70000) + eof,
        "bsize_short": _bgzf_block(header() + record(), bsize_delta=-40) + eof,
        "subfield_past_xlen": _bgzf_block(header() + record())[:10] + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 200, 0) + eof,
        "negative_l_text": _bgzf_block(header(l_text=-8) + record()) + eof,
        "negative_n_ref": _bgzf_block(header(n_ref=-1) + record()) + eof,
        "huge_n_ref": _bgzf_block(header(n_ref=1 << 28) + record()) + eof,
        "negative_l_name": _bgzf_block(header(l_name=-4) + record()) + eof,
        "record_l_seq": _bgzf_block(header() + record(l_seq=4000)) + eof,
        "record_n_cig": _bgzf_block(header() + record(n_cig=60000)) + eof,
        "record_block_size": _bgzf_block(header() + record(bs_delta=5000)) + eof,
        "record_name_unterminated": _bgzf_block(header() + record(name=b"r1x")) + eof,
    }
    good = _bgzf_block(header() + record()) + eof
    path = tmp_path / "good.bam"
    path.write_bytes(good)
    assert bamio.read_bam(str(path)).n_reads == 1
    for name, blob in cases.items():
        path = tmp_path / f"{name}.bam"
        path.write_bytes(blob)
        with pytest.raises((ValueError, RuntimeError, OSError)):
            bamio.read_bam(str(path))
    # truncations and byte flips of a real file
    real = open(f"{GOLD}/mini_illumina.bam", "rb").read()
    rng = np.random.default_rng(9)
    for i in range(40):
        blob = bytearray(real[: int(rng.integers(30, len(real)))]) if i % 2 else bytearray(real)
        for _ in range(int(rng.integers(1, 6))):
            blob[int(rng.integers(0, len(blob)))] = int(rng.integers(0, 256))
        path = tmp_path / "fuzz.bam"
        path.write_bytes(bytes(blob))
        try:
            bamio.read_bam(str(path))
        except (ValueError, RuntimeError, OSError):
            pass
    # a CIGAR kept in the CG tag (placeholder <l_seq>S<span>N) is refused, not piled up as a reference skip
    body = struct.pack("<iiBBHHHIiii", 0, 10, 3, 60, 0, 2, 0, 4, -1, -1, 0) + b"r1\0"
    body += struct.pack("<II", (4 << 4) | 4, (100 << 4) | 3) + bytes(2) + bytes(4)
    path = tmp_path / "cg.bam"
    path.write_bytes(_bgzf_block(header() + struct.pack("<i", len(body)) + body) + eof)
    with pytest.raises(OSError) as ei:
        bamio.read_bam(str(path))
    assert "CG tag" in str(ei.value)


def test_readbatch_slice_and_records():
    from trueconsense_b200.reads import ReadBatch

    recs = [dict(pos=i, cigar="3S10M2I5M1D4M", seq="ACGT" * 6, qual=list(range(24))) for i in range(10)]
    b = ReadBatch.from_records(recs)
    assert b.cigar_string(3) == "3S10M2I5M1D4M" and b.seq_string(3) == "ACGT" * 6
    assert b.ref_spans().tolist() == [20] * 10
    s = b.slice(4, 7)
    s.validate()
    assert s.n_reads == 3 and s.pos.tolist() == [4, 5, 6] and s.cigar_string(0) == b.cigar_string(4)
    assert s.seq_string(2) == b.seq_string(6)
    assert b.algorithmic_bytes(100) == 10 * (12 + 4 * 6 + 16) + 3200


# ---------------------------------------------------------------------------- host walk bookkeeping
def test_gff_tracker_matches_plain_correct_gff():
    """The incremental stop-codon tracker of the host walk gives the same GFF updates as the plain
    re-scan (both restate TrueConsense/ORFs.py:111-192; the plain one is the oracle's)."""
    import copy

    from oracle import call
    from trueconsense_b200 import ORFs

    rng = np.random.default_rng(9)
    for case in range(300):
        L = int(rng.integers(8, 90))
        gff = {}
        for k in range(int(rng.integers(1, 4))):
            s = int(rng.integers(1, L))
            e = int(rng.integers(s, L + 1))
            gff[k] = {"start": s, "end": e, "strand": "+" if rng.random() < 0.8 else "-", "attributes": f"ID=f{k}"}
        inserts = None
        if rng.random() < 0.5:
            inserts = {int(p): {str(int(rng.integers(1, 10))): "".join(rng.choice(list("ACGT"), int(rng.integers(1, 14))))}
                       for p in rng.choice(np.arange(1, L + 1), size=min(3, L), replace=False)}
        new_a, new_b, new_c = copy.deepcopy(gff), copy.deepcopy(gff), copy.deepcopy(gff)
        tracker = ORFs.GffTracker(gff, new_b)
        cons_a = []
        for p in range(1, L + 1):
            ch = str(rng.choice(list("ACGTacgt-N"), p=[.18, .14, .14, .18, .03, .03, .03, .03, .2, .04]))
            cov = int(rng.choice([0, 5, 50]))
            cons_a.append(ch); tracker.append(ch)
            last = ch
            if inserts and p in inserts and rng.random() < 0.7 and cov > 10:
                s_ins = list(inserts[p].values())[0]
                cons_a.append(s_ins); tracker.append(s_ins)
                last = s_ins
            new_a = call.correct_gff(gff, new_a, cons_a, p, inserts, 10, cov)
            tracker.correct(p, last, inserts, 10, cov)
            new_c = ORFs.CorrectGFF(gff, new_c, cons_a, p, inserts, 10, cov)
            assert {k: (v["start"], v["end"]) for k, v in new_a.items()} == {k: (v["start"], v["end"]) for k, v in new_b.items()}, (case, p)
            assert {k: (v["start"], v["end"]) for k, v in new_a.items()} == {k: (v["start"], v["end"]) for k, v in new_c.items()}, (case, p)


def test_orf_helpers():
    from trueconsense_b200 import ORFs

    g = {0: {"start": 4, "end": 24, "strand": "+"}}
    assert ORFs.in_orf(4, g) and ORFs.in_orf(23, g) and not ORFs.in_orf(24, g) and not ORFs.in_orf(3, g)
    assert ORFs.SolveTripletLength([1, 2, 3], [0]) is False
    assert ORFs.SolveTripletLength([1, 2], [0]) is True
    assert ORFs.SolveTripletLength([1], [0, 9]) is True
    assert ORFs.SolveTripletLength([1, 2, 3, 4], [0, 9]) is True
    assert ORFs.split_to_codons("ATGAAAT") == ["ATG", "AAA", "T"]
    h = {0: {"start": 10}, 1: {"start": 3}}
    ORFs.CorrectStartPositions(h, "3", 5)
    assert h[0]["start"] == 13 and h[1]["start"] == 3


def test_synthetic_configs_shapes(host_libs):
    from trueconsense_b200 import synth

    for idx, scale in ((0, 0.02), (1, 0.001), (2, 0.001), (3, 0.0001), (4, 0.002)):
        w = synth.config(idx, scale=scale)
        b = synth.generate_reads(w.params, w.ref)
        b.validate()
        assert b.n_reads > 0 and b.sorted and b.aligned_bases == b.count_aligned_bases(0)
        assert int(b.pos.max()) < len(w.ref) and int((b.pos.astype(np.int64) + b.ref_spans()).max()) <= len(w.ref)
    # deterministic for a fixed seed, independent of the thread count
    w = synth.config(1, scale=0.002)
    a1 = synth.generate_reads(w.params, w.ref, threads=1)
    a8 = synth.generate_reads(w.params, w.ref, threads=8)
    assert np.array_equal(a1.seq4, a8.seq4) and np.array_equal(a1.cigar, a8.cigar) and np.array_equal(a1.qual, a8.qual)


# ---------------------------------------------------------------------------- CLI surface and writers
def test_cli_argument_surface(tmp_path, capsys, monkeypatch):
    """Same flags, validators and exit behaviour as TrueConsense/TrueConsense.py:25-209."""
    from trueconsense_b200 import TrueConsense

    bam = tmp_path / "a.bam"; fa = tmp_path / "r.fasta"; gff = tmp_path / "f.gff"; txt = tmp_path / "a.txt"
    for p in (bam, fa, gff, txt):
        p.write_text("x")
    base = ["-i", str(bam), "-ref", str(fa), "-gff", str(gff), "-cov", "30", "-name", "s", "-o", str(tmp_path / "c.fa")]
    a = TrueConsense.GetArgs(base + ["-vcf", "v.vcf", "-doc", "d.tsv", "-ogff", "o.gff", "-t", "3", "-noambig"])
    assert (a.input, a.reference, a.features, a.coverage_level, a.samplename) == (str(bam), str(fa), str(gff), 30, "s")
    assert (a.variants, a.depth_of_coverage, a.output_gff, a.threads, a.noambiguity, a.index_override) == ("v.vcf", "d.tsv", "o.gff", 3, True, None)
    with pytest.raises(SystemExit) as e:
        TrueConsense.GetArgs(["-i", str(tmp_path / "missing.bam")] + base[2:])
    assert e.value.code == -1
    with pytest.raises(SystemExit) as e:
        TrueConsense.GetArgs(base[:2] + ["-ref", str(tmp_path / "missing.fa")] + base[4:])
    assert e.value.code == 1
    with pytest.raises(SystemExit) as e:
        TrueConsense.GetArgs(["-i", str(txt)] + base[2:])          # wrong extension -> parser.error
    assert e.value.code == 2
    with pytest.raises(SystemExit) as e:
        TrueConsense.GetArgs(base[:-2])                             # --output is required
    assert e.value.code == 2
    monkeypatch.setattr("sys.argv", ["TrueConsense"])
    with pytest.raises(SystemExit) as e:
        TrueConsense.main([])
    assert e.value.code == 1
    with pytest.raises(SystemExit) as e:
        TrueConsense.GetArgs(["--version"])
    assert e.value.code == 0
    capsys.readouterr()


def test_gff_writer_and_fasta_reader(tmp_path):
    from trueconsense_b200 import Outputs, indexing

    g = indexing.Gffindex(f"{GOLD}/mini_ont.gff")
    df = g.df
    df["seqid"] = "mini_ont"
    out = tmp_path / "o.gff"
    Outputs.WriteGFF(g.header, df.to_dict("index"), str(out), "mini_ont")
    # the reference CLI run on the same GFF moved no coordinate of this fixture's features except through
    # CorrectGFF; lines whose coordinates are unchanged must be byte-identical
    gold = open(f"{GOLD}/cli_mini_ont.gff").read().splitlines()
    mine = out.read_text().splitlines()
    assert mine[:2] == gold[:2] and len(mine) == len(gold)
    assert all(a.split("\t")[:3] == b.split("\t")[:3] and a.split("\t")[5:] == b.split("\t")[5:] for a, b in zip(mine, gold))
    feat = {"seqid": "s", "source": "x", "type": "CDS", "start": 1, "end": 9, "score": ".", "strand": "+", "phase": 0,
            "attributes": "ID=a;Name=b;", "Extra": 5}
    assert Outputs._gff_line(feat) == "s\tx\tCDS\t1\t9\t.\t+\t0\tID=a;Name=b;extra=5\n"
    with pytest.raises(ValueError):
        Outputs._gff_line(dict(feat, attributes="ID=a=b"))
    fa = tmp_path / "r.fa"
    fa.write_text(">chr1 some description\nACGT\nacgt\n>chr2\nTTTT\n")
    assert Outputs._first_fasta_record(str(fa)) == ("chr1", list("ACGTacgt"))


def test_bench_reference_arm_line(host_libs):
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the contract's keys,
    on a tiny slice so the CPU suite stays short."""
    import json
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--scale", "0.005", "--ref-reads", "2000"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "aligned_bases_per_sec_pileup_consensus"
    assert line["unit"] == "aligned bases/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["workload"].startswith("cfg2")


def test_bench_needs_a_gpu_and_says_so():
    """Our arm of bench.py has no CPU path: without a CUDA device it stops with a message instead of printing a line."""
    import subprocess
    import sys

    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--scale", "0.001"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode != 0 and "CUDA device" in (out.stderr + out.stdout)
    assert not out.stdout.strip().startswith("{")


def test_clock_sampler_reports_only_samples_inside_the_timed_windows():
    import importlib.util
    import time

    spec = importlib.util.spec_from_file_location("tc_bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    class Quiet:
        stdout = ()

        def terminate(self): pass
        def wait(self, timeout=None): return 0
        def kill(self): pass

    s = bench.ClockSampler(0)
    s.proc = Quiet()
    now = time.time()
    row = lambda mhz, cap: ["0", str(mhz), "1965", "700", "0x0", "Not Active", "Not Active", "Not Active", cap]
    s.rows = [(now - 10.0, row(300, "Not Active")),            # idle, long before the timed region: ignored
              (now - 1.00, row(1965, "Not Active")), (now - 0.98, row(1950, "Active")),
              (now + 30.0, row(200, "Not Active"))]
    s.windows = [[now - 1.01, now - 0.97]]
    r = s.stop()
    assert r["samples"] == 2 and r["sm_mhz"] == pytest.approx(1957.5) and r["sm_max_mhz"] == 1965.0
    assert r["reasons"] == ["sw_power_cap"]


def test_seq2_pack_reconstructs_seq4(host_libs):
    """tc_seq2_pack (compact SEQ transport): expanding the 2-bit codes and patching the exception words gives seq4 back bit for
    bit — valid bases, zero padding, every non-ACGT code."""
    from trueconsense_b200 import synth
    from trueconsense_b200.reads import ReadBatch

    def rebuild(c, b):
        h = c.seq2.astype(np.uint32)
        v = np.zeros_like(b.seq4)
        nw = np.diff(b.seq_off.astype(np.int64))
        widx = np.arange(b.seq4.shape[0])
        rid = np.repeat(np.arange(b.n_reads), nw)
        valid = b.l_seq[rid] - 8 * (widx - b.seq_off[rid].astype(np.int64))
        for j in range(8):
            nib = (1 << ((h >> (2 * j)) & 3)).astype(np.uint32)
            v |= np.where(j < valid, nib << (8 * (j >> 1) + (0 if j & 1 else 4)), 0).astype(np.uint32)
        v[c.seq_exc_idx] = c.seq_exc_val
        return v

    w = synth.config(0, scale=0.05)
    b = synth.generate_reads(w.params, w.ref)
    c = b.with_seq2()
    assert c is not b and c.seq2.shape == b.seq4.shape and np.all(np.diff(c.seq_exc_idx.astype(np.int64)) > 0)
    assert np.array_equal(rebuild(c, b), b.seq4)
    odd = ReadBatch.from_records([dict(pos=3, cigar="16M", seq="=ACMGRSVTWYHKDBN"), dict(pos=4, cigar="5M", seq="ACGTN"),
                                  dict(pos=6, cigar="7M", seq="*"), dict(pos=7, cigar="1M", seq="T"), dict(pos=9, cigar="9M", seq="ACGTACGTA")])
    c = odd.with_seq2(max_exception_fraction=1.0)
    assert np.array_equal(rebuild(c, odd), odd.seq4)
    assert c.c_struct().n_seq_exc == c.seq_exc_idx.size > 0
    assert odd.with_seq2(max_exception_fraction=0.0) is odd
    from trueconsense_b200 import bamio

    g = bamio.read_bam(f"{GOLD}/mini_ont.bam", compact=True)
    assert g.seq2 is not None and g.cigar16 is not None and np.array_equal(rebuild(g, g), g.seq4)
