"""The oracle against the golden vectors generated from the reference (oracle/make_golden.py).

CPU only.  This is what pins oracle/call.py (ranking, IsAmbiguous, MinorityDel, ListInserts,
ExtractInserts, the consensus walk, CorrectGFF) and the classifier half of
oracle/pileup_oracle.c to the reference's behaviour.
"""
import numpy as np
import pytest

from conftest import GOLD, load_golden_counts, load_golden_json
from oracle import bam_py, call, pileup

KAT = load_golden_json("kat.json")
MINIS = ("quirk", "mini_illumina", "mini_ont", "mini_long", "mini_overlap")


def _col_counts(c):
    counts = np.zeros((8, 1), dtype=np.int64)
    for k, r in call.ROW.items():
        counts[r, 0] = c[k]
    return counts


def test_ranking_kat():
    for case in KAT["ranking"]:
        counts = _col_counts(case["col"])
        got = [list(call.get_nucleotide(counts, 1, k)) for k in range(1, 6)]
        assert got == case["ranks"], case


def test_is_ambiguous_kat():
    n_true = 0
    for case in KAT["is_ambiguous"]:
        ranks = [tuple(t) for t in case["ranks"]]
        if case["status"] == "raise":
            with pytest.raises(Exception) as ei:
                call.is_ambiguous(*ranks, case["cov"])
            assert type(ei.value).__name__ == case["result"]
            continue
        try:
            got = list(call.is_ambiguous(*ranks, case["cov"]))
        except KeyError:
            # duplicate / X-in-pair letter sets make the reference fall through with an
            # unbound local; those inputs cannot come from a ranking and are not part of the contract
            continue
        assert got == case["result"], case
        n_true += bool(got[0])
    assert n_true > 20


def test_ieee_threshold_quirks():
    # these cannot be reproduced with integer cross-multiplication
    assert call.is_ambiguous(("A", 55), ("C", 45), ("T", 0), ("G", 0), 100) == (False, None)
    assert call.is_ambiguous(("A", 6), ("C", 5), ("T", 0), ("G", 0), 10) == (True, "M")
    assert call.insert_candidates(_col_counts(dict(coverage=100, A=100, T=0, C=0, G=0, X=0, I=55)), 30) == [1]


def test_minority_del_kat():
    for case in KAT["minority_del"]:
        counts = _col_counts(dict(coverage=case["cov"], A=0, T=0, C=0, G=0, X=case["X"], I=0))
        if case["status"] == "raise":
            with pytest.raises(ZeroDivisionError):
                call.minority_del(counts, 1)
        else:
            assert call.minority_del(counts, 1) == case["result"], case


def test_list_inserts_kat():
    for case in KAT["list_inserts"]:
        counts = _col_counts(dict(coverage=case["cov"], A=case["cov"], T=0, C=0, G=0, X=0, I=case["I"]))
        has, pos = call.list_inserts(counts, case["mincov"], lambda p: ["A+3TTT"] * 6 + ["A"] * 4)
        exp_has, exp_pos = case["result"]
        assert has == exp_has
        assert (pos if pos is None else {str(k): v for k, v in pos.items()}) == exp_pos


def test_extract_inserts_kat():
    for case in KAT["extract_inserts"]:
        assert list(call.extract_insert(case["strings"])) == case["result"], case


def test_walk_cases():
    cases = load_golden_json("walk_cases.json")
    n_ok = n_raise = 0
    for c in cases:
        counts = np.zeros((8, c["L"]), dtype=np.int64)
        counts[:7] = np.array(c["counts"], dtype=np.int64)
        gff = {int(k): v for k, v in c["gff"].items()}
        cols = {int(k): v for k, v in c["columns"].items()}

        def column_strings(p):
            return cols.get(p - 1)

        for inc in (True, False):
            exp = c[f"ins{int(inc)}"]
            inserts = call.list_inserts(counts, c["mincov"], column_strings)
            if exp["status"] == "raise":
                with pytest.raises(Exception) as ei:
                    call.build_consensus(c["mincov"], counts, gff, c["include_ambig"], inserts, inc)
                assert type(ei.value).__name__ == exp["exc"]
                n_raise += 1
            else:
                cons, newgff = call.build_consensus(c["mincov"], counts, gff, c["include_ambig"], inserts, inc)
                assert cons == exp["consensus"]
                assert {str(k): [v["start"], v["end"]] for k, v in newgff.items()} == exp["gff"]
                n_ok += 1
    assert n_ok > 300 and n_raise > 50


@pytest.mark.parametrize("name", MINIS)
def test_pileup_counts_golden(name, host_libs):
    b = bam_py.read_bam(f"{GOLD}/{name}.bam")
    exp = load_golden_counts(name)
    got = pileup.pileup_counts(b, exp.shape[1])
    assert np.array_equal(got[:7], exp)
    got4 = pileup.pileup_counts(b, exp.shape[1], threads=4)
    assert np.array_equal(got4, got)


@pytest.mark.parametrize("name", MINIS)
def test_inserts_and_consensus_golden(name, host_libs):
    meta = load_golden_json(f"{name}.json")
    b = bam_py.read_bam(f"{GOLD}/{name}.bam")
    counts = load_golden_counts(name).astype(np.int64)
    counts = np.vstack([counts, np.zeros((1, counts.shape[1]), np.int64)])

    def column_strings(p):
        cols = pileup.pileup_columns(b, region=(p - 1, p), **pileup.EXTRACTINSERTS)
        return cols[0][1] if cols else None

    has, pos = call.list_inserts(counts, meta["mincov"], column_strings)
    assert [has, None if pos is None else {str(k): v for k, v in pos.items()}] == meta["list_inserts"]
    if counts.shape[1] > 4000:
        return      # the plain O(ORF^2) walk is only run on the small genomes
    from trueconsense_b200 import synth  # noqa: F401  (GFF fixtures are plain text)
    gff = {}
    with open(f"{GOLD}/{name}.gff") as fh:
        for line in fh:
            if line.startswith("#"):
                continue
            f = line.rstrip("\n").split("\t")
            gff[len(gff)] = {"seqid": name, "source": f[1], "type": f[2], "start": int(f[3]), "end": int(f[4]),
                             "score": f[5], "strand": f[6], "phase": f[7], "attributes": f[8]}
    for amb in (True, False):
        for inc in (True, False):
            exp = meta[f"consensus_amb{int(amb)}_ins{int(inc)}"]
            if exp["status"] == "raise":
                with pytest.raises(Exception) as ei:
                    call.build_consensus(meta["mincov"], counts, gff, amb, (has, pos), inc)
                assert type(ei.value).__name__ == exp["exc"]
            else:
                cons, newgff = call.build_consensus(meta["mincov"], counts, gff, amb, (has, pos), inc)
                assert cons == exp["consensus"]
                assert {str(k): [v["start"], v["end"]] for k, v in newgff.items()} == exp["gff"]


def test_call_table_consistency():
    """call_table's flags/xrun restate what the walk computes on the fly."""
    rng = np.random.default_rng(5)
    from oracle import fixtures

    for _ in range(50):
        L = int(rng.integers(5, 80))
        counts = np.zeros((8, L), np.int64)
        counts[:8] = fixtures.random_counts(rng, L, 50)
        t = call.call_table(counts, 10, True)
        for p in range(1, L + 1):
            try:
                run = len(call.walk_forward(counts, p))
                assert not (t["flags"][p - 1] & call.CF_XRUN_OFF_END)
                assert t["xrun"][p - 1] == run
            except KeyError:
                assert t["flags"][p - 1] & call.CF_XRUN_OFF_END
            assert bool(t["flags"][p - 1] & call.CF_PRIMARY_X) == (call.get_nucleotide(counts, p, 1)[0] == "X")


def test_mate_overlap_rewrite_kat():
    """htslib's mate-overlap quality rewriting (pysam ignore_overlaps=True, the default of Events.py:66) on hand-derived
    cases: the oracle's ExtractInserts column must hold exactly the strings htslib's rule leaves (oracle/fixtures.py)."""
    from collections import Counter

    from oracle import fixtures
    from trueconsense_b200.reads import ReadBatch

    recs, exp = fixtures.overlap_kat_records()
    b = ReadBatch.from_records(recs)
    c = fixtures.OVERLAP_COL
    cols = pileup.pileup_columns(b, region=(c, c + 1), **pileup.EXTRACTINSERTS)
    assert len(cols) == 1 and cols[0][0] == c
    assert Counter(cols[0][1]) == Counter(exp)
    # ignore_overlaps=False: every mate of a proper pair whose own quality passes is there
    off = pileup.pileup_columns(b, region=(c, c + 1), reserved=1 << 8, **pileup.EXTRACTINSERTS)
    assert len(off[0][1]) > len(exp)
    # the rule of htslib <= 1.12: the first-arrived mate always keeps (agreement: the sum; disagreement: the better, a on ties)
    old = Counter(pileup.pileup_columns(b, region=(c, c + 1), reserved=2 << 8, **pileup.EXTRACTINSERTS)[0][1])
    # solo reads, mate_far, mate a of agree_low / agree_high / dis_a_better / dis_tie; mate b only where it is the better base
    assert old["a+2tt"] == 0 and old["A+2TT"] == 12 + 1 + 4 and old["c+2tt"] == 1
    # BuildIndex (min_base_quality = 0) cannot see any of it
    full = pileup.pileup_counts(b, fixtures.OVERLAP_REF_LEN)
    assert full[0, c] == len(recs) - 1          # every read but mate_far's second mate covers the column (orphans included: nofilter)
