"""Generate tests/golden/* by running the UNMODIFIED reference (/root/reference) in this container
(TEST INFRASTRUCTURE ONLY — see oracle/__init__.py).

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden

What is pinned, and by what:
* kat.json          — the reference's own per-position functions (Sequences.GetNucleotide,
                      Ambig.IsAmbiguous, Events.MinorityDel / ListInserts / ExtractInserts,
                      indexing.BuildIndex's classifier) on the known-answer inputs of SURVEY.md §4.3.
* walk_cases.json   — Sequences.BuildConsensus (+ ORFs.CorrectGFF) on seeded random count tables,
                      GFFs and insertion columns: consensus, corrected GFF coordinates or the
                      exception type the reference raises.
* mini_*.{bam,fasta,gff,npz,json}, quirk.* — small BAMs; expected count table = the reference's
                      indexing.BuildIndex run on a stand-in ``pysam`` backed by the restated htslib
                      engine (oracle/pileup_oracle.c).  These pin the reference's classifier,
                      insertion caller, walk and writers; they do NOT pin htslib itself, whose
                      source is not in this container (parity unpinned at that boundary).
* cli_*             — the reference's CLI main() run end to end on the mini BAMs (FASTA, VCF, GFF,
                      coverage TSV).
"""
from __future__ import annotations

import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _exc_name(fn):
    try:
        return ("ok", fn())
    except Exception as e:       # noqa: BLE001 - the exception type is the datum
        return ("raise", type(e).__name__)


def make_kat(ref) -> dict:
    S, A, E = ref.Sequences, ref.Ambig, ref.Events
    out = {}

    def col(cov=0, A_=0, T=0, C=0, G=0, X=0, I=0):
        return {"coverage": cov, "A": A_, "T": T, "C": C, "G": G, "X": X, "I": I}

    rank_inputs = [col(40, 10, 10, 10, 10, 0), col(0), col(100, 50, 0, 0, 0, 50), col(9, 1, 2, 3, 2, 1),
                   col(7, 0, 0, 3, 3, 1), col(5, 5, 0, 0, 0, 0), col(20, 0, 5, 5, 5, 5)]
    out["ranking"] = [{"col": c, "ranks": [list(S.GetNucleotide({1: c}, 1, k)) for k in range(1, 6)]} for c in rank_inputs]

    amb_inputs = []
    for cov, cs in [(100, (55, 45, 0, 0)), (1000, (550, 450, 0, 0)), (300, (165, 135, 0, 0)), (10, (6, 5, 0, 0)),
                    (30, (18, 15, 0, 0)), (70, (42, 35, 0, 0)), (90, (50, 41, 0, 0)), (110, (61, 50, 0, 0)),
                    (130, (72, 59, 0, 0)), (100, (50, 50, 0, 0)), (100, (60, 40, 0, 0)), (99, (33, 33, 33, 0)),
                    (100, (25, 25, 25, 25)), (0, (0, 0, 0, 0)), (1000, (5, 3, 0, 0)), (100, (40, 31, 29, 0)),
                    (100, (34, 33, 33, 0)), (50, (20, 15, 10, 5)), (7, (3, 2, 1, 1)), (3, (1, 1, 1, 0))]:
        for letters in ("ACGT", "ATCG", "CGTA", "GTAC", "TGCA", "ACXG", "AXCG", "XACG", "ATGX", "CGXA", "TCAG", "GATC"):
            tup = [[letters[i], cs[i]] for i in range(4)]
            res = _exc_name(lambda: A.IsAmbiguous(*[tuple(t) for t in tup], cov))
            amb_inputs.append({"ranks": tup, "cov": cov, "result": list(res[1]) if res[0] == "ok" else res[1], "status": res[0]})
    out["is_ambiguous"] = amb_inputs

    md = []
    for cov, x in [(100, 15), (100, 14), (20, 3), (20, 2), (7, 1), (7, 2), (1000, 150), (1000, 149), (13, 2), (0, 0), (200, 30)]:
        res = _exc_name(lambda: E.MinorityDel({1: col(cov, X=x)}, 1))
        md.append({"cov": cov, "X": x, "status": res[0], "result": res[1]})
    out["minority_del"] = md

    from .ref_stubs import FakeBam

    li = []
    for cov, ins, mincov in [(100, 55, 30), (100, 56, 30), (100, 54, 30), (20, 12, 11), (20, 11, 11), (20, 12, 21),
                             (1000, 550, 30), (1000, 551, 30), (40, 22, 30), (40, 23, 30), (0, 0, 0), (9, 5, 0), (200, 110, 30)]:
        bam = FakeBam({0: ["A+3TTT"] * 6 + ["A"] * 4})
        res = E.ListInserts({1: col(cov, A_=cov, I=ins)}, mincov, bam)
        li.append({"cov": cov, "I": ins, "mincov": mincov, "result": [res[0], res[1] if res[1] is None else {str(k): v for k, v in res[1].items()}]})
    out["list_inserts"] = li

    ei = []
    for strings in [["A+3TTT"] * 60 + ["A"] * 40, ["a+12acgtacgtacgt"] * 60 + ["A"] * 40,
                    ["A-2NN"] * 30 + ["A+1T"] * 20 + ["A+1G"] * 20 + ["A+1C"] * 20 + ["A"] * 10,
                    ["A+1T", "A+1G", "a+1g", "A+1T"], "", ["A"] * 5, ["*+2AC"] * 3 + ["*"] * 2, ["A+2.C"] * 3,
                    [">-2NN", "<-2nn", "A"], ["A+10ACGTACGTAC"] * 2 + ["A+1A"], ["T+3ACG", "t+3acg", "T+3ACC", "T+3ACC"],
                    ["G+2A.", "G+2A."], ["C+1N"] * 2]:
        bam = FakeBam({4: strings})
        ei.append({"strings": strings, "result": list(E.ExtractInserts(bam, 5))})
    ei.append({"strings": None, "result": list(E.ExtractInserts(FakeBam({}), 5))})
    out["extract_inserts"] = ei
    return out


def make_walk_cases(ref, n_cases=600, seed=20260118) -> list:
    from . import fixtures
    from .ref_stubs import FakeBam

    rng = np.random.default_rng(seed)
    S = ref.Sequences
    cases = []
    for _ in range(n_cases):
        c = fixtures.random_walk_case(rng)
        L = c["L"]
        counts = np.array(c["counts"], dtype=np.int64)
        entry = dict(c)
        entry["columns"] = {str(k): v for k, v in c["columns"].items()}
        entry["gff"] = {str(k): v for k, v in c["gff"].items()}
        for inc in (True, False):
            index = {p: {"coverage": int(counts[0, p - 1]), "A": int(counts[1, p - 1]), "T": int(counts[2, p - 1]),
                         "C": int(counts[3, p - 1]), "G": int(counts[4, p - 1]), "X": int(counts[5, p - 1]),
                         "I": int(counts[6, p - 1])} for p in range(1, L + 1)}
            bam = FakeBam(dict(c["columns"]))
            import copy
            gff = copy.deepcopy(c["gff"])
            res = _exc_name(lambda: S.BuildConsensus(c["mincov"], index, gff, c["include_ambig"], bam, inc))
            if res[0] == "ok":
                cons, newgff = res[1]
                entry[f"ins{int(inc)}"] = {"status": "ok", "consensus": cons,
                                           "gff": {str(k): [v["start"], v["end"]] for k, v in newgff.items()}}
            else:
                entry[f"ins{int(inc)}"] = {"status": "raise", "exc": res[1]}
        cases.append(entry)
    return cases


def make_bam_fixture(ref, name: str, batch, refseq: str, feats: list[dict], mincov: int) -> None:
    from trueconsense_b200 import bamio, synth

    bam = os.path.join(GOLD, f"{name}.bam")
    fa = os.path.join(GOLD, f"{name}.fasta")
    gff = os.path.join(GOLD, f"{name}.gff")
    bamio.write_bam(bam, batch, "ref", len(refseq), level=6)
    synth.write_fasta(fa, "ref", refseq)
    synth.write_gff(gff, "ref", len(refseq), feats)
    df = ref.indexing.BuildIndex(bam, fa)
    cols = ["coverage", "A", "T", "C", "G", "X", "I"]
    assert list(df.columns) == cols and df.index[0] == 1 and len(df) == len(refseq)
    table = np.stack([df[c].to_numpy() for c in cols]).astype(np.int32)
    index = df.to_dict("index")
    bamobj = ref.indexing.Readbam(bam)
    has, pos = ref.Events.ListInserts(index, mincov, bamobj)
    gffobj = ref.indexing.Gffindex(gff)
    gdf = gffobj.df
    gdf["seqid"] = name
    gdict = gdf.to_dict("index")
    meta = {"name": name, "mincov": mincov, "ref_len": len(refseq), "n_reads": int(batch.n_reads),
            "list_inserts": [has, None if pos is None else {str(k): v for k, v in pos.items()}],
            "insert_pileup_calls": [list(c) for c in bamobj.pileup_calls]}
    for amb in (True, False):
        for inc in (True, False):
            res = _exc_name(lambda: ref.Sequences.BuildConsensus(mincov, df.to_dict("index"), gdict, amb, bamobj, inc))
            key = f"consensus_amb{int(amb)}_ins{int(inc)}"
            if res[0] == "ok":
                meta[key] = {"status": "ok", "consensus": res[1][0],
                             "gff": {str(k): [v["start"], v["end"]] for k, v in res[1][1].items()}}
            else:
                meta[key] = {"status": "raise", "exc": res[1]}
    np.savez_compressed(os.path.join(GOLD, f"{name}.npz"), counts=table)
    with open(os.path.join(GOLD, f"{name}.json"), "w") as fh:
        json.dump(meta, fh, indent=1, sort_keys=True)
    # the CLI, end to end
    out = os.path.join(GOLD, f"cli_{name}")
    argv = ["--input", bam, "--reference", fa, "--features", gff, "--coverage-level", str(mincov), "--samplename", name,
            "--output", out + ".fasta", "--variants", out + ".vcf", "--output-gff", out + ".gff",
            "--depth-of-coverage", out + ".cov.tsv", "--threads", "2"]
    old_argv = sys.argv
    sys.argv = ["TrueConsense"] + ["<golden>"]
    try:
        res = _exc_name(lambda: ref.TrueConsense.main(argv))
    finally:
        sys.argv = old_argv
    with open(out + ".status.json", "w") as fh:
        json.dump({"status": res[0], "exc": res[1] if res[0] == "raise" else None}, fh)
    if os.path.exists(out + ".vcf"):       # strip the run date so the file is reproducible
        lines = open(out + ".vcf").read().splitlines(keepends=True)
        lines = [("##fileDate=<date>\n" if l.startswith("##fileDate=") else l) for l in lines]
        lines = [l.replace(GOLD + os.sep, "") for l in lines]
        open(out + ".vcf", "w").writelines(lines)


def main() -> None:
    warnings.filterwarnings("ignore")
    sys.dont_write_bytecode = True
    sys.path.insert(0, ROOT)
    from . import fixtures, ref_stubs

    ref = ref_stubs.load_reference()
    os.makedirs(GOLD, exist_ok=True)
    with open(os.path.join(GOLD, "kat.json"), "w") as fh:
        json.dump(make_kat(ref), fh, indent=1)
    with open(os.path.join(GOLD, "walk_cases.json"), "w") as fh:
        json.dump(make_walk_cases(ref), fh)
    qb = fixtures.quirk_batch()
    qref = "".join("ACGT"[i % 4] for i in range(fixtures.QUIRK_REF_LEN))
    make_bam_fixture(ref, "quirk", qb, qref, [{"name": "q", "start": 4, "end": 63, "strand": "+"}], 1)
    for name in fixtures.MINI_NAMES:
        w, batch = fixtures.mini_workload(name)
        make_bam_fixture(ref, w.name, batch, w.ref, w.feats, w.mincov)
    name, oref, ofeats, omincov, ob = fixtures.overlap_workload()
    make_bam_fixture(ref, name, ob, oref, ofeats, omincov)
    print("golden vectors written to", GOLD)


if __name__ == "__main__":
    main()
