"""The reference's OWN CPU path, timed in this container: the unmodified /root/reference package —
indexing.BuildIndex (pysam pileup -> get_query_sequences -> parse_query_sequences, indexing.py:96-143) and
Sequences.BuildConsensus (Sequences.py:168-322) — on stand-in pysam / AminoExtract modules (oracle/ref_stubs.py; the pileup
strings come from oracle/pileup_oracle.c, the classifier, the frame building and the walk are the reference's own Python).

    PYTHONDONTWRITEBYTECODE=1 python scripts/reference_python_rate.py [reads] > profiles/r2_reference_python_rate.json

/root/reference does not exist on the GPU box, so this number cannot be taken inside bench.py; it is recorded here, next to
the C port's rate on the same prefix (kind "port": what bench.py's cpu_baseline and --impl reference report).
"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True

from oracle import pileup, ref_stubs  # noqa: E402
from trueconsense_b200 import bamio, synth  # noqa: E402


def main():
    n_reads = int(float(sys.argv[1])) if len(sys.argv) > 1 else 40_000
    ref = ref_stubs.load_reference()
    pileup.build()
    out = {"host": {"cpus": os.cpu_count(), "note": "CPU container (no GPU); one core: the reference's hot loops are single-threaded"}}
    for idx, name in ((0, "configs[0]"), (1, "configs[1] prefix")):
        w = synth.config(idx, scale=1.0 if idx == 0 else n_reads / 2_000_000)
        b = synth.generate_reads(w.params, w.ref)
        L = len(w.ref)
        bases = b.count_aligned_bases(0x4)
        with tempfile.TemporaryDirectory() as tmp:
            bam, fa, gff = (os.path.join(tmp, f"s.{e}") for e in ("bam", "fasta", "gff"))
            bamio.write_bam(bam, b, "ref", L, level=1)
            synth.write_fasta(fa, "ref", w.ref)
            synth.write_gff(gff, "ref", L, w.feats)
            t0 = time.perf_counter()
            df = ref.indexing.BuildIndex(bam, fa)
            t_index = time.perf_counter() - t0
            index = df.to_dict("index")
            gdf = ref.indexing.Gffindex(gff).df
            gdf["seqid"] = "s"
            bamobj = ref.indexing.Readbam(bam)
            t0 = time.perf_counter()
            try:
                ref.Sequences.BuildConsensus(w.mincov, index, gdf.to_dict("index"), True, bamobj, True)
                walk = "ok"
            except Exception as e:      # noqa: BLE001 — the reference raises on some inputs; time it all the same
                walk = type(e).__name__
            t_walk = time.perf_counter() - t0
        t0 = time.perf_counter()
        pileup.pileup_counts(b, L, threads=1)
        t_port = time.perf_counter() - t0
        out[name] = {
            "workload": w.name, "reads": int(b.n_reads), "aligned_bases": int(bases),
            "reference_BuildIndex_s": t_index, "reference_BuildIndex_aligned_bases_per_s": bases / t_index,
            "reference_BuildConsensus_s": t_walk, "reference_BuildConsensus": walk,
            "reference_path_aligned_bases_per_s": bases / (t_index + t_walk),
            "port_pileup_counts_s": t_port, "port_aligned_bases_per_s": bases / t_port,
            "kind": {"reference_*": "reference (its own Python on the stand-in pysam)", "port_*": "port (oracle/pileup_oracle.c, 1 thread)"},
        }
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
