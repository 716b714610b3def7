"""Ad-hoc stress of tc_pileup_counts against the CPU oracle (TEST AID, not part of the product): many seeds of the
synthetic generator at sizes that exercise sub-tile tails, window moves, all three geometries and the long-read path.
    python scripts/stress_pileup.py [n_seeds]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pileup  # noqa: E402
from trueconsense_b200 import gpu, synth  # noqa: E402

pileup.build()
ctx = gpu.Context(0)
n_seeds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
rng = np.random.default_rng(99)
bad = 0
for seed in range(n_seeds):
    L = int(rng.integers(7000, 12000))
    kind = seed % 5
    kw = [dict(read_len=150, read_len_jitter=int(rng.integers(0, 60)), paired=bool(seed & 1)),
          dict(read_len=400, read_len_jitter=int(rng.integers(0, 40)), n_amplicons=int(rng.integers(1, 6)), amplicon_jitter=int(rng.integers(0, 12)), indel_rate=1 / 30),
          dict(read_len=int(rng.integers(500, 900)), read_len_jitter=80, indel_rate=1 / 40, indel_maxlen=3),
          dict(read_len=int(rng.integers(1500, 5000)), read_len_jitter=500, indel_rate=1 / 40, indel_maxlen=2),
          dict(read_len=int(rng.integers(30, 120)), read_len_jitter=20, indel_rate=0.05, indel_maxlen=4, refskip_rate=0.05)][kind]
    n_reads = int(rng.integers(1, 40000 if kind != 3 else 1500))
    p = synth.SynthParams(seed=1000 + seed, n_reads=n_reads, ref_len=L, softclip_rate=0.2, softclip_max=int(rng.integers(1, 40)),
                          n_rate=0.003, iupac_rate=0.002, sub_rate=0.02, special_flag_rate=0.03, **kw)
    ref, _ = synth.make_genome(L, seed, "sars2")
    b = synth.generate_reads(p, ref)
    exp = pileup.pileup_counts(b, L, threads=8)
    span = int(b.max_ref_span)
    for hint in (span if span > 0 else -1, -1):          # with the generator's span bound (folded span pass), then without
        b.max_ref_span = hint
        got = ctx.pileup_counts(b, L, gpu.buildindex_params(0))
        ok = np.array_equal(got, exp)
        bad += not ok
        print(f"seed {seed:3d} kind {kind} reads {b.n_reads:6d} L {L} span bound {hint}: {'ok' if ok else 'MISMATCH'}")
print("mismatches:", bad)
sys.exit(1 if bad else 0)
