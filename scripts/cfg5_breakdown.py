"""Kernel-by-kernel time of tc_pileup_counts on BASELINE config 5 (10-kb reads): run under
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_cfg5.csv python scripts/cfg5_breakdown.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trueconsense_b200 import gpu, synth  # noqa: E402

ctx = gpu.Context(0)
w = synth.config(4, scale=0.5)
b = synth.generate_reads(w.params, w.ref)
L = len(w.ref)
dev = ctx.upload(b, with_qual=False)
out = torch.empty((8, L), dtype=torch.int32, device="cuda")
p = gpu.buildindex_params(0)
for _ in range(3):
    ctx.pileup_counts(dev, L, p, out=out)
torch.cuda.synchronize()
print(w.name, b.n_reads, "reads", b.count_aligned_bases(0x4), "aligned bases")
