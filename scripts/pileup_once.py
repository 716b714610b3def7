"""One device-resident tc_pileup_counts on a chosen BASELINE config (for ncu captures).
    python scripts/pileup_once.py <config-index 0..4> <scale> [kernel]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trueconsense_b200 import gpu, synth  # noqa: E402

idx, scale = int(sys.argv[1]), float(sys.argv[2])
kernel = int(sys.argv[3]) if len(sys.argv) > 3 else 0
ctx = gpu.Context(0)
w = synth.config(idx, scale=scale)
b = synth.generate_reads(w.params, w.ref)
L = len(w.ref)
dev = ctx.upload(b, with_qual=False)
out = torch.empty((8, L), dtype=torch.int32, device="cuda")
for _ in range(2):
    ctx.pileup_counts(dev, L, gpu.buildindex_params(kernel), out=out)
torch.cuda.synchronize()
print(w.name, b.n_reads, "reads", b.count_aligned_bases(0x4), "aligned bases")
