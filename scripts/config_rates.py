"""Pileup-kernel rates on BASELINE.json's other configs (device-resident reads, CUDA events), for DESIGN.md.
    python scripts/config_rates.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trueconsense_b200 import gpu, synth  # noqa: E402

ctx = gpu.Context(0)
KERNELS = [int(k) for k in sys.argv[1].split(',')] if len(sys.argv) > 1 else [0, 1]
for idx, scale in ((0, 1.0), (1, 1.0), (2, 1.0), (3, 0.1), (4, 0.5)):
    w = synth.config(idx, scale=scale)
    b = synth.generate_reads(w.params, w.ref)
    L = len(w.ref)
    dev = ctx.upload(b, with_qual=False)
    out = torch.empty((8, L), dtype=torch.int32, device="cuda")
    bases = b.count_aligned_bases(0x4)
    for kernel in KERNELS:
        if idx == 4 and kernel == 3:
            kernel = 4                           # long reads: pieces through variant 3
        p = gpu.buildindex_params(kernel)
        for _ in range(3):
            ctx.pileup_counts(dev, L, p, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ctx.pileup_counts(dev, L, p, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"{w.name} scale {scale}: {b.n_reads} reads, {bases} aligned bases, alg bytes {b.algorithmic_bytes(L)}, kernel={kernel}: "
              f"{ms:.3f} ms per tc_pileup_counts -> {bases / ms / 1e9:.3f} T bases/s, {b.algorithmic_bytes(L) / ms / 1e6:.1f} GB/s")
