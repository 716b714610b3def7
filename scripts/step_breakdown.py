"""Wall-clock breakdown of one device-resident hot-path step (config 2), per public call.
    python scripts/step_breakdown.py [scale]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trueconsense_b200 import gpu, synth  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
ctx = gpu.Context(0)
w = synth.config(1, scale=scale)
b = synth.generate_reads(w.params, w.ref)
L = len(w.ref)
dev = ctx.upload(b.pin())
stream = torch.cuda.current_stream().cuda_stream
counts = torch.empty((8, L), dtype=torch.int32, device="cuda")
cc = torch.empty(L, dtype=torch.uint8, device="cuda"); fl = torch.empty(L, dtype=torch.uint8, device="cuda")
xr = torch.empty(L, dtype=torch.int32, device="cuda"); rl = torch.empty((4, L), dtype=torch.uint8, device="cuda")
rc = torch.empty((4, L), dtype=torch.int32, device="cuda"); am = torch.empty(L, dtype=torch.uint8, device="cuda")
table = gpu.CallTable(cc.data_ptr(), fl.data_ptr(), xr.data_ptr(), rl.data_ptr(), rc.data_ptr(), am.data_ptr())
p = gpu.buildindex_params()
acc = {}
for it in range(13):
    t = [time.perf_counter()]
    ctx.pileup_counts(dev, L, p, out=counts, stream=stream); t.append(time.perf_counter())
    ctx.call_device(counts, L, w.mincov, True, table, stream=stream); t.append(time.perf_counter())
    cands = ctx.list_insert_candidates(fl, L); t.append(time.perf_counter())
    ins = ctx.extract_inserts(dev, L, cands); t.append(time.perf_counter())
    if it >= 3:
        for name, a0, a1 in zip(("pileup_counts", "call_device", "list_insert_candidates", "extract_inserts"), t, t[1:]):
            acc[name] = acc.get(name, 0.0) + (a1 - a0)
tot = sum(acc.values())
for k, v in acc.items():
    print(f"{k:24s} {1e3 * v / 10:8.3f} ms")
print(f"{'step':24s} {1e3 * tot / 10:8.3f} ms   (pileup kernel alone {ctx.last_pileup_kernel_ms():.3f} ms)")
