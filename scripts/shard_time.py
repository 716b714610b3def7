"""One GPU: tc_pileup_counts on single read-range shards of config 4 (what a rank of `bench.py --gpus N` piles up), to separate
the shard's own cost from multi-process effects.    python scripts/shard_time.py [world] [ranks...]"""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from trueconsense_b200 import gpu, sharding, synth  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ranks = [int(x) for x in sys.argv[2:]] or [0, world // 2, world - 1]
w = synth.config(3, scale=1.0)
n, L = int(w.params.n_reads), len(w.ref)
ctx = gpu.Context(0)
ctx.set_timing(True)
out = torch.empty((gpu.TC_NROWS, L), dtype=torch.int32, device="cuda")
p = gpu.buildindex_params()
p.max_depth = 0
for rank in ranks:
    lo, hi = sharding.read_range(n, rank, world)
    shard = synth.generate_reads(w.params, w.ref, read_range=(lo, hi))
    dev = ctx.upload(shard, with_qual=False)
    for _ in range(5):
        ctx.pileup_counts(dev, L, p, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k = []
    e0.record()
    for _ in range(20):
        ctx.pileup_counts(dev, L, p, out=out)
        k.append(ctx.last_pileup_kernel_ms())
    e1.record()
    torch.cuda.synchronize()
    print(f"world {world} rank {rank}: reads {hi - lo}, columns {int(shard.pos.min())}..{int(shard.pos.max())}, "
          f"tc_pileup_counts {e0.elapsed_time(e1) / 20:.3f} ms, pileup kernel {statistics.mean(k):.3f} ms", flush=True)
