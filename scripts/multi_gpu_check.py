"""Read-range sharding across GPUs (torchrun, one rank per GPU): every rank piles up its slice, the tables are
summed with tc_allreduce_counts (NCCL) and must equal the table one GPU computes from all reads.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from trueconsense_b200 import gpu, sharding, synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = gpu.Context(local)
    comm = sharding.NcclComm(rank, world, local)
    w = synth.config(1, scale=float(os.environ.get("TC_SCALE", "0.25")))
    batch = synth.generate_reads(w.params, w.ref)
    L = len(w.ref)
    full = torch.empty((gpu.TC_NROWS, L), dtype=torch.int32, device="cuda")
    ctx.pileup_counts(batch, L, out=full)
    out = sharding.pileup_counts_read_range(ctx, batch, L, rank, world, comm)
    torch.cuda.synchronize()
    ok = bool(torch.equal(out, full))
    # timing of the sharded pass with the slice resident on the device
    lo, hi = sharding.read_range(batch.n_reads, rank, world)
    dev = ctx.upload(batch.slice(lo, hi), with_qual=False)
    p = gpu.buildindex_params()
    p.max_depth = 0
    for _ in range(3):
        ctx.pileup_counts(dev, L, p, out=out)
        ctx.allreduce_counts(out, comm)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ctx.pileup_counts(dev, L, p, out=out)
        ctx.allreduce_counts(out, comm)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 10], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    flags = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"read-range sharding over {world} GPUs: summed table == single-GPU table: {bool(flags.item())}; "
              f"{batch.n_reads} reads, pileup+allreduce {ms.item():.3f} ms per pass (max over ranks)")
    comm.close()
    dist.destroy_process_group()
    if not flags.item():
        sys.exit(1)


if __name__ == "__main__":
    main()
