"""Multi-GPU host logic on CPU: the partitions of sharding.py, and — with two gloo ranks — that read-range
shards summed by an all-reduce give exactly the single-process count table (the oracle stands in for the
per-rank pileup here; the GPU path does the same sum with tc_allreduce_counts over NCCL)."""
import os
import socket

import numpy as np
import pytest

from conftest import ROOT


def test_partitions_tile_the_input():
    from trueconsense_b200 import sharding

    for n in (0, 1, 7, 1000, 2_000_003):
        for world in (1, 2, 3, 8):
            cuts = [sharding.read_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            assert max(hi - lo for lo, hi in cuts) - min(hi - lo for lo, hi in cuts) <= 1
    assert sorted(sum((sharding.samples_of_rank(96, r, 8) for r in range(8)), [])) == list(range(96))
    assert sharding.samples_of_rank(5, 3, 4) == [3]
    with pytest.raises(ValueError):
        sharding.read_range(10, 2, 2)


def test_depth_cap_is_checked_on_the_summed_coverage():
    from trueconsense_b200 import gpu, sharding

    sharding.check_depth_cap(4_000_000, 0, 10_000_000)
    with pytest.raises(gpu.TcError) as ei:
        sharding.check_depth_cap(5_000_000, 0, 10_000_000)
    assert ei.value.code == -4


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, out_dir):
    import sys

    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from oracle import pileup
    from trueconsense_b200 import sharding, synth

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = synth.config(1, scale=0.002)            # the amplicon workload: deep columns, indels, insertions
        batch = synth.generate_reads(w.params, w.ref)
        L = len(w.ref)
        lo, hi = sharding.read_range(batch.n_reads, rank, world)
        part = torch.from_numpy(pileup.pileup_counts(batch.slice(lo, hi), L, threads=2).astype(np.int32))
        dist.all_reduce(part, op=dist.ReduceOp.SUM)
        full = pileup.pileup_counts(batch, L, threads=2)
        ok = np.array_equal(part.numpy(), full)
        # every rank holds the same summed table
        gathered = [torch.zeros_like(part) for _ in range(world)]
        dist.all_gather(gathered, part)
        same = all(torch.equal(g, part) for g in gathered)
        sharding.check_depth_cap(int(part[0].max()), 0, 10_000_000)
        with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as fh:
            fh.write(f"{int(ok)} {int(same)} {lo} {hi} {batch.n_reads}")
    finally:
        dist.destroy_process_group()


def test_read_range_shards_sum_to_the_full_table_gloo(tmp_path, host_libs):
    import torch.multiprocessing as mp

    world = 2
    mp.spawn(_rank_main, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    cuts = []
    for r in range(world):
        ok, same, lo, hi, n = (int(x) for x in open(tmp_path / f"rank{r}.txt").read().split())
        assert ok == 1 and same == 1
        cuts.append((lo, hi, n))
    assert cuts[0][0] == 0 and cuts[0][1] == cuts[1][0] and cuts[1][1] == cuts[1][2]
