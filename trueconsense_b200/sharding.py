"""Multi-GPU sharding of the pileup (nothing like it exists in the reference).

Two ways to spread the path over the GPUs of one box, one process per GPU:

* **by sample** — a multi-sample batch (the 96-sample plate): every rank takes whole samples
  (``samples_of_rank``); there is no data-path collective at all.
* **by read range** — one ultra-deep sample: every rank piles up a contiguous slice of the start-sorted reads
  into a full-length count table, and the tables are summed with ONE ``ncclAllReduce(int32, sum)`` over
  NVLink (``tc_allreduce_counts``).  Counts are additive integers, so the sum is bit-identical to the
  single-GPU table; the call kernel then runs replicated.  What is order dependent in htslib (the
  ``max_depth`` push rule) is proven non-binding on the SUMMED coverage, like on one GPU.

``NcclComm`` creates the communicator the C-ABI needs straight from libnccl (ctypes): the unique id is made
on rank 0 and broadcast through ``torch.distributed`` (any backend), so the same code runs under
``torchrun`` next to torch's own process group.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import gpu
from .reads import ReadBatch


def read_range(n_reads: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slice [lo, hi) of the start-sorted reads owned by ``rank``; slices tile 0..n_reads."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside 0..{world - 1}")
    return rank * n_reads // world, (rank + 1) * n_reads // world


def samples_of_rank(n_samples: int, rank: int, world: int) -> list[int]:
    """Round-robin assignment of whole samples to ranks."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside 0..{world - 1}")
    return list(range(rank, n_samples, world))


def check_depth_cap(coverage_max: int, n_zero_span: int, max_depth: int) -> None:
    """The bulk pileup does not emulate htslib's order-dependent depth cap; it must be provably non-binding on
    the whole sample (same bound tc_pileup_counts applies per call: pileup.cu status_to_rc)."""
    if max_depth > 0 and 2 * int(coverage_max) + int(n_zero_span) + 1 > max_depth:
        raise gpu.TcError(-4, f"coverage {coverage_max} could reach max_depth {max_depth}: the pileup depth cap may bind "
                              "and is not emulated by the bulk kernel")


def _nccl_path() -> str:
    p = os.environ.get("TC_NCCL_LIB")
    if p:
        return p
    try:
        import torch

        cand = os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "nccl", "lib", "libnccl.so.2")
        if os.path.exists(cand):
            return os.path.abspath(cand)
    except Exception:
        pass
    return "libnccl.so.2"


class _UniqueId(C.Structure):
    _fields_ = [("internal", C.c_byte * 128)]


class NcclComm:
    """An ``ncclComm_t`` of this process for (rank, world) — what ``tc_allreduce_counts`` takes."""

    def __init__(self, rank: int, world: int, device: int):
        import torch
        import torch.distributed as dist

        path = _nccl_path()
        os.environ.setdefault("TC_NCCL_LIB", path)          # the C side resolves ncclAllReduce from the same library
        self._lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
        self._lib.ncclGetErrorString.restype = C.c_char_p
        uid = _UniqueId()
        if rank == 0:
            self._check(self._lib.ncclGetUniqueId(C.byref(uid)))
        if world > 1:
            use_cuda = dist.get_backend() == "nccl"
            t = torch.frombuffer(bytearray(bytes(uid.internal)), dtype=torch.uint8).clone()
            if use_cuda:
                t = t.cuda(device)
            dist.broadcast(t, src=0)
            C.memmove(C.byref(uid), bytes(t.cpu().numpy().tobytes()), 128)
        torch.cuda.set_device(device)
        self.handle = C.c_void_p()
        self._check(self._lib.ncclCommInitRank(C.byref(self.handle), C.c_int(world), uid, C.c_int(rank)))
        self.rank, self.world = rank, world

    def _check(self, rc: int) -> None:
        if rc != 0:
            raise RuntimeError("NCCL: " + self._lib.ncclGetErrorString(rc).decode())

    def close(self) -> None:
        if getattr(self, "handle", None):
            self._lib.ncclCommDestroy(self.handle)
            self.handle = None


def pileup_counts_read_range(ctx: "gpu.Context", batch: ReadBatch, ref_len: int, rank: int, world: int, comm: NcclComm | None,
                             out=None, stream: int = 0, params: "gpu.PileupParams | None" = None):
    """Count table of ``batch`` with the reads sharded by range over ``world`` ranks.  ``out``: a torch CUDA
    int32 tensor [8][ref_len] (created if None).  Every rank returns the full, summed table."""
    import torch

    params = params or gpu.buildindex_params()
    if not batch.sorted:
        raise ValueError("Unsorted input. Pileup aborts")
    lo, hi = read_range(batch.n_reads, rank, world)
    if out is None:
        out = torch.empty((gpu.TC_NROWS, ref_len), dtype=torch.int32, device="cuda")
    shard_params = gpu.PileupParams.from_buffer_copy(params)
    shard_params.max_depth = 0                  # the cap is checked on the summed coverage below, not per shard
    if world > 1:
        if comm is None:
            raise ValueError("a communicator is needed to sum the per-rank tables")
        ctx.pileup_counts_allreduce(batch.slice(lo, hi), ref_len, shard_params, out, comm, stream=stream)
    else:
        ctx.pileup_counts(batch.slice(lo, hi), ref_len, shard_params, out=out, stream=stream)
    cov = out[0]
    zero_span = int(np.count_nonzero(batch.ref_spans() == 0)) if params.max_depth and params.max_depth < 4 * batch.n_reads else 0
    check_depth_cap(int(cov.max().item()) if ref_len else 0, zero_span, int(params.max_depth))
    return out
