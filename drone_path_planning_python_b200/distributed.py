"""Multi-GPU plumbing: trajectories are independent, so the batch is block-partitioned by
trajectory index over the ranks (one process per GPU) and the only collective is the final
all-gather of coefficients and collision flags.  The gather runs chunk by chunk so
NCCL moves chunk c while the kernels of chunk c+1 run.  Everything here is backend agnostic
(``nccl`` on the GPU box, ``gloo`` in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, world: int, rank: int, group: int = 1) -> Tuple[int, int]:
    """Half-open range of trajectory indices owned by ``rank``: contiguous blocks, sizes as
    equal as possible in units of ``group`` trajectories (a formation's drones, which share
    one time vector, are never split across GPUs)."""
    if total % group != 0:
        raise ValueError("total must be a multiple of the time-sharing group size")
    units = total // group
    base, extra = divmod(units, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo * group, hi * group


def chunk_plan(count: int, chunks: int, group: int = 1) -> List[Tuple[int, int]]:
    """Split ``count`` local trajectories into at most ``chunks`` contiguous pieces (multiples
    of ``group``) for the compute/gather overlap."""
    units = count // group
    chunks = max(1, min(chunks, units if units > 0 else 1))
    out, lo = [], 0
    for c in range(chunks):
        n = units // chunks + (1 if c < units % chunks else 0)
        out.append((lo * group, (lo + n) * group))
        lo += n
    return [p for p in out if p[1] > p[0]]


class ChunkedAllGather:
    """All-gather of per-trajectory result tensors, chunked for overlap.

    Every rank must own the same number of trajectories (pad the batch if needed).
    ``run(compute_chunk)`` calls ``compute_chunk(lo, hi)`` for every local chunk — it must
    return the tuple of result tensors for trajectories ``[lo, hi)`` — issues the asynchronous
    all-gathers right behind it, and finally returns tensors in GLOBAL trajectory order:
    ``out[k][r * count + i]`` is trajectory ``i`` of rank ``r``.
    """

    def __init__(self, count: int, world: int, chunks: int, templates: Sequence[torch.Tensor], group: int = 1):
        self.count, self.world = count, world
        self.plan = chunk_plan(count, chunks, group)
        self.staging = []
        for lo, hi in self.plan:
            self.staging.append([torch.empty((world, hi - lo) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
                                 for t in templates])
        self.out = [torch.empty((world * count,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
                    for t in templates]

    def run(self, compute_chunk: Callable[[int, int], Sequence[torch.Tensor]], assemble: bool = True):
        handles = []
        for c, (lo, hi) in enumerate(self.plan):
            results = compute_chunk(lo, hi)
            for k, res in enumerate(results):
                stage = self.staging[c][k]
                # concatenated form ([world * n, ...]) is accepted by every backend
                handles.append(dist.all_gather_into_tensor(stage.view((-1,) + tuple(stage.shape[2:])),
                                                           res.contiguous(), async_op=True))
        for h in handles:
            h.wait()
        if not assemble:
            return self.staging
        for c, (lo, hi) in enumerate(self.plan):
            for k, full in enumerate(self.out):
                view = full.view((self.world, self.count) + tuple(full.shape[1:]))
                view[:, lo:hi].copy_(self.staging[c][k])
        return self.out
