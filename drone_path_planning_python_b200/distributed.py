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


def chunk_plan(count: int, chunks: int, group: int = 1, taper=False) -> List[Tuple[int, int]]:
    """Split ``count`` local trajectories into at most ``chunks`` contiguous pieces (multiples
    of ``group``) for the compute/gather overlap.  ``taper`` = True / ``"tail"``: every piece half the
    size of the one before it (8 : 4 : 2 : 1 ...), so that the push of the LAST piece — the only one
    no later compute hides — is short (compute-bound steps).  ``"head"``: the mirror image
    (1 : 2 : 4 ...), so that pushing starts early (exchange-bound steps: the step is then the first
    piece's compute plus all the pushes)."""
    units = count // group
    chunks = max(1, min(chunks, units if units > 0 else 1))
    if taper:
        weights = [2 ** (chunks - 1 - c) for c in range(chunks)]
        if taper == "head":
            weights.reverse()
        total, sizes, used = sum(weights), [], 0
        for c, w in enumerate(weights):
            n = units - used if c == chunks - 1 else max(1, (units * w) // total) if units - used > 0 else 0
            n = min(n, units - used)
            sizes.append(n)
            used += n
    else:
        sizes = [units // chunks + (1 if c < units % chunks else 0) for c in range(chunks)]
    out, lo = [], 0
    for n in sizes:
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


class PeerPushAllGather:
    """All-gather by PEER PUSH over NVLink with the copy engines — no collective kernel.

    Why not NCCL here: the pipeline's kernels are persistent (one resident CTA set per SM that
    walks all tiles), so an NCCL all-gather kernel launched next to them gets no SM until they
    retire and the "overlap" serialises (measured at 2 GPUs: 11.7 ms compute + 4.6 ms gather =
    16.3 ms).  Every rank instead owns a symmetric buffer ``[world * count, ...]`` that all
    peers have mapped (``torch.distributed._symmetric_memory``: CUDA VMM / NVLink peer
    mappings through NVSwitch).  A rank's kernels write their results straight into its own
    slot of its own buffer; behind every chunk, on a side stream, plain device-to-device
    copies push that slot into the same position of every peer's buffer.  Copies run on the
    DMA engines, so they overlap the next chunk's kernels, and the 18 NVLink-5 links of a
    B200 are driven without spending SMs.  A device-side barrier on the signal pads closes
    the step.  After ``run`` the local buffer holds all ranks' results in GLOBAL order.
    """

    def __init__(self, count: int, world: int, rank: int, chunks: int, templates: Sequence[torch.Tensor],
                 group: int = 1, streams: int = 1, taper=False):
        import torch.distributed._symmetric_memory as symm
        self.count, self.world, self.rank = count, world, rank
        self.plan = chunk_plan(count, chunks, group, taper)
        self.buffers, self.handles, self.peers = [], [], []
        gname = dist.group.WORLD.group_name
        for t in templates:
            shape = (world * count,) + tuple(t.shape[1:])
            buf = symm.empty(shape, dtype=t.dtype, device=t.device)
            hdl = symm.rendezvous(buf, gname)
            self.buffers.append(buf)
            self.handles.append(hdl)
            self.peers.append([buf if p == rank else hdl.get_buffer(p, shape, t.dtype) for p in range(world)])
        # few side streams: pushing to all 7 peers concurrently from per-peer streams was measured
        # slower at 8 GPUs than one stream (33.9 vs 27.6 ms per step); see profiles/r1_scaling.md
        self.comm = [torch.cuda.Stream(device=templates[0].device) for _ in range(max(1, min(streams, world - 1)))]
        self.out = self.buffers

    def local_slot(self, k: int, lo: int = 0, hi: int = None) -> torch.Tensor:
        """View of this rank's own rows ``[lo, hi)`` inside buffer ``k`` (kernels write here)."""
        hi = self.count if hi is None else hi
        base = self.rank * self.count
        return self.buffers[k][base + lo:base + hi]

    def run(self, compute_chunk: Callable[[int, int], None]):
        """``compute_chunk(lo, hi)`` must write trajectories ``[lo, hi)`` into ``local_slot``."""
        cur = torch.cuda.current_stream()
        self.handles[0].barrier(channel=0)          # peers are done reading the previous step
        base = self.rank * self.count
        for lo, hi in self.plan:
            compute_chunk(lo, hi)
            for side in self.comm:
                side.wait_stream(cur)
            for shift in range(1, self.world):       # staggered so the links are loaded evenly
                p = (self.rank + shift) % self.world
                with torch.cuda.stream(self.comm[(shift - 1) % len(self.comm)]):
                    for k, buf in enumerate(self.buffers):
                        self.peers[k][p][base + lo:base + hi].copy_(buf[base + lo:base + hi], non_blocking=True)
        for side in self.comm:
            cur.wait_stream(side)
        self.handles[0].barrier(channel=1)          # every peer's pushes into this buffer have landed
        return self.out


class PeerStoreGather:
    """All-gather by PEER STORES from inside the pipeline kernel (``mst_pipeline_wire``).

    Every rank owns symmetric buffers ``[world * count, ...]`` for the wire outputs — the float32
    polynomial matrix the reference's ``path_to_pol`` emits (scripts/drones_pols_generator.py:63-77)
    and / or the collision flags — mapped by all peers over NVLink
    (``torch.distributed._symmetric_memory``).  The single-pass kernel stores each finished
    trajectory's rows through all ``world`` base pointers (its own buffer first, then the peers,
    staggered by rank so the links load evenly): the exchange is spread over the whole kernel, no
    copy engine or collective kernel runs, and a device-side barrier on the signal pads closes the
    step.  ``mode``: ``"pol_matrix_f32"`` (matrix + flags) or ``"flags"`` (flags only).
    After ``run`` the local buffers hold all ranks' results in GLOBAL trajectory order.
    """

    def __init__(self, count: int, world: int, rank: int, n: int, K: int, S: int, device, mode: str = "pol_matrix_f32"):
        import torch.distributed._symmetric_memory as symm
        from .batch import make_wire_targets
        if mode not in ("pol_matrix_f32", "flags"):
            raise ValueError("mode must be 'pol_matrix_f32' or 'flags'")
        if world > 8:
            raise ValueError("at most 8 destinations (one NVSwitch domain)")
        self.count, self.world, self.rank, self.mode = count, world, rank, mode
        gname = dist.group.WORLD.group_name
        rows = world * count
        specs = []
        if mode == "pol_matrix_f32":
            specs.append(("pol_matrix", (rows, n, 1 + 8 * K), torch.float32))
        specs += [("hit", (rows, S), torch.uint8), ("any_hit", (rows,), torch.uint8)]
        self.buffers, self.handles, order = {}, [], [(rank + shift) % world for shift in range(world)]
        targets = {}
        for name, shape, dtype in specs:
            buf = symm.empty(shape, dtype=dtype, device=device)
            hdl = symm.rendezvous(buf, gname)
            self.buffers[name] = buf
            self.handles.append(hdl)
            targets[name] = [buf if p == rank else hdl.get_buffer(p, shape, dtype) for p in order]
        self.wire = make_wire_targets(targets.get("pol_matrix"), targets["hit"], targets["any_hit"],
                                      row_offset=rank * count)
        self._targets = targets

    def bytes_per_trajectory(self, n: int, K: int, S: int) -> int:
        return (n * (1 + 8 * K) * 4 if self.mode == "pol_matrix_f32" else 0) + S + 1

    def run(self, launch: Callable[[object], None]):
        """``launch(wire)`` must issue ``pipeline_wire(..., wire)`` for this rank's trajectories on
        the current stream."""
        self.handles[0].barrier(channel=0)          # peers are done reading the previous step
        launch(self.wire)
        self.handles[0].barrier(channel=1)          # every peer's stores into this buffer have landed
        return self.buffers
