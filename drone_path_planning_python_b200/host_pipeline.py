"""End-to-end pipeline over HOST buffers: what a caller holding numpy / pinned torch arrays
gets from the plugin.  The batch is cut into chunks that rotate through a small ring of
device-side slots; each slot has its own stream, so chunk i+1's host->device copy and chunk
i-1's device->host copy overlap chunk i's kernels (PCIe is full duplex; the copy engines and
the SMs run concurrently).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _abi
from .batch import Mesh, PipelineResult, pack_pol_matrix, pipeline


class HostPipeline:
    """Preallocated chunked solve+sample+collide over host tensors.

    ``run(wp, t, out)`` takes pinned host tensors ``wp[B, n+1, K]`` / ``t[B/G, n+1]`` and
    fills the pinned host tensors of ``out`` (a ``PipelineResult`` of host tensors).  ``run`` is
    synchronous for the HOST: it returns after the last device->host copy of every slot has
    completed (an event per slot), so ``out`` can be read right away.  ``run(..., wait=False)``
    returns the slot events instead for callers that overlap further work."""

    def __init__(self, n: int, K: int, S: int, robot: Mesh, env: Mesh, chunk: int = 65536,
                 share_time_group: int = 1, solver: str = "auto", slots: int = 3, wire: str = "f64"):
        """``wire="f64"`` returns coefficients and durations in double precision;
        ``wire="pol_matrix_f32"`` returns the reference's own output format instead — the
        float32 ``(n, 1 + 8K)`` matrix ``[T | x | y | z (| yaw)]`` that ``path_to_pol`` writes and
        publishes (scripts/drones_pols_generator.py:63-87), packed on the device
        (``mst_pack_pol_matrix``) — which halves the device->host traffic."""
        if wire not in ("f64", "pol_matrix_f32"):
            raise ValueError("wire must be 'f64' or 'pol_matrix_f32'")
        self.wire = wire
        self.dev = _abi.require_cuda()
        self.n, self.K, self.S, self.G = n, K, S, int(share_time_group)
        self.robot, self.env, self.solver = robot, env, solver
        if chunk < self.G:
            raise ValueError("chunk (%d) must hold at least one time-sharing group of %d trajectories" % (chunk, self.G))
        self.chunk = chunk - chunk % self.G
        f64, dev = torch.float64, self.dev
        self.slots = []
        for _ in range(slots):
            self.slots.append({
                "stream": torch.cuda.Stream(device=dev),
                "done": torch.cuda.Event(),
                "wp": torch.empty((self.chunk, n + 1, K), dtype=f64, device=dev),
                "t": torch.empty((self.chunk // self.G, n + 1), dtype=f64, device=dev),
                "mat": torch.empty((self.chunk, n, 1 + 8 * K), dtype=torch.float32, device=dev)
                if wire == "pol_matrix_f32" else None,
                "res": PipelineResult(torch.empty((self.chunk, n, K, 8), dtype=f64, device=dev),
                                      torch.empty((self.chunk, n), dtype=f64, device=dev),
                                      torch.empty((self.chunk,), dtype=torch.int32, device=dev),
                                      torch.empty((self.chunk, S), dtype=torch.uint8, device=dev),
                                      torch.empty((self.chunk,), dtype=torch.uint8, device=dev)),
            })

    @staticmethod
    def alloc_host_result(B: int, n: int, K: int, S: int, wire: str = "f64") -> PipelineResult:
        """Pinned host buffers for ``run``.  With ``wire="pol_matrix_f32"`` the ``coef`` field is
        the float32 ``[B, n, 1 + 8K]`` matrix and ``dur`` is None (column 0 of the matrix)."""
        pin = dict(pin_memory=True)
        if wire == "pol_matrix_f32":
            return PipelineResult(torch.empty((B, n, 1 + 8 * K), dtype=torch.float32, **pin), None,
                                  torch.empty((B,), dtype=torch.int32, **pin),
                                  torch.empty((B, S), dtype=torch.uint8, **pin),
                                  torch.empty((B,), dtype=torch.uint8, **pin))
        return PipelineResult(torch.empty((B, n, K, 8), dtype=torch.float64, **pin),
                              torch.empty((B, n), dtype=torch.float64, **pin),
                              torch.empty((B,), dtype=torch.int32, **pin),
                              torch.empty((B, S), dtype=torch.uint8, **pin),
                              torch.empty((B,), dtype=torch.uint8, **pin))

    def bytes_per_trajectory(self):
        """(host->device, device->host) bytes moved per trajectory."""
        h2d = (self.n + 1) * self.K * 8 + (self.n + 1) * 8 / self.G
        if self.wire == "pol_matrix_f32":
            d2h = self.n * (1 + 8 * self.K) * 4 + 4 + self.S + 1
        else:
            d2h = self.n * self.K * 64 + self.n * 8 + 4 + self.S + 1
        return h2d, d2h

    def run(self, wp: torch.Tensor, t: torch.Tensor, out: Optional[PipelineResult] = None, wait: bool = True):
        B = wp.shape[0]
        if B % self.G != 0 or t.shape[0] != B // self.G:
            raise ValueError("t must be [B/share_time_group, n+1]")
        if out is None:
            out = self.alloc_host_result(B, self.n, self.K, self.S, self.wire)
        caller = torch.cuda.current_stream()
        for i, b0 in enumerate(range(0, B, self.chunk)):
            nb = min(self.chunk, B - b0)
            slot = self.slots[i % len(self.slots)]
            st = slot["stream"]
            if i < len(self.slots):
                st.wait_stream(caller)
            with torch.cuda.stream(st):
                # stream order makes the slot's previous device->host copies finish first
                slot["wp"][:nb].copy_(wp[b0:b0 + nb], non_blocking=True)
                g0, ng = b0 // self.G, nb // self.G
                slot["t"][:ng].copy_(t[g0:g0 + ng], non_blocking=True)
                r = slot["res"]
                view = PipelineResult(r.coef[:nb], r.dur[:nb], r.info[:nb], r.hit[:nb], r.any_hit[:nb])
                mat = slot["mat"][:nb] if self.wire == "pol_matrix_f32" else None
                pipeline(slot["wp"][:nb], slot["t"][:ng], self.S, self.robot, self.env,
                         share_time_group=self.G, solver=self.solver, out=view, pol_matrix=mat)
                if self.wire == "pol_matrix_f32":   # packed by the solver kernel itself
                    out.coef[b0:b0 + nb].copy_(mat, non_blocking=True)
                else:
                    out.coef[b0:b0 + nb].copy_(view.coef, non_blocking=True)
                    out.dur[b0:b0 + nb].copy_(view.dur, non_blocking=True)
                out.info[b0:b0 + nb].copy_(view.info, non_blocking=True)
                out.hit[b0:b0 + nb].copy_(view.hit, non_blocking=True)
                out.any_hit[b0:b0 + nb].copy_(view.any_hit, non_blocking=True)
                slot["done"].record(st)
        used = self.slots[:min(len(self.slots), (B + self.chunk - 1) // self.chunk)]
        for slot in used:
            caller.wait_stream(slot["stream"])
        if not wait:
            return out, [slot["done"] for slot in used]
        # the copies above are asynchronous: order the HOST behind each slot's last device->host copy
        for slot in used:
            slot["done"].synchronize()
        return out
