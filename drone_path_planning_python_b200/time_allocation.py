"""Batched time-allocation search (BASELINE configs[2]: "free-time-allocation
re-solves").  NOT in the reference, whose callers hand fixed stamps to ``calculate_trajectory4D``
(src/optimizations/calculatingTrajectories.py:200-213): this is the extension the survey names as
the place where the snap objective belongs, built from the path's own operators — every
evaluation is one ``mst_solve_batch`` + ``mst_snap_cost`` launch pair over ALL problems and ALL
candidate allocations at once.

Method (the test suite holds a numpy restatement of the same steps as its checker):
projected gradient descent on the piece durations ``T[B, n]`` with each problem's first stamp and
total duration fixed.

* directional derivatives of ``J(T)`` = snap cost of the min-snap solve through the waypoints along
  the ``n`` tangent directions ``u_i = e_i - 1/n`` by forward differences (``n`` re-solves per
  problem, one batch): exactly the components of the gradient projected on ``sum T = const``;
* candidates ``T - alpha_k g`` with ``alpha_k = cap / 2^k`` (``L`` per problem, one batch), ``cap``
  chosen so that no duration falls below ``min_fraction`` of the mean duration;
* the best candidate replaces ``T`` where it lowers the cost; otherwise the problem keeps its ``T``.

Allocations that push a problem onto the pivoted solver (duration spread > 4, ``mst.h``) are
handled by ``mst_solve_batch`` itself; a candidate whose solve fails counts as cost ``+inf``.
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import _abi
from .batch import _f64, snap_cost, solve_batch


def _cost(wp_rep: torch.Tensor, t0: torch.Tensor, T: torch.Tensor, solver: str) -> torch.Tensor:
    """``J`` for stacked candidates: ``wp_rep[N, n+1, K]``, ``t0[N]``, ``T[N, n]`` -> ``[N]``."""
    t = torch.cat([t0[:, None], t0[:, None] + torch.cumsum(T, dim=1)], dim=1)
    coef, dur, info = solve_batch(wp_rep, t, solver=solver)
    cost = snap_cost(coef, dur)
    return torch.where((info == 0) & torch.isfinite(cost), cost, torch.full_like(cost, float("inf")))


def optimize_time_allocation(wp, t, iters: int = 8, line_search: int = 6, rel_step: float = 1e-4,
                             min_fraction: float = 0.1, solver: str = "auto"
                             ) -> Tuple[torch.Tensor, torch.Tensor]:
    """``wp[B, n+1, K]``, ``t[B, n+1]`` -> ``(t_new[B, n+1], cost[iters + 1, B])``.

    ``t_new`` keeps ``t[:, 0]`` and ``t[:, -1]``; ``cost[0]`` is the snap cost of the given stamps,
    ``cost[j]`` the cost after ``j`` iterations (non-increasing per problem)."""
    dev = _abi.require_cuda()
    wp = _f64(wp, dev)
    t = _f64(t, dev)
    if wp.dim() != 3 or t.dim() != 2 or t.shape[0] != wp.shape[0] or t.shape[1] != wp.shape[1]:
        raise ValueError("wp must be [B, n+1, K] and t [B, n+1]")
    B, m, K = wp.shape
    n = m - 1
    if n < 2 or B == 0:
        J = _cost(wp, t[:, 0], t[:, 1:] - t[:, :-1], solver) if B else torch.zeros((0,), dtype=torch.float64, device=dev)
        return t.clone(), J[None].repeat(iters + 1, 1)
    t0 = t[:, 0].contiguous()
    T = (t[:, 1:] - t[:, :-1]).contiguous()
    total = T.sum(dim=1)
    floor = min_fraction * total / n
    h = rel_step * total / n
    eye = torch.eye(n, dtype=torch.float64, device=dev) - 1.0 / n                    # rows u_i
    scale = 0.5 ** torch.arange(line_search, dtype=torch.float64, device=dev)         # alpha_k / cap
    wp_n = wp[:, None].expand(B, n, m, K).reshape(B * n, m, K)
    wp_l = wp[:, None].expand(B, line_search, m, K).reshape(B * line_search, m, K)
    history = [_cost(wp, t0, T, solver)]
    for _ in range(iters):
        J0 = history[-1]
        probe = T[:, None, :] + h[:, None, None] * eye[None]                         # [B, n, n]
        Jp = _cost(wp_n, t0.repeat_interleave(n), probe.reshape(B * n, n), solver).reshape(B, n)
        g = (Jp - J0[:, None]) / h[:, None]
        g = torch.where(torch.isfinite(g), g, torch.zeros_like(g))
        g = g - g.mean(dim=1, keepdim=True)                                          # rounding only
        gmax = g.abs().amax(dim=1)
        cap = 0.5 * (T - floor[:, None]).amin(dim=1).clamp_min(0.0) / gmax.clamp_min(1e-300)
        cand = T[:, None, :] - (cap[:, None] * scale[None])[:, :, None] * g[:, None, :]   # [B, L, n]
        Jc = _cost(wp_l, t0.repeat_interleave(line_search), cand.reshape(B * line_search, n), solver)
        Jc = Jc.reshape(B, line_search)
        best = Jc.argmin(dim=1)
        Jb = Jc.gather(1, best[:, None])[:, 0]
        better = Jb < J0
        Tb = cand.gather(1, best[:, None, None].expand(B, 1, n))[:, 0]
        T = torch.where(better[:, None], Tb, T)
        history.append(torch.where(better, Jb, J0))
    t_new = torch.cat([t0[:, None], t0[:, None] + torch.cumsum(T, dim=1)], dim=1)
    t_new[:, -1] = t[:, -1]   # the total is preserved up to rounding: pin the last stamp exactly
    return t_new, torch.stack(history)
