"""Batched time-allocation search (BASELINE configs[2]: "free-time-allocation
re-solves").  NOT in the reference, whose callers hand fixed stamps to ``calculate_trajectory4D``
(src/optimizations/calculatingTrajectories.py:200-213): this is the extension the survey names as
the place where the snap objective belongs, built from the path's own operators — every
evaluation is one ``mst_solve_batch`` + ``mst_snap_cost`` launch pair over ALL problems and ALL
candidate allocations at once, and the gradient comes from the coefficients of one solve
(``mst_time_gradient``).

Method (the test suite holds a numpy restatement of the same steps as its checker):
projected gradient descent on the piece durations ``T[B, n]`` with each problem's first stamp and
total duration fixed.

* gradient of ``J(T)`` = snap cost of the min-snap solve through the waypoints: the knot derivatives
  are what the solve optimises, so only the explicit dependence on ``T_i`` counts (envelope theorem)
  and ``dJ/dT_i`` is minus the Hamiltonian of piece ``i`` — read off the coefficients, no
  finite-difference re-solves; it is then projected on ``sum T = const``;
* candidates ``T - alpha_k g`` with ``alpha_k = cap / 2^k`` (``L`` per problem, one batch), ``cap``
  chosen so that no duration falls below ``min_fraction`` of the mean duration;
* the best candidate replaces ``T`` where it lowers the cost; otherwise the problem keeps its ``T``.

Allocations that push a problem onto the pivoted solver (wide duration spreads, ``csrc/condensed_core.cuh``) are
handled by ``mst_solve_batch`` itself; a candidate whose solve fails counts as cost ``+inf``.
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import _abi
from .batch import _f64, snap_cost, solve_batch, time_gradient


def _evaluate(wp_rep: torch.Tensor, t0: torch.Tensor, T: torch.Tensor, solver: str):
    """Cost and gradient for stacked candidates: ``wp_rep[N, n+1, K]``, ``t0[N]``, ``T[N, n]`` ->
    ``J[N]`` (``+inf`` where the solve failed), ``dJ/dT[N, n]``."""
    t = torch.cat([t0[:, None], t0[:, None] + torch.cumsum(T, dim=1)], dim=1)
    coef, dur, info = solve_batch(wp_rep, t, solver=solver)
    cost = snap_cost(coef, dur)
    ok = (info == 0) & torch.isfinite(cost)
    grad = time_gradient(coef)
    grad = torch.where(ok[:, None] & torch.isfinite(grad), grad, torch.zeros_like(grad))
    return torch.where(ok, cost, torch.full_like(cost, float("inf"))), grad


def optimize_time_allocation(wp, t, iters: int = 8, line_search: int = 6, min_fraction: float = 0.1,
                             solver: str = "auto") -> Tuple[torch.Tensor, torch.Tensor]:
    """``wp[B, n+1, K]``, ``t[B, n+1]`` -> ``(t_new[B, n+1], cost[iters + 1, B])``.

    ``t_new`` keeps ``t[:, 0]`` and ``t[:, -1]``; ``cost[0]`` is the snap cost of the given stamps,
    ``cost[j]`` the cost after ``j`` iterations (non-increasing per problem).  One iteration costs
    ``line_search`` solves per problem."""
    dev = _abi.require_cuda()
    wp = _f64(wp, dev)
    t = _f64(t, dev)
    if wp.dim() != 3 or t.dim() != 2 or t.shape[0] != wp.shape[0] or t.shape[1] != wp.shape[1]:
        raise ValueError("wp must be [B, n+1, K] and t [B, n+1]")
    B, m, K = wp.shape
    n = m - 1
    if B == 0:
        return t.clone(), torch.zeros((iters + 1, 0), dtype=torch.float64, device=dev)
    t0 = t[:, 0].contiguous()
    T = (t[:, 1:] - t[:, :-1]).contiguous()
    J, g = _evaluate(wp, t0, T, solver)
    if n < 2:
        return t.clone(), J[None].repeat(iters + 1, 1)
    total = T.sum(dim=1)
    floor = min_fraction * total / n
    L = line_search
    scale = 0.5 ** torch.arange(L, dtype=torch.float64, device=dev)                  # alpha_k / cap
    wp_l = wp[:, None].expand(B, L, m, K).reshape(B * L, m, K)
    t0_l = t0.repeat_interleave(L)
    history = [J]
    for _ in range(iters):
        J0 = history[-1]
        g = g - g.mean(dim=1, keepdim=True)                                          # projection on sum T = const
        gmax = g.abs().amax(dim=1)
        cap = 0.5 * (T - floor[:, None]).amin(dim=1).clamp_min(0.0) / gmax.clamp_min(1e-300)
        cand = T[:, None, :] - (cap[:, None] * scale[None])[:, :, None] * g[:, None, :]   # [B, L, n]
        Jc, gc = _evaluate(wp_l, t0_l, cand.reshape(B * L, n), solver)
        Jc, gc = Jc.reshape(B, L), gc.reshape(B, L, n)
        best = Jc.argmin(dim=1)
        pick = best[:, None, None].expand(B, 1, n)
        Jb = Jc.gather(1, best[:, None])[:, 0]
        better = Jb < J0
        T = torch.where(better[:, None], cand.gather(1, pick)[:, 0], T)
        g = torch.where(better[:, None], gc.gather(1, pick)[:, 0], g)
        history.append(torch.where(better, Jb, J0))
    t_new = torch.cat([t0[:, None], t0[:, None] + torch.cumsum(T, dim=1)], dim=1)
    t_new[:, -1] = t[:, -1]   # the total is preserved up to rounding: pin the last stamp exactly
    return t_new, torch.stack(history)
