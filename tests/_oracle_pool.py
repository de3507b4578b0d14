"""TEST INFRASTRUCTURE: the CPU oracle over many trajectories on a process pool (the numpy
restatement of calculate_trajectory1D + PiecewisePolynomial.eval for coefficients and sampled
positions, the C restatement of the collision test for flags)."""
from __future__ import annotations

import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_STATE = {}


def _init(robot, env, S):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    _STATE.update(robot=robot, env=env, S=S)


def _work(args):
    from oracle import build_oracle, minsnap_oracle as mo
    wp, t = args
    S = _STATE["S"]
    coefs, poss, flags = [], [], []
    for b in range(wp.shape[0]):
        coef, dur = mo.solve_waypoints(wp[b], t[b])
        ts = mo.uniform_sample_times(dur, S)
        pos = mo.sample_trajectory(coef, dur, ts)
        K = pos.shape[1]
        poses = pos if K == 4 else np.concatenate([pos[:, :3], np.zeros((S, 1))], axis=1)
        coefs.append(coef)
        poss.append(pos)
        flags.append(build_oracle.c_collide_poses(_STATE["robot"], _STATE["env"], poses))
    return np.stack(coefs), np.stack(poss), np.stack(flags)


def oracle_pipeline(wp, t, S, robot_tris, env_tris, procs=None):
    """``wp[B, n+1, K]``, ``t[B, n+1]`` -> oracle ``coef[B, n, K, 8]``, ``pos[B, S, K]``, ``hit[B, S]``."""
    from oracle import build_oracle
    build_oracle.build()
    procs = procs or max(1, min(16, len(os.sched_getaffinity(0))))
    parts = [p for p in np.array_split(np.arange(wp.shape[0]), procs * 4) if len(p)]
    with mp.get_context("spawn").Pool(procs, initializer=_init, initargs=(robot_tris, env_tris, S)) as pool:
        out = pool.map(_work, [(wp[p], t[p]) for p in parts])
    return (np.concatenate([o[0] for o in out]), np.concatenate([o[1] for o in out]),
            np.concatenate([o[2] for o in out]))
