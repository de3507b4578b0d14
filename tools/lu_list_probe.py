"""List-mode cost of the pivoted solver: a large batch of which a few percent have wide duration spreads."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drone_path_planning_python_b200 as mst

rng = np.random.default_rng(9)
B, n, K = 393216, 20, 3


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


wp = torch.as_tensor(np.cumsum(rng.normal(0, 1.0, (B, n + 1, K)), 1), device="cuda")
for share in (0.0, 0.03, 0.1):
    T = rng.uniform(0.6, 1.6, (B, n))
    wide = rng.uniform(size=B) < share
    T[wide, 0] = 0.12
    t = torch.as_tensor(np.concatenate([np.zeros((B, 1)), np.cumsum(T, 1)], 1), device="cuda")
    ms_auto = timeit(lambda: mst.solve_batch(wp, t))
    idx = torch.as_tensor(np.nonzero(wide)[0], device="cuda")
    ms_sub = timeit(lambda: mst.solve_batch(wp[idx], t[idx], solver="banded_lu")) if len(idx) else 0.0
    _, _, info = mst.solve_batch(wp, t)
    print("share %.2f: auto %.3f ms; its %d wide groups alone on the pivoted solver %.3f ms; failures %d" % (
        share, ms_auto, int(wide.sum()), ms_sub, int((info != 0).sum())), flush=True)
