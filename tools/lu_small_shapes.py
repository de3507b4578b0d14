"""Small pivoted-solver runs (written for compute-sanitizer, which this pool does not allow; kept as a quick
regression run): shapes that cover one and several
right-hand sides per lane, the t[0] != 0 quirk, a singular group, bad stamps, list mode through AUTO."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drone_path_planning_python_b200 as mst
from oracle import minsnap_oracle as mo

rng = np.random.default_rng(11)
for B, n, K, G in ((64, 10, 3, 1), (60, 20, 4, 5), (36, 3, 3, 6), (40, 1, 3, 1), (34, 49, 3, 1)):
    groups = B // G
    T = np.clip(rng.uniform(0.5, 2, (groups, n)) * np.exp(rng.normal(size=(groups, n))), 0.05, 5.0)
    t = np.concatenate([np.zeros((groups, 1)), np.cumsum(T, 1)], 1)
    T[1] = rng.uniform(0.5, 2, n)    # benign durations under the quirk: with short pieces its matrix is near singular
    t = np.concatenate([np.zeros((groups, 1)), np.cumsum(T, 1)], 1)
    t[1] += 0.1                      # t[0] != 0
    if n > 2:
        t[2, 2] = t[2, 1]            # zero-length piece: singular
        t[3, 2] = t[3, 1] - 0.5      # decreasing
    wp = np.cumsum(rng.normal(0, 0.3, (B, n + 1, K)), 1)
    for solver in ("banded_lu", "auto"):
        coef, dur, info = mst.solve_batch(wp, t, share_time_group=G, solver=solver)
        torch.cuda.synchronize()
        info = info.cpu().numpy()
        ok = np.nonzero(info == 0)[0]
        worst = 0.0
        for b in ok[:6]:
            ref, _ = mo.solve_waypoints(wp[b], t[b // G])
            worst = max(worst, float((np.abs(coef[b].cpu().numpy() - ref).max(axis=(0, 2)) / np.abs(ref).max(axis=(0, 2))).max()))
        print(B, n, K, G, solver, "ok %d of %d, worst %.1e" % (len(ok), B, worst), flush=True)
        assert worst < 1e-9
print("done")
