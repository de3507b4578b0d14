"""Throughput of the pivoted banded solver alone (solver="banded_lu"), plus a spot check against the oracle."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drone_path_planning_python_b200 as mst
from oracle import minsnap_oracle as mo

rng = np.random.default_rng(5)
out = {}


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


CASES = ((65536, 20, 3, 1), (262144, 10, 3, 1), (65536, 10, 4, 1), (65536 * 5, 10, 3, 5), (16384, 49, 3, 1))
if len(sys.argv) > 1:
    CASES = CASES[:int(sys.argv[1])]
for B, n, K, G in CASES:
    groups = B // G
    T = np.clip(rng.uniform(0.5, 2, (groups, n)) * np.exp(rng.normal(size=(groups, n))), 0.05, 5.0)
    t_np = np.concatenate([np.zeros((groups, 1)), np.cumsum(T, 1)], 1)
    wp_np = np.cumsum(rng.normal(0, 0.3, (B, n + 1, K)), 1)
    t, wp = torch.as_tensor(t_np, device="cuda"), torch.as_tensor(wp_np, device="cuda")
    ms = timeit(lambda: mst.solve_batch(wp, t, share_time_group=G, solver="banded_lu"))
    coef, dur, info = mst.solve_batch(wp, t, share_time_group=G, solver="banded_lu")
    assert int((info != 0).sum()) == 0
    worst = 0.0
    for b in list(range(0, B, max(1, B // 24)))[:24]:
        ref, _ = mo.solve_waypoints(wp_np[b], t_np[b // G])
        got = coef[b].cpu().numpy()
        worst = max(worst, float((np.abs(got - ref).max(axis=(0, 2)) / np.abs(ref).max(axis=(0, 2))).max()))
    key = "B=%d n=%d K=%d G=%d" % (B, n, K, G)
    out[key] = {"ms": ms, "M_trajectories_per_s": B / ms / 1e3, "M_factorisations_per_s": groups / ms / 1e3,
                "worst_normwise_error_vs_oracle": worst}
    print(key, json.dumps(out[key]), flush=True)
print(json.dumps(out))
