"""K = 4 (x, y, z, yaw) pipeline timing per stage (tool): solver, sample+collide, both pipelines."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drone_path_planning_python_b200 as mst
from bench import mesh_soups
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
rng = np.random.default_rng(1)
def timeit(fn, reps=reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
robot_s, env_s = mesh_soups()
robot, env = mst.Mesh(robot_s), mst.Mesh(env_s)
n, K, S = 10, 4, 100
T = rng.uniform(0.5, 2, (B, n)); t = torch.as_tensor(np.concatenate([np.zeros((B, 1)), np.cumsum(T, 1)], 1), device='cuda')
wp4 = np.zeros((B, n + 1, K)); wp4[:, :, :3] = rng.uniform([-2.2, 2.8, 0.5], [2.2, 5.0, 2.5], (B, 1, 3)) + np.cumsum(rng.normal(0, 0.3, (B, n + 1, 3)), 1)
wp4[:, :, 3] = np.cumsum(rng.normal(0, 0.1, (B, n + 1)), 1)
wp4 = torch.as_tensor(wp4, device='cuda')
out = mst.pipeline(wp4, t, S, robot, env)
scale = (1 << 20) / B
print("K=4 per 1M trajectories:")
print("  pipeline (two launches)  %.2f ms" % (timeit(lambda: mst.pipeline(wp4, t, S, robot, env, out=out)) * scale))
print("  pipeline (single pass)   %.2f ms" % (timeit(lambda: mst.pipeline(wp4, t, S, robot, env, out=out, solver="auto_one_pass")) * scale))
print("  solve_batch              %.2f ms" % (timeit(lambda: mst.solve_batch(wp4, t)) * scale))
print("  collide_trajectories     %.2f ms" % (timeit(lambda: mst.collide_trajectories(out.coef, out.dur, S, robot, env)) * scale))
print("  any-hit rate %.3f" % float(out.any_hit.float().mean()))
