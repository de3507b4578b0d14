"""What would a two-kernel split of sample+collide cost?  (a) the fused kernel with the obstacle moved
far away = its sampling half alone, (b) the pose-batch kernel on exactly the near poses of the
benchmark workload = its collision half alone on a dense list."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
import drone_path_planning_python_b200 as mst

def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

B = 1 << 20
wp, t = bench.make_workload(B, 1)
rs, es = bench.mesh_soups()
robot, env = mst.Mesh(rs), mst.Mesh(es)
far = mst.Mesh(es + np.array([1000.0, 0, 0]))
coef, dur, info = mst.solve_batch(wp, t)
S = bench.S_SAMPLES
print("fused, real obstacle : %.3f ms" % timeit(lambda: mst.collide_trajectories(coef, dur, S, robot, env)))
print("fused, obstacle far  : %.3f ms (sampling half alone)" % timeit(lambda: mst.collide_trajectories(coef, dur, S, robot, far)))
pos = mst.sample_batch(coef, dur, S=S).reshape(-1, 3)
verts = torch.as_tensor(np.unique(rs.reshape(-1, 3), axis=0), device=pos.device)
lo, hi = verts.amin(0), verts.amax(0)
e = torch.as_tensor(es.reshape(-1, 3), device=pos.device)
near = ((pos + hi >= e.amin(0)) & (pos + lo <= e.amax(0))).all(dim=1)
dense = pos[near].contiguous()
print("near poses: %d (%.1f %%)" % (dense.shape[0], 100.0 * dense.shape[0] / pos.shape[0]))
print("pose kernel on the near poses only: %.3f ms" % timeit(lambda: mst.collide_poses(robot, env, dense)))
print("pose kernel on all poses          : %.3f ms" % timeit(lambda: mst.collide_poses(robot, env, pos)))
print("sample_batch alone (writes 2.5 GB of positions): %.3f ms" % timeit(lambda: mst.sample_batch(coef, dur, S=S)))
