"""Single-call latency of the drop-in surface (tool): what one OMPL ``isStateValid`` callback
(RB_planning_sep_coll_check.py:208-226) and one scalar ``Polynomial.eval`` cost through the CUDA
library, next to the batched calls.  Prints one JSON line; the numbers go to profiles/."""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import drone_path_planning_python_b200 as mst
    from drone_path_planning_python_b200 import meshio
    sys.path.insert(0, mst.dropin_path())
    import optimizations as o
    from RigidBodyPlanners.fcl_checker import Fcl_checker
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        ef, rf = os.path.join(tmp, "env.stl"), os.path.join(tmp, "robot.stl")
        meshio.write_stl(ef, meshio.shipped_mesh("env-scene-ltu-experiment"))
        meshio.write_stl(rf, meshio.shipped_mesh("custom_triangle_robot"))
        checker = Fcl_checker(ef, rf)
    rng = np.random.default_rng(0)
    states = np.concatenate([rng.uniform([-2.2, 2.8, 0.5], [2.2, 5.0, 2.5], (4000, 3)), rng.uniform(-3, 3, (4000, 1))], 1)
    quats = [(0.0, 0.0, float(np.sin(y / 2)), float(np.cos(y / 2))) for y in states[:, 3]]
    for i in range(200):
        checker.check_collision(states[i, :3], quats[i])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(4000):
        checker.check_collision(states[i, :3], quats[i])
    out["Fcl_checker.check_collision_us"] = (time.perf_counter() - t0) / 4000 * 1e6
    # the previous implementation of the same call: tensor API, host->device copy + launch + blocking read
    t0 = time.perf_counter()
    for i in range(500):
        int(mst.collide_poses(checker.robot.m, checker.env.m, np.concatenate([states[i, :3], quats[i]])[None])[0])
    out["collide_poses_single_via_tensors_us"] = (time.perf_counter() - t0) / 500 * 1e6
    t0 = time.perf_counter()
    for _ in range(20):
        checker.check_collision_batch(states)
    out["check_collision_batch_4000_states_us_per_state"] = (time.perf_counter() - t0) / 20 / 4000 * 1e6
    pol = o.Polynomial([1.0, 0.5, -0.25, 0.125, 0.0, 0.3, -0.1, 0.01])
    for _ in range(50):
        pol.eval(0.7)
    t0 = time.perf_counter()
    for _ in range(500):
        pol.eval(0.7)
    out["Polynomial.eval_us"] = (time.perf_counter() - t0) / 500 * 1e6
    pieces = [o.Polynomial(np.random.default_rng(i).normal(size=(8, 1))) for i in range(10)]
    pc = o.PiecewisePolynomial(pieces, [1.0] * 10)
    for _ in range(50):
        pc.eval(3.3)
    t0 = time.perf_counter()
    for i in range(500):
        pc.eval(0.01 * i)
    out["PiecewisePolynomial.eval_us"] = (time.perf_counter() - t0) / 500 * 1e6
    tr = o.Trajectory()
    tr.polynomials = [o.Polynomial4D(1.0, *[np.random.default_rng(7 * i + k).normal(size=8) for k in range(4)]) for i in range(10)]
    tr.duration = 10.0
    for _ in range(50):
        tr.eval(3.3)
    t0 = time.perf_counter()
    for i in range(500):
        tr.eval(0.01 * i)
    out["Trajectory.eval_us"] = (time.perf_counter() - t0) / 500 * 1e6
    ts = np.linspace(0, 1, 4000)
    t0 = time.perf_counter()
    for _ in range(20):
        pol.eval_many(ts)
    out["Polynomial.eval_many_4000_us_per_value"] = (time.perf_counter() - t0) / 20 / 4000 * 1e6
    print(json.dumps(out))


if __name__ == "__main__":
    main()
