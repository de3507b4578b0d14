"""Generate tests/golden/collision_anchors.npz: robot-pose-vs-environment collision answers
decided in EXACT rational arithmetic on the reference's shipped mesh pairs.

TEST INFRASTRUCTURE ONLY.  Run as ``python oracle/make_collision_anchors.py`` (needs only the
shipped meshes stored in drone_path_planning_python_b200/data/stl_meshes.npz; a few minutes of
``fractions`` arithmetic).  The CPU tests re-derive a subset and hold both restatements
(numpy, C) to every answer; the GPU tests hold the kernels to them.

What is anchored: ``Fcl_checker.check_collision`` (src/RigidBodyPlanners/fcl_checker.py:93-100) as
``isStateValid`` drives it (RB_planning_sep_coll_check.py:208-215) — robot mesh at a pose against
the environment at identity, "some triangle pair shares a point".  FCL itself is absent
(collision_oracle.py header), so the anchor is the mathematical predicate, not FCL's rounding:
  * lattice poses: translations on a 1/16 grid inside / around the obstacle, rotations that are
    exact in double (identity and half turns about x, y, z given as quaternions), so the posed
    robot corners are exact doubles and the rational decision IS the answer for the very inputs
    the double-precision code sees;
  * vertex-on-vertex poses: a robot corner placed exactly on an environment corner (touching:
    must be reported as a collision);
  * the one pose the reference itself evaluates, fcl_checker.py:124-136
    (robot-scene-triangle vs env-scene-ltu-experiment, T = [-1.21917, -0.441611, -0.0462389],
    q = [-0.298798, 0.00548747, 0.0160421, 0.954166]), with the rotation matrix FCL builds from
    that (non-unit) quaternion taken as exact.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

PAIRS = [("custom_triangle_robot", "env-scene-ltu-experiment"),
         ("custom_triangle_robot", "env-scene-narrow"),
         ("custom_triangle_robot", "env-scene-hole"),
         ("robot-scene-triangle", "env-scene-ltu-experiment")]
HALF_TURNS = np.array([[0, 0, 0, 1], [0, 0, 1, 0], [1, 0, 0, 0], [0, 1, 0, 0]], dtype=np.float64)  # xyzw, exact R
REFERENCE_POSE = np.array([-1.21917, -0.441611, -0.0462389, -0.298798, 0.00548747, 0.0160421, 0.954166])


def lattice_poses(rng, robot, env, count):
    lo, hi = env.reshape(-1, 3).min(0), env.reshape(-1, 3).max(0)
    reach = np.abs(robot.reshape(-1, 3)).max()
    pos = np.round(rng.uniform(lo - 0.6 * reach, hi + 0.6 * reach, (count, 3)) * 16) / 16
    quat = HALF_TURNS[rng.integers(0, 4, count)]
    return np.concatenate([pos, quat], axis=1)


def vertex_poses(rng, robot, env, count):
    rv, evs = robot.reshape(-1, 3), env.reshape(-1, 3)
    out = []
    for _ in range(count):
        q = HALF_TURNS[rng.integers(0, 4)]
        from oracle import collision_oracle as co
        R = co.quat_to_matrix(q)
        a, b = rv[rng.integers(len(rv))], evs[rng.integers(len(evs))]
        out.append(np.concatenate([b - R @ a, q]))       # exact: signed copies and one subtraction of float32 values
    return np.asarray(out)


def main():
    from drone_path_planning_python_b200 import meshio
    from oracle import collision_oracle as co, exact_geometry as xg
    rng = np.random.default_rng(20261020)
    out = {}
    for pi, (rname, ename) in enumerate(PAIRS):
        robot = co.mesh_triangles(meshio.shipped_mesh(rname))
        env = co.mesh_triangles(meshio.shipped_mesh(ename))
        poses = np.concatenate([lattice_poses(rng, robot, env, 220), vertex_poses(rng, robot, env, 30)])
        kind = np.array([0] * 220 + [1] * 30)
        if (rname, ename) == ("robot-scene-triangle", "env-scene-ltu-experiment"):
            poses = np.concatenate([poses, REFERENCE_POSE[None]])
            kind = np.concatenate([kind, [2]])
        R, T = co.pose_matrices(poses)
        exact = np.array([xg.robot_meets_env(robot, env, R[i], T[i]) for i in range(len(poses))], dtype=np.uint8)
        flags, margin = co.collide_poses(robot, env, poses, with_margin=True)
        key = "pair%d" % pi
        out[key + "__robot"] = np.array(rname)
        out[key + "__env"] = np.array(ename)
        out[key + "__poses"] = poses
        out[key + "__kind"] = kind
        out[key + "__exact"] = exact
        out[key + "__margin"] = margin
        agree = flags == exact
        print("%-24s vs %-26s: %3d poses, %3d collide exactly, restatement agrees on %d (disagreements at |margin| <= %.1e)"
              % (rname, ename, len(poses), int(exact.sum()), int(agree.sum()),
                 float(np.abs(margin[~agree]).max()) if (~agree).any() else 0.0))
        assert (exact[kind == 1] == 1).all(), "a shared corner is a collision"
    np.savez(os.path.join(ROOT, "tests", "golden", "collision_anchors.npz"), **out)


if __name__ == "__main__":
    main()
