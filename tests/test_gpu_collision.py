"""GPU parity tests for the collision half and the fused pipeline.

The oracle is a restatement of FCL's published mesh-mesh test (parity UNPINNED: FCL is not
in the reference tree — see oracle/collision_oracle.py).  Flags must be bit-exact except for
poses whose margin to the touching boundary is below EPS (north star).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

EPS = 1e-9
BOUNDS_LO = np.array([-2.2, 2.8, 0.5])   # RB_planning_sep_coll_check.py:102-110
BOUNDS_HI = np.array([2.2, 5.0, 2.5])


def _soup(name):
    from drone_path_planning_python_b200 import meshio
    verts, _, tris = meshio.ingest_mesh(meshio.shipped_mesh(name))
    return meshio.triangle_soup(verts, tris)


def _random_poses(rng, P, dim, env_tris=None):
    if env_tris is None:
        pos = rng.uniform(BOUNDS_LO - 0.5, BOUNDS_HI + 0.5, (P, 3))
    else:  # around the obstacle, whatever frame the mesh was modelled in
        flat = env_tris.reshape(-1, 3)
        pos = rng.uniform(flat.min(axis=0) - 0.8, flat.max(axis=0) + 0.8, (P, 3))
    if dim == 3:
        return pos
    yaw = rng.uniform(-np.pi, np.pi, P)
    if dim == 4:
        return np.concatenate([pos, yaw[:, None]], axis=1)
    q = rng.normal(size=(P, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return np.concatenate([pos, q], axis=1)


@pytest.mark.parametrize("env_name", ["env-scene-ltu-experiment", "env-scene-narrow", "env-scene-hole"])
@pytest.mark.parametrize("dim", [4, 7])
def test_collide_poses_vs_oracle(env_name, dim):
    from oracle import collision_oracle as co
    import drone_path_planning_python_b200 as mst
    rng = np.random.default_rng(sum(env_name.encode()) * 10 + dim)
    robot_tris, env_tris = _soup("custom_triangle_robot"), _soup(env_name)
    robot, env = mst.Mesh(robot_tris), mst.Mesh(env_tris)
    P = 3000 if len(env_tris) > 24 else 6000
    poses = _random_poses(rng, P, dim, env_tris)
    hit = mst.collide_poses(robot, env, poses).cpu().numpy()
    ref, margin = co.collide_poses(robot_tris, env_tris, poses, with_margin=True)
    clear = np.abs(margin) > EPS
    assert clear.mean() > 0.99
    assert np.array_equal(hit[clear], ref[clear])
    assert 0.02 < ref.mean() < 0.9     # both outcomes are exercised


@pytest.mark.parametrize("env_name", ["env-scene-ltu-experiment", "env-scene-hole-narrow"])
@pytest.mark.parametrize("dim", [3, 4])
def test_poses_grazing_triangle_edges_vs_oracle(env_name, dim):
    """The cursor culls robot triangles that lie wholly beyond one of an env triangle's edge
    planes.  Poses that put a robot vertex within millimetres of an env triangle's edge (inside,
    outside, along it) are where a wrong cull would flip an answer; 56-triangle scenes also walk
    two 32-triangle blocks of the box mask."""
    from oracle import collision_oracle as co
    import drone_path_planning_python_b200 as mst
    rng = np.random.default_rng(77 + dim + len(env_name))
    robot_tris, env_tris = _soup("custom_triangle_robot"), _soup(env_name)
    robot, env = mst.Mesh(robot_tris), mst.Mesh(env_tris)
    P = 4000
    verts = np.unique(robot_tris.reshape(-1, 3), axis=0)
    e = rng.integers(0, len(env_tris), P)
    k = rng.integers(0, 3, P)
    a, b = env_tris[e, k], env_tris[e, (k + 1) % 3]
    on_edge = a + rng.uniform(-0.1, 1.1, (P, 1)) * (b - a)
    yaw = rng.uniform(-np.pi, np.pi, P) if dim == 4 else np.zeros(P)
    v = verts[rng.integers(0, len(verts), P)]
    c, s_ = np.cos(yaw), np.sin(yaw)
    v_world = np.stack([c * v[:, 0] - s_ * v[:, 1], s_ * v[:, 0] + c * v[:, 1], v[:, 2]], axis=1)
    pos = on_edge - v_world + rng.normal(0, 1.0, (P, 3)) * 10.0 ** rng.uniform(-6, -1, (P, 1))
    poses4 = np.concatenate([pos, yaw[:, None]], axis=1)
    hit = mst.collide_poses(robot, env, pos if dim == 3 else poses4).cpu().numpy()
    ref, margin = co.collide_poses(robot_tris, env_tris, poses4, with_margin=True)
    clear = np.abs(margin) > EPS
    assert clear.mean() > 0.9
    assert np.array_equal(hit[clear], ref[clear])
    assert 0.05 < ref.mean() < 0.999


def test_collide_translation_only_and_degenerate_robot():
    from oracle import collision_oracle as co
    import drone_path_planning_python_b200 as mst
    rng = np.random.default_rng(3)
    robot_tris = _soup("robot-scene-triangle")      # has 4 zero-area facets
    env_tris = _soup("env-scene-ltu-experiment")
    robot, env = mst.Mesh(robot_tris), mst.Mesh(env_tris)
    pos = _random_poses(rng, 4000, 3)
    hit = mst.collide_poses(robot, env, pos).cpu().numpy()
    ref, margin = co.collide_poses(robot_tris, env_tris, np.concatenate([pos, np.zeros((4000, 1))], 1),
                                   with_margin=True)
    clear = np.abs(margin) > EPS
    assert np.array_equal(hit[clear], ref[clear])


def test_known_answers_live_pair():
    """Weak anchors from SURVEY §8c: the wall of env-scene-ltu-experiment spans
    x in [-2,2], y in [3.9,4.1], z in [0,1.6]; a robot centred in it straddles both faces;
    the planner's start/goal states and a pass well above the wall are free."""
    import drone_path_planning_python_b200 as mst
    robot, env = mst.Mesh(_soup("custom_triangle_robot")), mst.Mesh(_soup("env-scene-ltu-experiment"))
    poses = np.array([[0.0, 4.0, 1.0, 0.0],     # inside the wall -> collision
                      [0.0, 4.0, 1.0, 1.2],
                      [0.0, 3.0, 1.0, 0.0],     # planner start (scripts/rigidBodyPath.py:146)
                      [0.0, 5.0, 1.0, 0.0],     # planner goal  (:147)
                      [0.0, 4.0, 2.17, 0.3]])   # over the wall, as the shipped path does
    hit = mst.collide_poses(robot, env, poses).cpu().numpy()
    assert hit.tolist() == [1, 1, 0, 0, 0]


def test_pipeline_equals_separate_stages():
    import drone_path_planning_python_b200 as mst
    rng = np.random.default_rng(9)
    robot, env = mst.Mesh(_soup("custom_triangle_robot")), mst.Mesh(_soup("env-scene-ltu-experiment"))
    for K in (3, 4):
        B, n, S = 300, 10, 100
        T = rng.uniform(0.5, 2.0, (B, n))
        t = np.concatenate([np.zeros((B, 1)), np.cumsum(T, axis=1)], axis=1)
        start = rng.uniform(BOUNDS_LO, BOUNDS_HI, (B, 1, 3))
        wp = np.zeros((B, n + 1, K))
        wp[:, :, :3] = start + np.cumsum(rng.normal(0, 0.3, (B, n + 1, 3)), axis=1)
        if K == 4:
            wp[:, :, 3] = np.cumsum(rng.normal(0, 0.1, (B, n + 1)), axis=1)
        res = mst.pipeline(wp, t, S, robot, env)
        coef, dur, info = mst.solve_batch(wp, t)
        assert np.array_equal(res.coef.cpu().numpy(), coef.cpu().numpy())
        assert np.array_equal(res.info.cpu().numpy(), info.cpu().numpy())
        pos = mst.sample_batch(coef, dur, S=S)
        hit = mst.collide_poses(robot, env, pos.reshape(B * S, K)).reshape(B, S)
        assert np.array_equal(res.hit.cpu().numpy(), hit.cpu().numpy())
        assert np.array_equal(res.any_hit.cpu().numpy(), hit.cpu().numpy().max(axis=1))
        assert 0 < res.any_hit.float().mean() < 1


def test_pipeline_samples_match_oracle_end_to_end():
    from oracle import collision_oracle as co
    from oracle import minsnap_oracle as mo
    import drone_path_planning_python_b200 as mst
    rng = np.random.default_rng(21)
    robot_tris, env_tris = _soup("custom_triangle_robot"), _soup("env-scene-ltu-experiment")
    robot, env = mst.Mesh(robot_tris), mst.Mesh(env_tris)
    B, n, K, S = 24, 10, 3, 100
    T = rng.uniform(0.5, 2.0, (B, n))
    t = np.concatenate([np.zeros((B, 1)), np.cumsum(T, axis=1)], axis=1)
    wp = rng.uniform(BOUNDS_LO, BOUNDS_HI, (B, 1, 3)) + np.cumsum(rng.normal(0, 0.3, (B, n + 1, 3)), axis=1)
    res = mst.pipeline(wp, t, S, robot, env)
    hit = res.hit.cpu().numpy()
    for b in range(B):
        coef, dur = mo.solve_waypoints(wp[b], t[b])
        ts = mo.uniform_sample_times(dur, S)
        pos = mo.sample_trajectory(coef, dur, ts)
        ref, margin = co.collide_poses(robot_tris, env_tris, np.concatenate([pos, np.zeros((S, 1))], 1),
                                       with_margin=True)
        clear = np.abs(margin) > 1e-7     # positions themselves carry ~1e-12 solver differences
        assert np.array_equal(hit[b][clear], ref[clear]), b


@pytest.mark.parametrize("dim", [3, 4, 7])
def test_every_pose_is_answered_once(dim):
    """Regression for the warp ring of the collision engine: 1 M poses, output prefilled with a
    sentinel — every entry must be overwritten with 0/1, identically on a second run and under a
    permutation of the poses (lost ring entries showed up as never-written flags)."""
    import torch
    import drone_path_planning_python_b200 as mst
    from drone_path_planning_python_b200 import _abi
    lib = _abi.load()
    rng = np.random.default_rng(24 + dim)
    robot_tris, env_tris = _soup("custom_triangle_robot"), _soup("env-scene-ltu-experiment")
    robot, env = mst.Mesh(robot_tris), mst.Mesh(env_tris)
    P = 1 << 20
    poses = torch.as_tensor(_random_poses(rng, P, dim, env_tris), device="cuda")
    stream = torch.cuda.current_stream().cuda_stream

    def run(p):
        out = torch.full((P,), 7, dtype=torch.uint8, device="cuda")
        assert lib.mst_collide_poses(robot.handle, env.handle, p.data_ptr(), P, dim, out.data_ptr(), stream) == 0
        return out
    first = run(poses)
    assert int((first > 1).sum()) == 0
    assert torch.equal(run(poses), first)
    perm = torch.randperm(P, device="cuda")
    assert torch.equal(run(poses[perm].contiguous()), first[perm])


def test_pipeline_answers_every_sample():
    import torch
    import drone_path_planning_python_b200 as mst
    rng = np.random.default_rng(31)
    robot, env = mst.Mesh(_soup("custom_triangle_robot")), mst.Mesh(_soup("env-scene-ltu-experiment"))
    for K in (3, 4):
        B, n, S = 50000, 10, 100
        T = rng.uniform(0.5, 2.0, (B, n))
        t = np.concatenate([np.zeros((B, 1)), np.cumsum(T, axis=1)], axis=1)
        wp = np.zeros((B, n + 1, K))
        wp[:, :, :3] = rng.uniform(BOUNDS_LO, BOUNDS_HI, (B, 1, 3)) + np.cumsum(rng.normal(0, 0.3, (B, n + 1, 3)), axis=1)
        if K == 4:
            wp[:, :, 3] = np.cumsum(rng.normal(0, 0.3, (B, n + 1)), axis=1)
        coef, dur, info = mst.solve_batch(wp, t)
        out = mst.PipelineResult(coef, dur, info, torch.full((B, S), 9, dtype=torch.uint8, device="cuda"),
                                 torch.full((B,), 9, dtype=torch.uint8, device="cuda"))
        hit, any_hit = mst.collide_trajectories(coef, dur, S, robot, env)
        res = mst.pipeline(wp, t, S, robot, env, out=out)
        assert int((res.hit > 1).sum()) == 0 and int((res.any_hit > 1).sum()) == 0
        assert torch.equal(res.hit, hit) and torch.equal(res.any_hit, any_hit)
        assert torch.equal(res.any_hit, res.hit.amax(dim=1))


def test_mesh_level_exact_anchors(golden_dir):
    """Kernels vs the robot-pose-vs-environment answers decided in exact rational arithmetic on the
    shipped mesh pairs (tests/golden/collision_anchors.npz, oracle/make_collision_anchors.py; the
    predicate fcl.collide decides at fcl_checker.py:93-100): lattice poses, exact half turns, shared
    corners (touching = collision) and the pose the reference itself evaluates (fcl_checker.py:133-136).
    Every answer must match — no epsilon band."""
    import os
    import sys
    import drone_path_planning_python_b200 as mst
    with np.load(os.path.join(golden_dir, "collision_anchors.npz")) as z:
        data = {k: z[k] for k in z.files}
    seen_reference_pose = False
    for i in range(4):
        key = "pair%d__" % i
        robot, env = mst.Mesh(_soup(str(data[key + "robot"]))), mst.Mesh(_soup(str(data[key + "env"])))
        poses, exact, kind = data[key + "poses"], data[key + "exact"], data[key + "kind"]
        hit = mst.collide_poses(robot, env, poses).cpu().numpy()
        assert np.array_equal(hit, exact), (str(data[key + "env"]), np.flatnonzero(hit != exact))
        # pure translations among them through the translation-only kernel as well
        ident = (poses[:, 3:] == [0, 0, 0, 1]).all(axis=1)
        hit3 = mst.collide_poses(robot, env, poses[ident, :3]).cpu().numpy()
        assert np.array_equal(hit3, exact[ident])
        if (kind == 2).any():
            seen_reference_pose = True
            j = int(np.flatnonzero(kind == 2)[0])
            assert exact[j] == 0 and hit[j] == 0
            # the same query through the drop-in, the way fcl_checker.py:124-136 issues it
            sys.path.insert(0, mst.dropin_path())
            try:
                from RigidBodyPlanners.fcl_checker import Fcl_checker
                import tempfile
                from drone_path_planning_python_b200 import meshio
                with tempfile.TemporaryDirectory() as tmp:
                    ef, rf = os.path.join(tmp, "env.stl"), os.path.join(tmp, "robot.stl")
                    meshio.write_stl(ef, meshio.shipped_mesh("env-scene-ltu-experiment"))
                    meshio.write_stl(rf, meshio.shipped_mesh("robot-scene-triangle"))
                    checker = Fcl_checker(ef, rf)
                    checker.set_robot_transform([-1.21917, -0.441611, -0.0462389],
                                                [-0.298798, 0.00548747, 0.0160421, 0.954166])
                    assert checker.check_collision() == 0
            finally:
                sys.path.remove(mst.dropin_path())
                for name in [m for m in sys.modules if m.split(".")[0] == "RigidBodyPlanners"]:
                    del sys.modules[name]
    assert seen_reference_pose


def test_state_validity_loop_equals_batch(tmp_path):
    """2,000 isStateValid-style single queries through the Fcl_checker drop-in
    (RB_planning_sep_coll_check.py:208-215: set pose, check_collision, `not collision`) give the
    same answers as one check_collision_batch launch and as the oracle outside the touching band."""
    import sys
    import drone_path_planning_python_b200 as mst
    from drone_path_planning_python_b200 import meshio
    from oracle import collision_oracle as co
    sys.path.insert(0, mst.dropin_path())
    try:
        from RigidBodyPlanners.fcl_checker import Fcl_checker
        ef, rf = str(tmp_path / "env.stl"), str(tmp_path / "robot.stl")
        meshio.write_stl(ef, meshio.shipped_mesh("env-scene-ltu-experiment"))
        meshio.write_stl(rf, meshio.shipped_mesh("custom_triangle_robot"))
        checker = Fcl_checker(ef, rf)
        rng = np.random.default_rng(99)
        states = _random_poses(rng, 2000, 4)
        single = np.zeros(2000, np.uint8)
        for i, st in enumerate(states):
            q = co.yaw_pose_quat(st[3])
            single[i] = checker.check_collision(st[:3], q)
        batch = checker.check_collision_batch(states)
        assert np.array_equal(single, batch)
        ref, margin = co.collide_poses(_soup("custom_triangle_robot"), _soup("env-scene-ltu-experiment"), states,
                                       with_margin=True)
        clear = np.abs(margin) > EPS
        assert np.array_equal(single[clear], ref[clear]) and 0.02 < single.mean() < 0.9
    finally:
        sys.path.remove(mst.dropin_path())
        for name in [m for m in sys.modules if m.split(".")[0] == "RigidBodyPlanners"]:
            del sys.modules[name]


def test_batched_motion_validation():
    """mst_collide_motions == checking every interpolated state separately, and == the oracle on
    the interpolated states (C restatement) outside the touching band."""
    import drone_path_planning_python_b200 as mst
    rng = np.random.default_rng(17)
    robot_tris, env_tris = _soup("custom_triangle_robot"), _soup("env-scene-ltu-experiment")
    robot, env = mst.Mesh(robot_tris), mst.Mesh(env_tris)
    M, steps = 5000, 37
    a = _random_poses(rng, M, 4)
    b = a + rng.normal(0, 0.4, (M, 4))
    invalid = mst.collide_motions(robot, env, a, b, steps).cpu().numpy()
    f = (np.arange(1, steps + 1) / steps)[None, :, None]
    states = a[:, None, :] + (b - a)[:, None, :] * f
    each = mst.collide_poses(robot, env, states.reshape(-1, 4)).cpu().numpy().reshape(M, steps)
    assert np.array_equal(invalid, each.max(axis=1))
    assert 0.05 < invalid.mean() < 0.95
    # against the oracle (not another kernel): the C restatement on the same interpolated states of a
    # subset of the motions; a motion may differ only if one of its states sits in the touching band
    from oracle import build_oracle, collision_oracle as co
    sub = slice(0, 600)
    ref_each = build_oracle.c_collide_poses(robot_tris, env_tris, states[sub].reshape(-1, 4)).reshape(-1, steps)
    ref_invalid = ref_each.max(axis=1)
    for m in np.flatnonzero(ref_invalid != invalid[sub]):
        _, margin = co.collide_poses(robot_tris, env_tris, states[m], with_margin=True)
        assert np.abs(margin).min() <= 1e-7, m
    assert (ref_invalid != invalid[sub]).mean() < 0.01
    # the planner's own straight line start -> goal crosses the wall (scripts/rigidBodyPath.py:146-147)
    assert mst.collide_motions(robot, env, [[0, 3, 1, 0]], [[0, 5, 1, 0]], 2000).cpu().numpy().tolist() == [1]
    assert mst.collide_motions(robot, env, [[0, 3, 1, 0]], [[1, 3.2, 1.2, 0.5]], 2000).cpu().numpy().tolist() == [0]
