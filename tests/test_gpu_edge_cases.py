"""GPU tier: edge cases — empty batches, single piece / single sample, long trajectories, the
brute-force collision path for robots the bit-mask engine cannot hold, empty meshes, non-finite
inputs, ragged formation groups."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _soup(name):
    from drone_path_planning_python_b200 import meshio
    verts, _, tris = meshio.ingest_mesh(meshio.shipped_mesh(name))
    return meshio.triangle_soup(verts, tris)


def _subdivide(tris, levels):
    for _ in range(levels):
        a, b, c = tris[:, 0], tris[:, 1], tris[:, 2]
        ab, bc, ca = (a + b) / 2, (b + c) / 2, (c + a) / 2
        tris = np.concatenate([np.stack([a, ab, ca], 1), np.stack([ab, b, bc], 1),
                               np.stack([ca, bc, c], 1), np.stack([ab, bc, ca], 1)])
    return tris


def test_empty_batches():
    import drone_path_planning_python_b200 as mst
    robot, env = mst.Mesh(_soup("custom_triangle_robot")), mst.Mesh(_soup("env-scene-ltu-experiment"))
    coef, dur, info = mst.solve_batch(np.zeros((0, 11, 3)), np.zeros((0, 11)))
    assert coef.shape == (0, 10, 3, 8) and dur.shape == (0, 10) and info.shape == (0,)
    assert mst.sample_batch(coef, dur, S=5).shape == (0, 5, 3)
    assert mst.collide_poses(robot, env, np.zeros((0, 4))).shape == (0,)
    res = mst.pipeline(np.zeros((0, 11, 3)), np.zeros((0, 11)), 100, robot, env)
    assert res.hit.shape == (0, 100) and res.any_hit.shape == (0,)
    with pytest.raises(ValueError):
        mst.solve_batch(np.zeros((5, 11, 3)), np.zeros((2, 11)), share_time_group=2)   # 5 % 2 != 0


@pytest.mark.parametrize("n,K", [(1, 1), (1, 4), (2, 2), (100, 3), (300, 1)])
def test_extreme_piece_counts(n, K):
    from oracle import minsnap_oracle as mo
    import drone_path_planning_python_b200 as mst
    rng = np.random.default_rng(n * 10 + K)
    B = 3
    T = rng.uniform(0.5, 2.0, (B, n))
    t = np.concatenate([np.zeros((B, 1)), np.cumsum(T, axis=1)], axis=1)
    wp = np.cumsum(rng.normal(0, 0.3, (B, n + 1, K)), axis=1)
    for solver in ("auto", "banded_lu"):
        coef, dur, info = mst.solve_batch(wp, t, solver=solver)
        assert (info.cpu().numpy() == 0).all(), solver
        ref, _ = mo.solve_waypoints(wp[0], t[0])
        got = coef[0].cpu().numpy()
        assert (np.abs(got - ref).max(axis=(0, 2)) / np.abs(ref).max(axis=(0, 2))).max() <= 1e-9, solver
    pos = mst.sample_batch(coef, dur, S=1)
    assert torch.equal(pos[:, 0, :].cpu(), torch.as_tensor(wp[:, 0, :]))      # t = 0 is the first waypoint


def test_pivoted_solver_reports_oversize():
    import drone_path_planning_python_b200 as mst
    from drone_path_planning_python_b200._abi import MstError
    from oracle import minsnap_oracle as mo
    n = 400                     # fits since only a window of the band is kept on chip (banded_core.cuh)
    rng = np.random.default_rng(400)
    t = np.concatenate([[0.0], np.cumsum(rng.uniform(0.3, 3.0, n))])[None]
    wp = np.cumsum(rng.normal(0, 0.3, (1, n + 1, 3)), axis=1)
    coef, _, info = mst.solve_batch(wp, t, solver="banded_lu")
    assert int(info[0]) == 0
    ref, _ = mo.solve_waypoints(wp[0], t[0])
    assert (np.abs(coef[0].cpu().numpy() - ref).max(axis=(0, 2)) / np.abs(ref).max(axis=(0, 2))).max() <= 1e-9
    n = 1000                    # window + power table + 3 right-hand sides of 8000 doubles do not fit one CTA
    t = np.arange(n + 1, dtype=np.float64)[None]
    with pytest.raises(MstError, match="does not fit"):
        mst.solve_batch(np.zeros((1, n + 1, 3)), t, solver="banded_lu")


def test_nonfinite_waypoints_propagate_without_hanging():
    import drone_path_planning_python_b200 as mst
    robot, env = mst.Mesh(_soup("custom_triangle_robot")), mst.Mesh(_soup("env-scene-ltu-experiment"))
    wp = np.zeros((4, 11, 3))
    wp[:, :, 1] = 4.0
    wp[1, 3, 0] = np.nan
    wp[2, 5, 2] = np.inf
    t = np.tile(np.arange(11.0), (4, 1))
    res = mst.pipeline(wp, t, 50, robot, env)
    info = res.info.cpu().numpy()
    assert (info == 0).all()                       # the reference would also "solve" these into NaNs
    assert np.isnan(res.coef[1].cpu().numpy()).any()
    assert int(res.any_hit[0]) == 1 and int(res.any_hit[3]) == 1     # stationary inside the wall
    assert int(res.any_hit[1]) in (0, 1) and int((res.hit > 1).sum()) == 0


@pytest.mark.parametrize("dim", [3, 4, 7])
def test_brute_force_path_for_large_robot_mesh(dim):
    """A robot with 128 triangles / > 32 unique vertices exceeds the engine's bit masks: the kernels
    fall back to the per-lane test over all pairs; same answers as the oracle."""
    from oracle import build_oracle
    import drone_path_planning_python_b200 as mst
    rng = np.random.default_rng(dim)
    robot_tris = _subdivide(_soup("custom_triangle_robot"), 2)
    env_tris = _soup("env-scene-narrow")
    assert len(robot_tris) == 128 and len(np.unique(robot_tris.reshape(-1, 3), axis=0)) > 32
    robot, env = mst.Mesh(robot_tris), mst.Mesh(env_tris)
    P = 3000
    flat = env_tris.reshape(-1, 3)
    pos = rng.uniform(flat.min(0) - 0.6, flat.max(0) + 0.6, (P, 3))
    if dim == 3:
        poses, ref_poses = pos, np.concatenate([pos, np.zeros((P, 1))], axis=1)
    elif dim == 4:
        poses = ref_poses = np.concatenate([pos, rng.uniform(-np.pi, np.pi, (P, 1))], axis=1)
    else:
        q = rng.normal(size=(P, 4))
        poses = ref_poses = np.concatenate([pos, q / np.linalg.norm(q, axis=1, keepdims=True)], axis=1)
    hit = mst.collide_poses(robot, env, poses).cpu().numpy()
    ref = build_oracle.c_collide_poses(robot_tris, env_tris, ref_poses)
    assert (hit != ref).sum() <= 1 and 0.1 < ref.mean() < 0.9    # same SAT arithmetic: at most a touching-band flip
    if dim != 7:
        # and through the fused pipeline: stationary "trajectories" sitting on the poses
        K = dim
        wp = np.repeat(poses[:256, None, :], 3, axis=1)
        t = np.tile(np.array([0.0, 1.0, 2.0]), (256, 1))
        res = mst.pipeline(wp, t, 4, robot, env)
        assert np.array_equal(res.any_hit.cpu().numpy(), hit[:256])


def test_empty_meshes_never_collide():
    import drone_path_planning_python_b200 as mst
    robot = mst.Mesh(_soup("custom_triangle_robot"))
    nothing = mst.Mesh(np.zeros((0, 3, 3)))
    poses = np.random.default_rng(0).uniform(-3, 6, (500, 4))
    assert int(mst.collide_poses(robot, nothing, poses).sum()) == 0
    assert int(mst.collide_poses(nothing, robot, poses).sum()) == 0


def test_sample_count_not_multiple_of_32_and_single_trajectory():
    import drone_path_planning_python_b200 as mst
    rng = np.random.default_rng(4)
    robot, env = mst.Mesh(_soup("custom_triangle_robot")), mst.Mesh(_soup("env-scene-ltu-experiment"))
    for B, S in ((1, 7), (3, 33), (5, 1), (2, 257)):
        T = rng.uniform(0.5, 2.0, (B, 6))
        t = np.concatenate([np.zeros((B, 1)), np.cumsum(T, axis=1)], axis=1)
        wp = np.array([0.0, 3.0, 1.0]) + np.cumsum(rng.normal(0, 0.4, (B, 7, 3)), axis=1)
        res = mst.pipeline(wp, t, S, robot, env)
        pos = mst.sample_batch(res.coef, res.dur, S=S)
        hit = mst.collide_poses(robot, env, pos.reshape(-1, 3)).reshape(B, S)
        assert torch.equal(res.hit, hit) and torch.equal(res.any_hit, hit.amax(dim=1))


@pytest.mark.parametrize("n,S,K", [(1, 100, 3), (10, 1500, 3), (10, 1024, 4), (30, 64, 3), (300, 40, 3),
                                    (7, 31, 4), (12, 2, 3)])
def test_fused_piece_lookup_equals_sampler(n, S, K):
    """The fused kernel finds a sample's piece through per-trajectory tables (first-sample
    thresholds, and a piece-of-sample byte table when n <= 255 and S <= 1024; a bisection over
    the thresholds otherwise).  Whatever the path, its flags must equal "sample with
    mst_sample_batch, then collide" — including zero-length pieces, where several thresholds
    coincide, a zero total duration, and tiles of every size."""
    import drone_path_planning_python_b200 as mst
    rng = np.random.default_rng(1000 * n + S)
    robot, env = mst.Mesh(_soup("custom_triangle_robot")), mst.Mesh(_soup("env-scene-ltu-experiment"))
    B = 77
    dur = rng.uniform(0.1, 1.0, (B, n))
    dur[rng.random((B, n)) < 0.2] = 0.0          # zero-length pieces (leading, trailing, in a row)
    dur[3] = 0.0                                 # a trajectory of zero total duration
    dur[4, :] = 0.0
    dur[4, n // 2] = 0.7                         # everything in one piece
    coef = np.zeros((B, n, K, 8))
    coef[..., 0] = np.array([0.0, 3.5, 1.2, 0.3][:K]) + rng.normal(0, 0.6, (B, n, K))
    coef[..., 1] = rng.normal(0, 1.0, (B, n, K))
    coef[..., 2] = rng.normal(0, 0.5, (B, n, K))
    hit, any_hit = mst.collide_trajectories(coef, dur, S, robot, env)
    pos = mst.sample_batch(coef, dur, S=S)
    want = mst.collide_poses(robot, env, pos.reshape(B * S, K)).reshape(B, S)
    assert torch.equal(hit, want)
    assert torch.equal(any_hit, want.amax(dim=1))
    assert 0 < int(want.sum()) < B * S


def _terrain(side=36, seed=5):
    """A bumpy height field over x in [-3, 3], y in [2, 6]: 2 * (side - 1)^2 triangles (2,450 for side 36)."""
    rng = np.random.default_rng(seed)
    xs, ys = np.linspace(-3.0, 3.0, side), np.linspace(2.0, 6.0, side)
    z = 0.6 + 0.5 * np.sin(1.7 * xs)[:, None] * np.cos(1.3 * ys)[None, :] + 0.15 * rng.normal(size=(side, side))
    pts = np.stack([np.broadcast_to(xs[:, None], (side, side)), np.broadcast_to(ys[None, :], (side, side)), z], axis=-1)
    tris = []
    for i in range(side - 1):
        for j in range(side - 1):
            a, b, c, d = pts[i, j], pts[i + 1, j], pts[i + 1, j + 1], pts[i, j + 1]
            tris.append([a, b, c])
            tris.append([a, c, d])
    order = rng.permutation(len(tris))          # no spatial order in the input: the library sorts
    return np.asarray(tris)[order]


@pytest.mark.parametrize("dim", [3, 4, 7])
def test_large_environment_block_boxes_vs_oracle(dim):
    """An environment of 2,450 triangles (far beyond what a CTA can stage: the mesh image stays in device
    memory, the cursor walks the block boxes above the per-triangle boxes) against the C restatement."""
    import drone_path_planning_python_b200 as mst
    from oracle import build_oracle, collision_oracle as co
    rng = np.random.default_rng(60 + dim)
    env_tris = _terrain()
    assert len(env_tris) >= 2000
    robot_tris = _soup("custom_triangle_robot")
    robot, env = mst.Mesh(robot_tris), mst.Mesh(env_tris)
    P = 30000
    pos = rng.uniform([-3.3, 1.7, -0.3], [3.3, 6.3, 1.8], (P, 3))
    if dim == 3:
        poses, oracle_poses = pos, np.concatenate([pos, np.zeros((P, 1))], axis=1)
    elif dim == 4:
        poses = oracle_poses = np.concatenate([pos, rng.uniform(-np.pi, np.pi, (P, 1))], axis=1)
    else:
        q = rng.normal(size=(P, 4))
        q /= np.linalg.norm(q, axis=1, keepdims=True)
        poses = oracle_poses = np.concatenate([pos, q], axis=1)
    hit = mst.collide_poses(robot, env, poses).cpu().numpy()
    ref = build_oracle.c_collide_poses(robot_tris, env_tris, oracle_poses)
    differ = np.flatnonzero(hit != ref)
    assert 0.1 < ref.mean() < 0.9
    for i in differ:           # only inside the touching band
        _, margin = co.collide_poses(robot_tris, env_tris, oracle_poses[i:i + 1], with_margin=True)
        assert abs(margin[0]) <= 1e-9, i
    assert len(differ) <= 3
    # the single synchronous query and the motion validator read the same large mesh
    for i in range(0, 200, 7):
        assert mst.collide_pose_now(robot, env, poses[i]) == hit[i] or i in differ
    if dim == 4:
        a, b = poses[:500], poses[500:1000]
        invalid = mst.collide_motions(robot, env, a, b, 9).cpu().numpy()
        f = (np.arange(1, 10) / 9)[None, :, None]
        states = (a[:, None, :] + (b - a)[:, None, :] * f).reshape(-1, 4)
        each = mst.collide_poses(robot, env, states).cpu().numpy().reshape(500, 9)
        assert np.array_equal(invalid, each.max(axis=1))


def test_large_environment_through_the_pipeline():
    """Trajectories over the 2,450-triangle terrain: pipeline flags == sample, then collide_poses; the
    single-pass kernel declines meshes it cannot stage and the two-launch pipeline takes over."""
    import drone_path_planning_python_b200 as mst
    rng = np.random.default_rng(8)
    robot, env = mst.Mesh(_soup("custom_triangle_robot")), mst.Mesh(_terrain())
    for K in (3, 4):
        B, n, S = 400, 10, 100
        T = rng.uniform(0.5, 2.0, (B, n))
        t = np.concatenate([np.zeros((B, 1)), np.cumsum(T, axis=1)], axis=1)
        wp = np.zeros((B, n + 1, K))
        wp[:, :, :3] = rng.uniform([-2.5, 2.5, 0.3], [2.5, 5.5, 1.6], (B, 1, 3)) + np.cumsum(rng.normal(0, 0.25, (B, n + 1, 3)), axis=1)
        if K == 4:
            wp[:, :, 3] = np.cumsum(rng.normal(0, 0.2, (B, n + 1)), axis=1)
        for solver in ("auto", "auto_one_pass"):
            res = mst.pipeline(wp, t, S, robot, env, solver=solver)
            pos = mst.sample_batch(res.coef, res.dur, S=S)
            hit = mst.collide_poses(robot, env, pos.reshape(B * S, K)).reshape(B, S)
            assert torch.equal(res.hit, hit) and torch.equal(res.any_hit, hit.amax(dim=1))
        assert 0.05 < float(res.any_hit.float().mean()) < 1.0
