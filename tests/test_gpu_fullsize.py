"""GPU tier, BASELINE.json full sizes (configs[1]..[4]): the oracle cannot reach these sizes in
seconds, so correctness is checked through size-independent properties of the reference's
system — interpolation, C^1..C^6 continuity, rest-to-rest ends, linearity in the waypoints,
time-group sharing, flag consistency — plus oracle spot checks on a seeded subset."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

LO = np.array([-2.2, 2.8, 0.5])
HI = np.array([2.2, 5.0, 2.5])
FACT = np.array([1, 1, 2, 6, 24, 120, 720, 5040], dtype=np.float64)


def _soup(name):
    from drone_path_planning_python_b200 import meshio
    verts, _, tris = meshio.ingest_mesh(meshio.shipped_mesh(name))
    return meshio.triangle_soup(verts, tris)


def _derivs_at(coef, tau):
    """All 8 derivatives of every piece at local time tau: coef [..., 8], tau [...] -> [..., 8]."""
    out = []
    c = coef
    for j in range(8):
        k = torch.arange(j, 8, device=coef.device, dtype=torch.float64)
        ff = torch.tensor([np.prod(np.arange(kk - j + 1, kk + 1)) if j else 1.0 for kk in range(j, 8)],
                          device=coef.device, dtype=torch.float64)
        out.append((c[..., j:] * ff * tau[..., None] ** (k - j)).sum(-1))
    return torch.stack(out, dim=-1)


def _check_system(coef, dur, wp, rtol=2e-9):
    """The reference's constraint rows (calculatingTrajectories.py:65-128), evaluated on the
    solution: residuals relative to the size of the terms involved."""
    B, n, K, _ = coef.shape
    T = dur[:, :, None].expand(B, n, K)
    end = _derivs_at(coef, T)                         # [B, n, K, 8] derivatives at piece end
    start = coef * torch.tensor(FACT, device=coef.device)   # derivatives at local time 0
    scale = (coef.abs() * T[..., None] ** torch.arange(8, device=coef.device)).amax(dim=(1, 3), keepdim=True)
    wpt = wp.transpose(1, 2) if False else wp        # [B, n+1, K]
    # interpolation at both ends of every piece
    assert ((start[..., 0] - wpt[:, :-1]).abs() <= rtol * scale[..., 0]).all()
    assert ((end[..., 0] - wpt[:, 1:]).abs() <= rtol * scale[..., 0]).all()
    # C^1..C^6 across interior knots, scaled per derivative by T^j
    for j in range(1, 7):
        jump = (end[:, :-1, :, j] - start[:, 1:, :, j]).abs() * T[:, :-1] ** j
        assert (jump <= 50 * rtol * scale[..., 0]).all(), j
    # rest to rest: velocity, acceleration, jerk vanish at both ends
    for j in range(1, 4):
        assert (start[:, 0, :, j].abs() * T[:, 0] ** j <= rtol * scale[:, 0, :, 0]).all()
        assert (end[:, -1, :, j].abs() * T[:, -1] ** j <= 50 * rtol * scale[:, 0, :, 0]).all()


def test_config1_shipped_example_with_third_drone(golden_dir):
    """BASELINE configs[0]: the shipped 3-drone formation example.  Rigid-body
    path recovered from Pol_matrix_{1,2}.csv (50 waypoints, T = 0.2, n = 49, K = 4); drones 1, 2 at
    (+-0.5, 0, 0) and the synthetic third drone at rb + R_z(yaw) (0, 0, -0.5) that the robot mesh
    implies; formation transform, min-snap solve with the formation sharing one time vector, float32
    packing and sampling on the GPU — against the oracle (restatement pinned bit for bit to the
    reference) in coefficient and position space, and against the shipped CSVs in position space."""
    import os
    from oracle import collision_oracle as co, minsnap_oracle as mo
    import drone_path_planning_python_b200 as mst
    with np.load(os.path.join(golden_dir, "shipped_pol_matrices.npz")) as z:
        m1, m2 = z["Pol_matrix_1"].astype(np.float64), z["Pol_matrix_2"].astype(np.float64)

    def waypoints(mat):
        c = mat[:, 1:].reshape(-1, 4, 8)
        w = np.zeros((50, 4))
        w[:49] = c[:, :, 0]
        w[49] = [mo.horner(c[48, k], mat[48, 0]) for k in range(4)]
        return w
    w1, w2 = waypoints(m1), waypoints(m2)
    rb = np.concatenate([0.5 * (w1[:, :3] + w2[:, :3]), w1[:, 3:4]], axis=1)[None]       # [1, 50, 4]
    offsets = np.array([[0.5, 0, 0], [-0.5, 0, 0], [0, 0, -0.5]])
    t = mo.uniform_times(50)                                                              # drones_pols_generator.py:44-56
    wp = mst.formation_waypoints(rb, offsets, K=4)
    ref_wp = mo.formation_waypoints(rb[0], offsets)
    assert np.abs(wp.cpu().numpy() - ref_wp).max() <= 1e-14
    assert np.abs(wp[0].cpu().numpy()[:, :3] - w1[:, :3]).max() < 2e-6 and np.abs(wp[1].cpu().numpy()[:, :3] - w2[:, :3]).max() < 2e-6
    # the third drone hangs 0.5 m below the rigid body whatever the yaw
    assert np.abs(wp[2].cpu().numpy()[:, :3] - (rb[0, :, :3] + [0, 0, -0.5])).max() <= 1e-14
    for solver in ("auto", "banded_lu"):
        coef, dur, info = mst.solve_batch(wp, t[None], share_time_group=3, solver=solver)
        assert int((info != 0).sum()) == 0
        ts = np.arange(0, 9.8, 0.1)
        pos = mst.sample_batch(coef, dur, ts=ts).cpu().numpy()
        for d in range(3):
            ref, rdur = mo.solve_waypoints(ref_wp[d], t)
            got = coef[d].cpu().numpy()
            # cond(A) ~ 2.7e8 at T = 0.2: normwise 1e-9 still holds for both solvers
            assert (np.abs(got - ref).max(axis=(0, 2)) / np.abs(ref).max(axis=(0, 2))).max() <= 1e-9, (solver, d)
            ref_pos = mo.sample_trajectory(ref, rdur, ts)
            assert np.abs(pos[d] - ref_pos).max() <= 1e-9 * max(1.0, np.abs(ref_pos).max()), (solver, d)
        packed = mst.pack_pol_matrix(coef, dur).cpu().numpy()
        assert packed.shape == (3, 49, 33) and packed.dtype == np.float32
        for d, mat in ((0, m1), (1, m2)):
            ours = mst.sample_batch(packed[d][None, :, 1:].reshape(1, 49, 4, 8).astype(np.float64),
                                    packed[d][None, :, 0].astype(np.float64), ts=ts).cpu().numpy()[0]
            theirs = mst.sample_batch(mat[None, :, 1:].reshape(1, 49, 4, 8), mat[None, :, 0], ts=ts).cpu().numpy()[0]
            assert np.abs(ours[:, :3] - theirs[:, :3]).max() < 5e-6, (solver, d)
    # the planned rigid-body states are collision-free for the 3-drone robot mesh (FCL accepted them)
    robot, env = mst.Mesh(_soup("custom_triangle_robot")), mst.Mesh(_soup("env-scene-ltu-experiment"))
    assert int(mst.collide_poses(robot, env, rb[0]).sum()) == 0
    # and the whole formation trajectory, sampled: pipeline on the rigid body's own trajectory
    res = mst.pipeline(rb, t[None], 100, robot, env)
    assert int(res.info[0]) == 0 and int(res.any_hit[0]) == 0


def test_config2_formations_4096x5():
    """4,096 formations x 5 drones x 10 pieces x 3 axes: one time vector per formation."""
    from oracle import minsnap_oracle as mo
    import drone_path_planning_python_b200 as mst
    rng = np.random.default_rng(20261018)
    F, D, n, K = 4096, 5, 10, 3
    offsets = np.array([[0.5, 0, 0], [-0.5, 0, 0], [0, 0, -0.5], [0, 0.5, 0], [0, -0.5, 0]])
    rb = np.zeros((F, n + 1, 4))
    rb[:, 0, :3] = rng.uniform(LO, HI, (F, 3))
    rb[:, 1:, :3] = rng.normal(0, 0.3, (F, n, 3))
    rb[:, :, :3] = np.cumsum(rb[:, :, :3], axis=1)
    rb[:, :, 3] = np.cumsum(rng.normal(0, 0.1, (F, n + 1)), axis=1)
    T = rng.uniform(0.5, 2.0, (F, n))
    t = np.concatenate([np.zeros((F, 1)), np.cumsum(T, axis=1)], axis=1)
    wp = mst.formation_waypoints(rb, offsets, K=K)
    assert wp.shape == (F * D, n + 1, K)
    # rigid formation: pairwise drone distances are constant along the path
    w = wp.view(F, D, n + 1, K)
    d01 = (w[:, 0] - w[:, 1]).norm(dim=-1)
    assert torch.allclose(d01, torch.ones_like(d01), atol=1e-12)
    coef, dur, info = mst.solve_batch(wp, t, share_time_group=D)
    assert int((info != 0).sum()) == 0
    assert torch.equal(dur.view(F, D, n)[:, 0], dur.view(F, D, n)[:, 4])
    _check_system(coef, dur, wp)
    # sharing the factorisation must not change the answer: solve drone 3 alone
    alone, _, _ = mst.solve_batch(w[:, 3].contiguous(), t)
    assert torch.equal(alone, coef.view(F, D, n, K, 8)[:, 3])
    for f in (0, 1234, F - 1):
        ref, _ = mo.solve_waypoints(w[f, 2].cpu().numpy(), t[f])
        got = coef.view(F, D, n, K, 8)[f, 2].cpu().numpy()
        assert (np.abs(got - ref).max(axis=(0, 2)) / np.abs(ref).max(axis=(0, 2))).max() <= 1e-9


def test_config3_free_time_resolves_65536x20():
    """65,536 problems x 20 pieces, re-solved with perturbed time allocations
    T_r = T_0 * exp(sigma_r * xi), sigma_r = 0.25 r, clipped to [0.05, 5] s (ill-conditioned stress)."""
    from oracle import minsnap_oracle as mo
    import drone_path_planning_python_b200 as mst
    rng = np.random.default_rng(3)
    B, n, K = 65536, 20, 3
    wp_np = np.cumsum(rng.normal(0, 0.3, (B, n + 1, K)), axis=1)
    wp = torch.as_tensor(wp_np, device="cuda")
    T0 = rng.uniform(0.5, 2.0, (B, n))
    xi = rng.normal(size=(B, n))
    report = []
    for r in (0, 2, 4, 7):
        T = np.clip(T0 * np.exp(0.25 * r * xi), 0.05, 5.0)
        t = np.concatenate([np.zeros((B, 1)), np.cumsum(T, axis=1)], axis=1)
        coef, dur, info = mst.solve_batch(wp, t)
        assert int((info != 0).sum()) == 0
        spread = T.max(axis=1) / T.min(axis=1)
        _check_system(coef, dur, wp, rtol=1e-8 if r else 2e-9)
        worst = 0.0
        for b in range(0, B, B // 8):
            ref, _ = mo.solve_waypoints(wp_np[b], t[b])
            got = coef[b].cpu().numpy()
            worst = max(worst, float((np.abs(got - ref).max(axis=(0, 2)) / np.abs(ref).max(axis=(0, 2))).max()))
        assert worst <= 1e-9, (r, worst)
        report.append((r, float((spread > 4).mean()), float(spread.max()), worst))
    print("config 3 (re-solve r, share on pivoted path, max spread, worst normwise error):", report)
    assert report[0][1] == 0.0 and report[-1][1] > 0.9


def test_linearity_and_translation_invariance_1M():
    import drone_path_planning_python_b200 as mst
    rng = np.random.default_rng(8)
    B, n, K = 1 << 20, 10, 3
    gen = torch.Generator(device="cuda").manual_seed(5)
    w1 = torch.randn((B, n + 1, K), generator=gen, device="cuda", dtype=torch.float64).cumsum(1)
    w2 = torch.randn((B, n + 1, K), generator=gen, device="cuda", dtype=torch.float64).cumsum(1)
    T = torch.rand((B, n), generator=gen, device="cuda", dtype=torch.float64) * 1.5 + 0.5
    t = torch.cat([torch.zeros((B, 1), device="cuda", dtype=torch.float64), T.cumsum(1)], dim=1)
    c1, dur, i1 = mst.solve_batch(w1, t)
    c2, _, _ = mst.solve_batch(w2, t)
    c3, _, _ = mst.solve_batch(2.0 * w1 - 0.5 * w2, t)
    assert int((i1 != 0).sum()) == 0
    scale = c3.abs().amax(dim=(1, 3), keepdim=True)
    assert ((c3 - (2.0 * c1 - 0.5 * c2)).abs() <= 1e-9 * scale).all()
    shift = torch.tensor([100.0, -50.0, 7.0], device="cuda", dtype=torch.float64)
    c4, _, _ = mst.solve_batch(w1 + shift, t)
    c4[..., 0] -= shift[None, None, :]
    s1 = c1.abs().amax(dim=(1, 3), keepdim=True)
    assert ((c4 - c1).abs() <= 1e-9 * s1).all()     # the solver only sees waypoint differences
    _check_system(c1[:: 64], dur[:: 64], w1[:: 64])
    del rng


@pytest.mark.parametrize("env_name", ["env-scene-ltu-experiment", "env-scene-narrow", "env-scene-hole"])
def test_config4_one_million_poses(env_name):
    """1,048,576 poses (x, y, z, yaw) vs each obstacle mesh: oracle parity on a 32,768 subset,
    permutation invariance and batch-split invariance on the whole set."""
    from oracle import build_oracle, collision_oracle as co
    import drone_path_planning_python_b200 as mst
    rng = np.random.default_rng(len(env_name))
    robot_tris, env_tris = _soup("custom_triangle_robot"), _soup(env_name)
    robot, env = mst.Mesh(robot_tris), mst.Mesh(env_tris)
    P = 1 << 20
    flat = env_tris.reshape(-1, 3)
    poses = np.concatenate([rng.uniform(flat.min(0) - 0.8, flat.max(0) + 0.8, (P, 3)),
                            rng.uniform(-np.pi, np.pi, (P, 1))], axis=1)
    dev = torch.as_tensor(poses, device="cuda")
    hit = mst.collide_poses(robot, env, dev)
    perm = torch.randperm(P, device="cuda")
    assert torch.equal(mst.collide_poses(robot, env, dev[perm]), hit[perm])
    halves = torch.cat([mst.collide_poses(robot, env, dev[: P // 3]), mst.collide_poses(robot, env, dev[P // 3:])])
    assert torch.equal(halves, hit)
    sub = slice(0, 32768)
    ref, margin = co.collide_poses(robot_tris, env_tris, poses[sub], with_margin=True)
    assert np.array_equal(build_oracle.c_collide_poses(robot_tris, env_tris, poses[sub]), ref)
    clear = np.abs(margin) > 1e-9
    assert clear.mean() > 0.999
    assert np.array_equal(hit[sub].cpu().numpy()[clear], ref[clear])
    print(env_name, "hit rate %.3f" % float(hit.float().mean()))


def test_config5_pipeline_one_million():
    """1,048,576 trajectories end to end: flag consistency, agreement with the separate stages on
    a slice, oracle spot checks, solver status."""
    from oracle import collision_oracle as co, minsnap_oracle as mo
    import drone_path_planning_python_b200 as mst
    rng = np.random.default_rng(20261019)
    B, n, K, S = 1 << 20, 10, 3, 100
    T = rng.uniform(0.5, 2.0, (B, n))
    t = np.concatenate([np.zeros((B, 1)), np.cumsum(T, axis=1)], axis=1)
    wp = rng.normal(0.0, 0.3, (B, n + 1, K))
    wp[:, 0, :] = rng.uniform(LO, HI, (B, K))
    np.cumsum(wp, axis=1, out=wp)
    robot_tris, env_tris = _soup("custom_triangle_robot"), _soup("env-scene-ltu-experiment")
    robot, env = mst.Mesh(robot_tris), mst.Mesh(env_tris)
    res = mst.pipeline(wp, t, S, robot, env)
    assert int((res.info != 0).sum()) == 0
    assert torch.equal(res.any_hit, res.hit.amax(dim=1))
    assert 0.3 < float(res.any_hit.float().mean()) < 0.9
    # a trajectory starts and ends at rest on its first / last waypoint
    first = res.coef[:, 0, :, 0]
    assert torch.equal(first.cpu(), torch.as_tensor(wp[:, 0, :]))
    sl = slice(500000, 500000 + 4096)
    pos = mst.sample_batch(res.coef[sl], res.dur[sl], S=S)
    hit = mst.collide_poses(robot, env, pos.reshape(-1, K)).reshape(-1, S)
    assert torch.equal(hit, res.hit[sl])
    again = mst.pipeline(wp, t, S, robot, env)
    assert torch.equal(again.hit, res.hit) and torch.equal(again.coef, res.coef)     # deterministic
    # ---- the oracle on 4,096 trajectories spread over the batch (process pool; C collision restatement)
    from _oracle_pool import oracle_pipeline
    pick = np.unique(np.concatenate([[0, 777777, B - 1], rng.choice(B, 4096, replace=False)]))
    ref_coef, ref_pos, ref_hit = oracle_pipeline(wp[pick], t[pick], S, robot_tris, env_tris)
    got = res.coef[torch.as_tensor(pick, device=res.coef.device)].cpu().numpy()
    err = np.abs(got - ref_coef).max(axis=(1, 3)) / np.abs(ref_coef).max(axis=(1, 3))      # per trajectory and axis
    assert err.max() <= 1e-9, err.max()
    hits = res.hit[torch.as_tensor(pick, device=res.hit.device)].cpu().numpy()
    differ = np.argwhere(hits != ref_hit)
    # flags are exact except for samples within EPS of touching (north star): the few that differ
    # must sit inside the band, measured with the numpy restatement's margin
    assert len(differ) <= 1e-4 * hits.size, len(differ)
    for q, s_ in differ:
        pose = np.concatenate([ref_pos[q, s_, :3], [0.0]])[None]
        _, margin = co.collide_poses(robot_tris, env_tris, pose, with_margin=True)
        assert abs(margin[0]) <= 1e-7, (int(pick[q]), int(s_), float(margin[0]))
    print("config 5: %d trajectories vs the oracle, worst coefficient error %.2e, %d of %d flags inside the band"
          % (len(pick), err.max(), len(differ), hits.size))
