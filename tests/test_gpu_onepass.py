"""GPU tier: the single-pass pipeline kernel (csrc/pipeline_onepass.cu) against the separate
stages it fuses — mst_solve_batch, then mst_collide_trajectories (the two-launch kernels) — and
against the oracle.  Everything must be IDENTICAL to the separate stages: coefficients bit for
bit, durations, status, every flag; including tiles that mix groups the condensed solver takes
with groups it hands to the pivoted solver, bad stamps, shared time vectors, partial last tiles."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

LO = np.array([-2.2, 2.8, 0.5])
HI = np.array([2.2, 5.0, 2.5])


def _soup(name):
    from drone_path_planning_python_b200 import meshio
    verts, _, tris = meshio.ingest_mesh(meshio.shipped_mesh(name))
    return meshio.triangle_soup(verts, tris)


def _meshes():
    import drone_path_planning_python_b200 as mst
    return mst.Mesh(_soup("custom_triangle_robot")), mst.Mesh(_soup("env-scene-ltu-experiment"))


def _workload(rng, B, n, K, G=1, wide_every=0, bad=()):
    F = B // G
    T = rng.uniform(0.5, 2.0, (F, n))
    if wide_every:
        T[::wide_every] *= np.exp(rng.normal(0, 1.2, (len(T[::wide_every]), n)))   # spread > 4: pivoted solver
        T = np.clip(T, 0.05, 5.0)
    t = np.concatenate([np.zeros((F, 1)), np.cumsum(T, axis=1)], axis=1)
    for f, kind in bad:
        if kind == "decreasing":
            t[f, 3] = t[f, 2] - 0.1
        elif kind == "nan":
            t[f, 2] = np.nan
        elif kind == "t0":
            t[f] += 0.4
        elif kind == "zero":
            t[f, 4] = t[f, 3]
    wp = np.zeros((B, n + 1, K))
    wp[:, :, :3] = rng.uniform(LO, HI, (B, 1, 3)) + np.cumsum(rng.normal(0, 0.3, (B, n + 1, 3)), axis=1)
    if K == 4:
        wp[:, :, 3] = np.cumsum(rng.normal(0, 0.2, (B, n + 1)), axis=1)
    return wp, t


def _separate(mst, wp, t, S, robot, env, G):
    coef, dur, info = mst.solve_batch(wp, t, share_time_group=G)
    hit, any_hit = mst.collide_trajectories(coef, dur, S, robot, env)
    return coef, dur, info, hit, any_hit


def _same(a, b):
    return torch.equal(a.view(torch.int64) if a.dtype == torch.float64 else a,
                       b.view(torch.int64) if b.dtype == torch.float64 else b)


@pytest.mark.parametrize("K,G,n,S,B", [(3, 1, 10, 100, 1003), (4, 1, 10, 100, 517), (3, 5, 10, 100, 1000),
                                         (3, 1, 2, 32, 77), (3, 1, 32, 257, 41), (3, 1, 20, 100, 300), (3, 1, 10, 257, 41), (4, 2, 7, 64, 258),
                                         (3, 10, 4, 100, 60), (3, 1, 10, 4096, 13)])
def test_single_pass_equals_separate_stages(K, G, n, S, B):
    import drone_path_planning_python_b200 as mst
    rng = np.random.default_rng(1000 * K + 10 * G + n)
    robot, env = _meshes()
    wp, t = _workload(rng, B, n, K, G)
    res = mst.pipeline(wp, t, S, robot, env, share_time_group=G, solver="auto_one_pass")
    coef, dur, info, hit, any_hit = _separate(mst, wp, t, S, robot, env, G)
    assert int((info != 0).sum()) == 0
    assert _same(res.coef, coef) and _same(res.dur, dur) and torch.equal(res.info, info)
    assert torch.equal(res.hit, hit) and torch.equal(res.any_hit, any_hit)
    assert torch.equal(res.any_hit, res.hit.amax(dim=1))
    if S == 100 and n == 10:
        assert 0.0 < float(res.any_hit.float().mean()) < 1.0


@pytest.mark.parametrize("K,G", [(3, 1), (4, 1), (3, 3)])
def test_single_pass_with_groups_for_the_pivoted_solver(K, G):
    """Tiles that mix solvable groups with wide duration spreads (pivoted solver + list-mode
    sampling), the t[0] != 0 quirk, a zero-length piece (singular), decreasing and NaN stamps."""
    import drone_path_planning_python_b200 as mst
    rng = np.random.default_rng(77 + K + G)
    robot, env = _meshes()
    B, n, S = 600 * G, 10, 100
    bad = ((5, "decreasing"), (17, "nan"), (18, "t0"), (301, "zero"), (599, "decreasing"))
    wp, t = _workload(rng, B, n, K, G, wide_every=3, bad=bad)
    out = mst.PipelineResult(torch.full((B, n, K, 8), 7.0, dtype=torch.float64, device="cuda"),
                             torch.full((B, n), 7.0, dtype=torch.float64, device="cuda"),
                             torch.full((B,), 99, dtype=torch.int32, device="cuda"),
                             torch.full((B, S), 9, dtype=torch.uint8, device="cuda"),
                             torch.full((B,), 9, dtype=torch.uint8, device="cuda"))
    res = mst.pipeline(wp, t, S, robot, env, share_time_group=G, out=out, solver="auto_one_pass")
    coef, dur, info, hit, any_hit = _separate(mst, wp, t, S, robot, env, G)
    assert torch.equal(res.info, info)
    codes = info.view(-1, G)[:, 0].cpu().numpy()
    assert codes[5] == -1 and codes[17] == -2 and codes[301] > 0 and codes[18] == 0
    nan_safe = lambda a, b: torch.equal(torch.nan_to_num(a, nan=-1.0).view(torch.int64), torch.nan_to_num(b, nan=-1.0).view(torch.int64))
    assert nan_safe(res.coef, coef) and nan_safe(res.dur, dur)
    assert torch.equal(res.hit, hit) and torch.equal(res.any_hit, any_hit)
    assert int((res.hit > 1).sum()) == 0 and int((res.any_hit > 1).sum()) == 0      # every flag written


def test_single_pass_vs_oracle():
    """The fused kernel against the oracle directly (not through another kernel): coefficients to
    1e-9 normwise, flags exact outside the touching band."""
    import drone_path_planning_python_b200 as mst
    from _oracle_pool import oracle_pipeline
    from oracle import collision_oracle as co
    rng = np.random.default_rng(4242)
    robot_tris, env_tris = _soup("custom_triangle_robot"), _soup("env-scene-ltu-experiment")
    robot, env = mst.Mesh(robot_tris), mst.Mesh(env_tris)
    for K in (3, 4):
        B, n, S = 512, 10, 100
        wp, t = _workload(rng, B, n, K)
        res = mst.pipeline(wp, t, S, robot, env, solver="auto_one_pass")
        ref_coef, ref_pos, ref_hit = oracle_pipeline(wp, t, S, robot_tris, env_tris)
        got = res.coef.cpu().numpy()
        err = np.abs(got - ref_coef).max(axis=(1, 3)) / np.abs(ref_coef).max(axis=(1, 3))
        assert err.max() <= 1e-9
        hits = res.hit.cpu().numpy()
        differ = np.argwhere(hits != ref_hit)
        assert len(differ) <= 1e-4 * hits.size
        for q, s_ in differ:
            pose = ref_pos[q, s_] if K == 4 else np.concatenate([ref_pos[q, s_, :3], [0.0]])
            _, margin = co.collide_poses(robot_tris, env_tris, pose[None], with_margin=True)
            assert abs(margin[0]) <= 1e-7


@pytest.mark.parametrize("K,G", [(3, 1), (4, 1), (3, 2)])
def test_wire_outputs_equal_packed_results(K, G):
    """mst_pipeline_wire on one GPU with three destination buffers standing in for the gather
    buffers of three ranks: every destination receives, at row_offset, the float32 polynomial
    matrix (== mst_pack_pol_matrix of the local results) and the flags — also for the groups the
    pivoted solver finishes (wire patch)."""
    import drone_path_planning_python_b200 as mst
    rng = np.random.default_rng(5 + K + G)
    robot, env = _meshes()
    B, n, S = 402 * G, 10, 100
    wp, t = _workload(rng, B, n, K, G, wide_every=7)
    rows, off = B + 100, 60
    mats = [torch.full((rows, n, 1 + 8 * K), -5.0, dtype=torch.float32, device="cuda") for _ in range(3)]
    hits = [torch.full((rows, S), 9, dtype=torch.uint8, device="cuda") for _ in range(3)]
    anys = [torch.full((rows,), 9, dtype=torch.uint8, device="cuda") for _ in range(3)]
    wire = mst.make_wire_targets(mats, hits, anys, row_offset=off)
    res = mst.pipeline_wire(wp, t, S, robot, env, wire, share_time_group=G)
    plain = mst.pipeline(wp, t, S, robot, env, share_time_group=G)
    assert _same(res.coef, plain.coef) and torch.equal(res.hit, plain.hit) and torch.equal(res.any_hit, plain.any_hit)
    packed = mst.pack_pol_matrix(res.coef, res.dur)
    for m, h, a in zip(mats, hits, anys):
        assert torch.equal(m[off:off + B], packed)
        assert torch.equal(h[off:off + B], res.hit) and torch.equal(a[off:off + B], res.any_hit)
        assert float(m[:off].max()) == -5.0 and float(m[off + B:].min()) == -5.0       # nothing outside this rank's rows
        assert int(h[:off].min()) == 9 and int(a[off + B:].min()) == 9
    # flags-only gather
    hits2 = [torch.full((rows, S), 9, dtype=torch.uint8, device="cuda")]
    anys2 = [torch.full((rows,), 9, dtype=torch.uint8, device="cuda")]
    res2 = mst.pipeline_wire(wp, t, S, robot, env, mst.make_wire_targets(None, hits2, anys2, row_offset=0), share_time_group=G)
    assert torch.equal(hits2[0][:B], res2.hit) and torch.equal(anys2[0][:B], res2.any_hit)


@pytest.mark.parametrize("K,G", [(3, 1), (4, 1), (3, 5)])
def test_pipeline_packed_matrix_equals_pack_of_results(K, G):
    """mst_pipeline_packed: the float32 polynomial matrix written by the solver kernel (and, for the groups
    the pivoted solver finishes, by the list-mode pass) == mst_pack_pol_matrix of the pipeline's results;
    everything else identical to the plain pipeline (far-piece culling included)."""
    import drone_path_planning_python_b200 as mst
    rng = np.random.default_rng(100 + K + G)
    robot, env = _meshes()
    B, n, S = 505 * G, 10, 100
    wp, t = _workload(rng, B, n, K, G, wide_every=6, bad=((3, "decreasing"), (40, "t0")))
    mat = torch.full((B, n, 1 + 8 * K), -7.0, dtype=torch.float32, device="cuda")
    res = mst.pipeline(wp, t, S, robot, env, share_time_group=G, pol_matrix=mat)
    plain = mst.pipeline(wp, t, S, robot, env, share_time_group=G)
    nan_safe = lambda a, b: torch.equal(torch.nan_to_num(a, nan=-1.0), torch.nan_to_num(b, nan=-1.0))
    assert nan_safe(res.coef, plain.coef) and torch.equal(res.hit, plain.hit) and torch.equal(res.info, plain.info)
    want = mst.pack_pol_matrix(res.coef, res.dur)
    ok = (res.info == 0)
    assert torch.equal(mat[ok], want[ok])
    assert nan_safe(mat, want)
    # and the culled two-launch pipeline against the separate stages (no culling in mst_collide_trajectories)
    coef, dur, info, hit, any_hit = _separate(mst, wp, t, S, robot, env, G)
    assert torch.equal(plain.hit[ok], hit[ok]) and torch.equal(plain.any_hit[ok], any_hit[ok])
