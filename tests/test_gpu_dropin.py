"""GPU tier: the reference's own Python call surface, served by the CUDA library.
These read like tests the reference would have: build Point_time lists, call
calculate_trajectory4D, evaluate polynomials, load CSVs, query the collision checker."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture()
def dropin():
    import drone_path_planning_python_b200 as mst
    path = mst.dropin_path()
    sys.path.insert(0, path)
    yield path
    sys.path.remove(path)
    for name in [m for m in sys.modules if m.split(".")[0] in ("optimizations", "RigidBodyPlanners", "scripts", "trajectory_visualising")]:
        del sys.modules[name]


def _load(golden_dir, name):
    with np.load(os.path.join(golden_dir, name)) as z:
        return {k: z[k] for k in z.files}


def test_calculate_trajectory4d_matches_reference_output(dropin, golden_dir):
    import optimizations as o
    from optimizations.calculatingTrajectories import calculate_trajectory1D
    z = _load(golden_dir, "solve_cases.npz")
    for case in ("reference_test_data", "rand_n10_r0", "uniform_n49", "stress_n20_r3", "t0_nonzero_n6"):
        wp, t, ref = z[case + "__wp"], z[case + "__t"], z[case + "__coef"]
        K = wp.shape[1]
        pts = [o.Point_time(o.Waypoint(*(list(wp[i]) + [0.0] * (4 - K))), t=float(t[i])) for i in range(len(t))]
        pols_coeffs, pc_pols = o.calculate_trajectory4D(pts)
        assert len(pols_coeffs) == 4 and len(pc_pols) == 4
        n = len(t) - 1
        for k in range(K):
            assert len(pols_coeffs[k]) == n and pc_pols[k].nOfPols == n
            got = np.stack([p.p.reshape(8) for p in pols_coeffs[k]])
            assert pols_coeffs[k][0].p.shape == (8, 1) and pols_coeffs[k][0].p.dtype == np.float64
            assert np.abs(got - ref[:, k]).max() <= 1e-9 * np.abs(ref[:, k]).max(), (case, k)
            assert pc_pols[k].time_durations == [float(d) for d in z[case + "__dur"]]
            assert isinstance(pc_pols[k].time_durations, list)
        # consumers call .reshape((1, 8)) on p (scripts/drones_pols_generator.py:72)
        assert pols_coeffs[0][0].p.reshape((1, 8)).shape == (1, 8)
        pieces, total = calculate_trajectory1D(pts, o.Waypoint.WP_TYPE_Y)
        one = np.stack([p.p.reshape(8) for p in pieces])
        four = np.stack([p.p.reshape(8) for p in pols_coeffs[1]])
        # the same axis solved alone or with its siblings (kernel tiling may differ: last-bit level)
        assert np.abs(one - four).max() <= 1e-12 * np.abs(four).max()


def test_reference_exceptions(dropin):
    import optimizations as o
    mk = lambda ts: [o.Point_time(o.Waypoint(float(i), 0.0, 0.0, 0.0), t=float(t)) for i, t in enumerate(ts)]
    with pytest.raises(np.linalg.LinAlgError):
        o.calculate_trajectory4D(mk([0.0, 1.0, 1.0, 2.0]))
    with pytest.raises(AssertionError):
        o.calculate_trajectory4D(mk([0.0, 1.0, 0.5, 2.0]))
    with pytest.raises(IndexError):
        o.calculate_trajectory4D(mk([0.0]))


def test_polynomial_methods(dropin, golden_dir):
    import optimizations as o
    z = _load(golden_dir, "time_power_rows.npz")
    for a, t in enumerate(z["t"]):
        pol = o.Polynomial([1, 1, 1, 1, 1, 1, 1, 1])
        for j in range(8):
            row = np.pad(np.array(pol.pol_coeffs_at_t(float(t))), (j, 0), "constant")
            assert np.array_equal(row, z["rows"][a, j]), (t, j)
            pol = pol.derivative()
        assert len(pol.p) == 0
    ev = _load(golden_dir, "piecewise_eval.npz")
    coef, dur = ev["coef"], ev["dur"]
    pols = [o.Polynomial(coef[i, 0].reshape(8, 1).copy()) for i in range(coef.shape[0])]
    pc = o.PiecewisePolynomial(pols, [float(d) for d in dur])
    for s in (0, 7, 20, 45, len(ev["t"]) - 1):
        v = pc.eval(float(ev["t"][s]))
        assert isinstance(v, np.ndarray) and v.shape == (1,)          # as the reference with (8,1) coefficients
        assert np.isclose(v[0], ev["values"][s, 0], rtol=1e-12, atol=0)
    inside = ev["t"] < dur.sum() - 1e-9
    assert np.array_equal(pc.eval_many(ev["t"])[inside], ev["values"][inside, 0])
    flat = o.Polynomial([1.0, -2.0, 0.5])
    assert flat.eval(2.0) == 1.0 - 4.0 + 2.0 and isinstance(flat.eval(2.0), float)
    assert flat.derivative().p == [-2.0, 1.0]


def test_trajectory_loadcsv_and_eval(dropin, golden_dir, tmp_path):
    import optimizations as o
    z = _load(golden_dir, "trajectory_eval.npz")
    rows = z["traj__file_rows"]
    path = tmp_path / "traj.csv"
    np.savetxt(str(path), rows, delimiter=",", header="duration,x^0,...", comments="")
    tr = o.Trajectory()
    tr.loadcsv(str(path))
    assert tr.n_pieces() == int(z["traj__n_pieces"]) and tr.duration == float(z["traj__duration"])
    for s in (0, 3, 31, len(z["traj__t"]) - 1):
        out = tr.eval(float(z["traj__t"][s]))
        assert np.allclose(out.pos, z["traj__pos"][s], rtol=1e-14, atol=1e-15)   # savetxt round trip
        assert np.allclose(out.vel, z["traj__vel"][s], rtol=1e-13, atol=1e-14)
        assert np.allclose(out.omega, z["traj__omega"][s], rtol=1e-10, atol=1e-12)
        assert np.isclose(out.yaw, z["traj__yaw"][s])
    many = tr.eval_many(z["traj__t"])
    assert np.allclose(many[:, 0:3], z["traj__pos"], rtol=1e-14, atol=1e-15)
    p4 = tr.polynomials[2]
    one = p4.eval(0.3)
    assert one.pos.shape == (3,) and one.acc.shape == (3,)
    with pytest.raises(AssertionError):
        tr.eval(-0.1)


def test_trajectory_visualising_nav_path(dropin, golden_dir, tmp_path):
    """get_nav_path_msg (src/trajectory_visualising/visualization.py:39-71) on the reference's own
    src/traj.csv: one pose per np.arange(0, duration, timestep) sample, position = Trajectory.eval(t).pos
    + offset, orientation = quaternion_from_euler(0, 0, -yaw) — against the unmodified reference's
    Trajectory.eval outputs stored in tests/golden/trajectory_eval.npz."""
    import math
    import trajectory_visualising as tv
    from trajectory_visualising import visualization
    z = _load(golden_dir, "trajectory_eval.npz")
    path = tmp_path / "traj.csv"
    np.savetxt(str(path), z["traj__file_rows"], delimiter=",", header="duration,x^0,...", comments="")
    tr = tv.Trajectory()
    tr.loadcsv(str(path))
    assert isinstance(tr.eval(0.0), tv.TrajectoryOutput)
    offset = [0.25, -1.0, 0.5]
    msg = tv.get_nav_path_msg(tr, 0.1, offset)
    ts = np.arange(0, tr.duration, 0.1)
    assert np.array_equal(ts, z["traj__t"][:-1]) and len(msg.poses) == len(ts)
    assert msg.header.frame_id == "world"
    for s, pose in enumerate(msg.poses):
        want = z["traj__pos"][s] + np.asarray(offset)
        got = np.array([pose.pose.position.x, pose.pose.position.y, pose.pose.position.z])
        assert np.array_equal(got, want), s                                   # bit-exact positions
        yaw = float(z["traj__yaw"][s])
        assert pose.pose.orientation.x == 0.0 and pose.pose.orientation.y == 0.0
        assert pose.pose.orientation.z == math.sin(-yaw / 2.0) and pose.pose.orientation.w == math.cos(-yaw / 2.0)
    # skiprows=1 quirk through this package too: a header-less polynomial matrix loses its first piece
    np.savetxt(str(path), z["pol1__file_rows"], delimiter=",")
    tr.loadcsv(str(path))
    assert tr.n_pieces() == int(z["pol1__n_pieces"]) == len(z["pol1__file_rows"]) - 1
    pos, quat = visualization.sample_path(tr, 0.1)
    assert np.array_equal(pos, z["pol1__pos"][:len(pos)])


def test_fcl_checker_dropin(dropin, tmp_path):
    from drone_path_planning_python_b200 import meshio
    from oracle import collision_oracle as co
    from RigidBodyPlanners.fcl_checker import Fcl_checker
    env_file, robot_file = tmp_path / "env.stl", tmp_path / "robot.stl"
    meshio.write_stl(str(env_file), meshio.shipped_mesh("env-scene-ltu-experiment"))
    meshio.write_stl(str(robot_file), meshio.shipped_mesh("custom_triangle_robot"))
    checker = Fcl_checker(str(env_file), str(robot_file))
    assert checker.robot.verts.shape == (6, 3) and checker.robot.tris.shape == (8, 3)
    assert checker.env.verts.shape == (8, 3) and checker.env.vecs.shape == (12, 3, 3)
    # isStateValid's calling sequence (RB_planning_sep_coll_check.py:208-215)
    for state, expected in (([0, 4, 1, 0.0], 1), ([0, 3, 1, 0.0], 0), ([0, 5, 1, 0.0], 0), ([0, 4, 2.17, 0.3], 0)):
        q = co.yaw_pose_quat(state[3])
        checker.set_robot_transform(state[:3], q)
        assert checker.check_collision() == expected
        assert checker.check_collision(state[:3], q) == expected
    rng = np.random.default_rng(0)
    poses = np.concatenate([rng.uniform([-2.5, 3, 0], [2.5, 5, 2.5], (500, 3)), rng.uniform(-3, 3, (500, 1))], axis=1)
    flags = checker.check_collision_batch(poses)
    robot_tris, env_tris = co.mesh_triangles(meshio.shipped_mesh("custom_triangle_robot")), \
        co.mesh_triangles(meshio.shipped_mesh("env-scene-ltu-experiment"))
    ref, margin = co.collide_poses(robot_tris, env_tris, poses, with_margin=True)
    clear = np.abs(margin) > 1e-9
    assert np.array_equal(flags[clear], ref[clear])


def test_node_scripts_reproduce_the_shipped_example(dropin, golden_dir, tmp_path, monkeypatch):
    """scripts pipeline on the shipped example: rigid-body path -> transform -> path_to_pol ->
    (49, 33) float32 matrices equal to the shipped CSVs in position space."""
    from oracle import minsnap_oracle as mo
    import drone_path_planning_python_b200 as mst
    from scripts import drones_pols_generator as pols, drones_traj_generator as gen
    from scripts._ros_compat import Path, make_pose
    z = _load(golden_dir, "shipped_pol_matrices.npz")
    m1, m2 = z["Pol_matrix_1"].astype(np.float64), z["Pol_matrix_2"].astype(np.float64)

    def waypoints(mat):
        c = mat[:, 1:].reshape(-1, 4, 8)
        wps = np.zeros((50, 4))
        wps[:49] = c[:, :, 0]
        wps[49] = [mo.horner(c[48, k], mat[48, 0]) for k in range(4)]
        return wps
    w1, w2 = waypoints(m1), waypoints(m2)
    rb_pos, yaw = 0.5 * (w1[:, :3] + w2[:, :3]), w1[:, 3]
    rb_path = Path()
    for i in range(50):
        rb_path.poses.append(make_pose(rb_pos[i, 0], rb_pos[i, 1], rb_pos[i, 2], mo.yaw_quat(yaw[i])))
    p1, p2 = gen.transform(rb_path)
    assert len(p1.poses) == 50
    got1 = np.array([[p.pose.position.x, p.pose.position.y, p.pose.position.z] for p in p1.poses])
    got2 = np.array([[p.pose.position.x, p.pose.position.y, p.pose.position.z] for p in p2.poses])
    assert np.abs(got1 - w1[:, :3]).max() < 2e-6 and np.abs(got2 - w2[:, :3]).max() < 2e-6   # float32 CSV inputs
    monkeypatch.setattr(pols, "OUTPUT_DIR", str(tmp_path))
    matrix, msg = pols.path_to_pol(p1, 1)
    assert matrix.shape == (49, 33) and matrix.dtype == np.float32
    assert msg.cf_id == 1 and len(msg.poly_x) == 49 * 8 and len(msg.durations) == 49
    written = np.loadtxt(str(tmp_path / "Pol_matrix_1.csv"), delimiter=",")
    assert np.array_equal(written.astype(np.float32), matrix)
    import io
    buf = io.BytesIO()
    np.savetxt(buf, matrix, delimiter=",")
    assert open(str(tmp_path / "Pol_matrix_1.csv"), "rb").read() == buf.getvalue()      # the bytes np.savetxt writes
    assert np.allclose(matrix[:, 0], 0.2, atol=1e-7)
    ts = np.linspace(0, 9.79, 300)
    ours = mst.sample_batch(matrix[None, :, 1:].reshape(1, 49, 4, 8).astype(np.float64),
                            matrix[None, :, 0].astype(np.float64), ts=ts).cpu().numpy()[0]
    theirs = mst.sample_batch(m1[None, :, 1:].reshape(1, 49, 4, 8), m1[None, :, 0], ts=ts).cpu().numpy()[0]
    assert np.abs(ours[:, :3] - theirs[:, :3]).max() < 5e-6
    assert np.abs(ours[:, 3] - theirs[:, 3]).max() < 5e-6
    many = pols.paths_to_matrices(np.stack([gen._path_array(p1), gen._path_array(p2)]))
    assert many.shape == (2, 49, 33) and np.array_equal(many[0], matrix)


def test_csv_emitter_is_byte_identical_to_savetxt(golden_dir, tmp_path):
    """mst_format_pol_matrix_csv against np.savetxt (scripts/drones_pols_generator.py:79-81 writes the
    float32 matrix with numpy's default '%.18e'), on the shipped matrices (whose first line as stored in
    the reference's own CSV files is a fixture), on random matrices with extreme magnitudes, and batched."""
    import io
    import drone_path_planning_python_b200 as mst
    z = _load(golden_dir, "shipped_pol_matrices.npz")

    def savetxt_bytes(mat):
        buf = io.BytesIO()
        np.savetxt(buf, mat, delimiter=",")
        return buf.getvalue()
    names = [k for k in z if not k.endswith("__first_line")]
    mats = np.stack([z[k] for k in names])                       # [6, 49, 33] float32
    blobs = mst.pol_matrix_csv(mats)
    for name, mat, blob in zip(names, mats, blobs):
        assert blob == savetxt_bytes(mat), name
        assert blob.split(b"\n")[0].decode() == str(z[name + "__first_line"]), name    # the reference's own file
    rng = np.random.default_rng(11)
    wild = (rng.normal(size=(5, 7, 25)) * 10.0 ** rng.integers(-40, 38, size=(5, 7, 25))).astype(np.float32)
    wild[0, 0, :6] = [0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45]
    for mat, blob in zip(wild, mst.pol_matrix_csv(wild)):
        assert blob == savetxt_bytes(mat)
    assert mst.pol_matrix_csv(mats[0])[0] == savetxt_bytes(mats[0])            # a single 2-D matrix
    # the node script writes its file through the same emitter
    coef = rng.normal(size=(2, 5, 4, 8)); dur = rng.uniform(0.1, 2, (2, 5))
    packed = mst.pack_pol_matrix(coef, dur)
    for b, blob in enumerate(mst.pol_matrix_csv(packed)):
        assert blob == savetxt_bytes(packed[b].cpu().numpy())


def test_pack_pol_matrix_matches_oracle_packing():
    from oracle import minsnap_oracle as mo
    import drone_path_planning_python_b200 as mst
    rng = np.random.default_rng(2)
    coef = rng.normal(size=(3, 5, 4, 8)) * 10.0 ** rng.integers(-3, 4, size=(3, 5, 4, 8))
    dur = rng.uniform(0.1, 2, (3, 5))
    got = mst.pack_pol_matrix(coef, dur).cpu().numpy()
    for b in range(3):
        assert np.array_equal(got[b], mo.pack_pol_matrix(coef[b], dur[b]))
