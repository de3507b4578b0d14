"""CPU tier: the drop-in packages expose the reference's module surface and import
without ROS / OMPL / FCL / matplotlib.  Numbers need the GPU (tests/test_gpu_dropin.py)."""
import importlib
import os
import sys

import numpy as np
import pytest


@pytest.fixture()
def dropin():
    import drone_path_planning_python_b200 as mst
    path = mst.dropin_path()
    sys.path.insert(0, path)
    for name in [m for m in sys.modules if m.split(".")[0] in ("optimizations", "RigidBodyPlanners", "scripts", "trajectory_visualising")]:
        del sys.modules[name]
    yield path
    sys.path.remove(path)
    for name in [m for m in sys.modules if m.split(".")[0] in ("optimizations", "RigidBodyPlanners", "scripts", "trajectory_visualising")]:
        del sys.modules[name]


def test_optimizations_star_import_surface(dropin):
    ns = {}
    exec("from optimizations import *", ns)   # what scripts/drones_pols_generator.py:16 does
    for name in ("normalize", "Polynomial", "TrajectoryOutput", "Polynomial4D", "Trajectory", "PiecewisePolynomial",
                 "Waypoint", "Point_time", "Point_time1D", "calculate_trajectory4D", "np"):
        assert name in ns, name
    ct = importlib.import_module("optimizations.calculatingTrajectories")
    for name in ("calculate_trajectory1D", "calculate_trajectory4D", "visualize_trajectory3D", "test_data", "timestep"):
        assert hasattr(ct, name), name
    assert len(ct.test_data) == 18 and ct.timestep == 2.0
    # the reference's own values (calculatingTrajectories.py:240-259), stored by oracle/make_golden.py
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "solve_cases.npz"))
    assert np.array_equal(np.asarray(ct.test_data, dtype=np.float64), gold["reference_test_data__wp"])


def test_trajectory_visualising_surface(dropin):
    """src/trajectory_visualising/__init__.py:1-2 and visualization.py's public names."""
    tv = importlib.import_module("trajectory_visualising")
    for name in ("Trajectory", "TrajectoryOutput", "get_nav_path_msg"):
        assert hasattr(tv, name), name
    vis = importlib.import_module("trajectory_visualising.visualization")
    for name in ("visualize_python", "get_nav_path_msg", "Trajectory", "TrajectoryOutput", "Polynomial4D", "np"):
        assert hasattr(vis, name), name
    ut = importlib.import_module("trajectory_visualising.uav_trajectory")
    for name in ("normalize", "Polynomial", "TrajectoryOutput", "Polynomial4D", "Trajectory"):
        assert hasattr(ut, name), name
    tr = tv.Trajectory()
    assert tr.polynomials is None and tr.duration is None


def test_plain_records_behave_like_the_reference(dropin, capsys):
    import optimizations as o
    wp = o.Waypoint(1.0, 2.0, 3.0, 0.5)
    assert (o.Waypoint.WP_TYPE_X, o.Waypoint.WP_TYPE_Y, o.Waypoint.WP_TYPE_Z, o.Waypoint.WP_TYPE_YAW) == (0, 1, 2, 3)
    assert [wp.getType(k) for k in range(4)] == [1.0, 2.0, 3.0, 0.5]
    assert wp.getType(7) is None and "Sorry, invalid type" in capsys.readouterr().out
    pt = o.Point_time(wp, 1.5)
    assert pt.wp is wp and pt.t == 1.5
    assert o.Point_time1D(2.0, 3.0).wp == 2.0
    pc = o.PiecewisePolynomial([o.Polynomial([1, 0]), o.Polynomial([0, 1])], [1, 2])
    assert pc.nOfPols == 2 and pc.time_durations == [1, 2] and len(pc.pols) == 2
    out = o.TrajectoryOutput()
    assert out.pos is None and out.omega is None
    p4 = o.Polynomial4D(1.0, [1] * 8, [2] * 8, [3] * 8, [0] * 8)
    assert p4.duration == 1.0 and list(p4.py.p) == [2] * 8
    assert np.allclose(o.normalize([3.0, 0.0, 4.0]), [0.6, 0.0, 0.8])
    with pytest.raises(AssertionError):
        o.normalize([0.0, 0.0, 0.0])
    with pytest.raises(AssertionError):
        o.Polynomial([1.0, 2.0]).eval(-1.0)         # the reference's `assert t >= 0`
    with pytest.raises(IndexError):
        o.calculate_trajectory4D([pt])               # single waypoint (quirk ii)


def test_trajectory_loadcsv_keeps_the_skiprows_quirk(dropin, tmp_path, golden_dir):
    import os
    import optimizations as o
    with np.load(os.path.join(golden_dir, "shipped_pol_matrices.npz")) as z:
        mat = z["Pol_matrix_1"]
    path = tmp_path / "Pol_matrix_1.csv"
    np.savetxt(str(path), mat, delimiter=",")
    tr = o.Trajectory()
    tr.loadcsv(str(path))
    assert tr.n_pieces() == 48                      # 49 rows written, first one eaten (quirk iii)
    assert abs(tr.duration - 9.6) < 1e-6
    with pytest.raises(AssertionError):
        tr.eval(tr.duration + 1.0)


def test_fcl_checker_module_surface(dropin):
    mod = importlib.import_module("RigidBodyPlanners.fcl_checker")
    for name in ("Fcl_mesh", "Fcl_checker", "visualize_meshes"):
        assert hasattr(mod, name)
    for meth in ("load_stl", "create_indexed_triangles", "create_fcl_mesh", "set_transform"):
        assert hasattr(mod.Fcl_mesh, meth)
    for meth in ("check_collision", "set_robot_transform", "check_collision_batch"):
        assert hasattr(mod.Fcl_checker, meth)


def test_node_scripts_import_without_ros(dropin):
    gen = importlib.import_module("scripts.drones_traj_generator")
    assert gen.drone_positions == [[0.5, 0, 0], [-0.5, 0, 0]]
    for name in ("drone_pose", "drone_pose2", "get_drone_positions", "transform", "callback", "listener",
                 "trajPub1", "trajPub2"):
        assert hasattr(gen, name)
    assert gen.drone_pose.pose.position.x == 0.5 and gen.drone_pose2.pose.position.x == -0.5
    poses = gen.get_drone_positions([[1, 2, 3], [4, 5, 6]])
    assert poses[0] is not poses[1] and poses[1].pose.position.z == 6
    pols = importlib.import_module("scripts.drones_pols_generator")
    for name in ("callback1", "callback2", "path_to_pol", "listener", "piece_pols_pub", "paths_to_matrices"):
        assert hasattr(pols, name)
    assert pols.callback1.counter == 0 and pols.callback2.counter == 0
