"""B200 drop-in for ``optimizations.calculatingTrajectories``
(reference: src/optimizations/calculatingTrajectories.py).

``calculate_trajectory1D`` / ``calculate_trajectory4D`` keep the reference's signature, return
structure ((8,1) float64 coefficient views, ``PiecewisePolynomial`` with a Python list of
durations) and exceptions, but the 8n x 8n system is never built on the host: the waypoints and
time stamps go to the GPU, ``mst_solve_batch`` assembles and solves it there (all axes in one
launch, one factorisation instead of the reference's four), and the coefficients come back.
"""
import numpy as np

import drone_path_planning_python_b200 as _mst

try:
    from uav_trajectory import *  # noqa: F401,F403  (same import dance as the reference, :8-11)
except ImportError:
    from .uav_trajectory import *  # noqa: F401,F403


def _solve(waypoints, axes):
    """Shared back end: list of Point_time -> (coef[n, len(axes), 8], durations[n]) on the host."""
    m = len(waypoints)
    if m < 2:
        # the reference indexes a 0 x 0 matrix for a single waypoint
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")
    values = np.zeros((1, m, len(axes)))
    stamps = np.zeros((1, m))
    for i, point in enumerate(waypoints):
        stamps[0, i] = point.t
        for k, axis in enumerate(axes):
            values[0, i, k] = point.wp.getType(axis)
    coef, dur, info = _mst.solve_batch(values, stamps)
    code = int(info[0])
    if code == _mst._abi.INFO_DECREASING or code == _mst._abi.INFO_NONFINITE:
        raise AssertionError("waypoint times must be non-negative and non-decreasing")  # uav_trajectory.py:30
    if code != 0:
        raise np.linalg.LinAlgError("Singular matrix")  # what np.linalg.solve raises at :137
    return coef[0].cpu().numpy(), dur[0].cpu().numpy()


def _pack(coef_axis, durations):
    """(piece_pols, total_pol) exactly as the reference returns them (:141-144,191,197)."""
    n = coef_axis.shape[0]
    column = np.ascontiguousarray(coef_axis.reshape(8 * n, 1))
    piece_pols = [Polynomial(column[8 * i:8 * (i + 1)]) for i in range(n)]
    time_points = [float(d) for d in durations]
    return piece_pols, PiecewisePolynomial(piece_pols, time_points)


def calculate_trajectory1D(waypoints, wp_type=Waypoint.WP_TYPE_X):
    """
    waypoints: list of Point_Time

    wp_type: specifies the type of waypoint (x,y,z or yaw)
    """
    coef, dur = _solve(waypoints, [wp_type])
    return _pack(coef[:, 0, :], dur)


def calculate_trajectory4D(waypoints):
    # waypoints: list of Point_time instances; one GPU solve for x, y, z and yaw together
    coef, dur = _solve(waypoints, [Waypoint.WP_TYPE_X, Waypoint.WP_TYPE_Y, Waypoint.WP_TYPE_Z,
                                   Waypoint.WP_TYPE_YAW])
    pols_coeffs, pc_pols = [], []
    for k in range(4):
        pieces, total = _pack(coef[:, k, :], dur)
        pols_coeffs.append(pieces)
        pc_pols.append(total)
    return pols_coeffs, pc_pols


def visualize_trajectory3D(pols):
    """Scatter plot of a 3-D trajectory sampled at 100 times in [0, 10] s (reference :216-237);
    matplotlib is imported lazily because only this function needs it."""
    import matplotlib.pyplot as plt
    from mpl_toolkits.mplot3d import Axes3D  # noqa: F401

    samples = np.linspace(0, 10, 100)
    xyz = [pols[k].eval_many(samples) for k in range(3)]
    fig = plt.figure()
    ax = fig.add_subplot(111, projection='3d')
    ax.set_xlabel('X')
    ax.set_ylabel('Y')
    ax.set_zlabel('Z')
    ax.scatter(xyz[0], xyz[1], xyz[2], c='r', marker='o')
    plt.show()


# Module-level smoke fixture with the names, shape and role of the reference's (:240-259): 18
# rigid-body waypoints (x, y, z, yaw) two seconds apart.  The values here are this package's own — a
# gentle descending arc through the reference's workspace bounds; the reference's values live in
# tests/golden/solve_cases.npz (case "reference_test_data"), where the parity tests use them.
timestep = 100/50
test_data = [[round(-1.0 + 0.08 * i, 6), round(5.0 - 0.14 * i - 0.0004 * i * i, 6), round(1.0 - 0.017 * i, 6),
              round(0.02 * i, 6)] for i in range(18)]

if __name__ == "__main__":
    traj_points = [Point_time(Waypoint(*row), t=i * timestep) for i, row in enumerate(test_data)]
    pieces, total = calculate_trajectory1D(traj_points, Waypoint.WP_TYPE_X)
    print("pieces:", len(pieces), "duration:", sum(total.time_durations))
