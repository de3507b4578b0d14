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


# Module-level fixture of the reference (src/optimizations/calculatingTrajectories.py:240-259): 18
# rigid-body waypoints (x, y, z, yaw) ``timestep`` seconds apart.  A drop-in exports the reference's
# VALUES (callers and the golden case "reference_test_data" in tests/golden/solve_cases.npz use them),
# so the table is kept verbatim.
timestep = 100/50
test_data = [
    [-1.0, 5.0, 1.0, 0.0],
    [-0.9105214656082087, 4.866527813557898, 0.9821609406403813, 0.02039080103534039],
    [-0.8225743363589189, 4.73288554489947, 0.964441662764501, 0.04077309721300639],
    [-0.7361272257466368, 4.5990008095166175, 0.946841139664208, 0.06113820463325185],
    [-0.6511925243577015, 4.464831726680945, 0.9293520786300249, 0.08147761001670759],
    [-0.5677385453806084, 4.3303072795034305, 0.9119680575644483, 0.10178299225536093],
    [-0.4857652247831987, 4.19536539455793, 0.8946893929748612, 0.12204500119949559],
    [-0.40523866600088, 4.05994725771634, 0.8775065705250998, 0.14225598708544368],
    [-0.3261547497876769, 3.923993481128284, 0.8604168009043632, 0.16240701194973708],
    [-0.2484752985307498, 3.78743548705188, 0.84341227981323, 0.1824899882345422],
    [-0.17219220053404993, 3.6502373470789204, 0.8264894614074699, 0.20249698043707148],
    [-0.09725801527295008, 3.51232995673728, 0.8096428181876703, 0.22241799607297275],
    [-0.02365621826047004, 3.37365871553904, 0.7928635266349502, 0.24224499181234838],
    [0.04864606691479989, 3.23417214717458, 0.7761488620638199, 0.2619709834334211],
    [0.11968750758167002, 3.0938169790745595, 0.75949244060815, 0.2815870183901405],
    [0.18949994071968002, 2.95253000758626, 0.7428843320283001, 0.3010839819518382],
    [0.2581249596866899, 2.8102848216266, 0.72632388402931, 0.32045501331803433],
    [0.32560330474396, 2.667021932848, 0.70980157478244, 0.3396910397023212]]

if __name__ == "__main__":
    traj_points = [Point_time(Waypoint(*row), t=i * timestep) for i, row in enumerate(test_data)]
    pieces, total = calculate_trajectory1D(traj_points, Waypoint.WP_TYPE_X)
    print("pieces:", len(pieces), "duration:", sum(total.time_durations))
