"""CPU oracle for the trajectory half of the hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it, and only as the checker / the timed CPU arm.  The
product path (``drone_path_planning_python_b200``) never imports this module and
fails loudly when its CUDA library is missing.

This is a plain numpy / pure-Python *restatement* of the reference's algorithm,
written from the reference's behaviour (file:line cited per function, paths
relative to the reference checkout).  Parity is PINNED: ``oracle/make_golden.py``
imports the unmodified reference in the build container, checks this restatement
against it bit for bit on seeded inputs, and commits the reference's own outputs
under ``tests/golden/`` (see that script's header).

Conventions (same as the reference): coefficients are in ASCENDING power order,
8 per piece; ``n`` pieces join ``m = n + 1`` waypoints; durations are
``T_i = t_{i+1} - t_i``.
"""
from __future__ import annotations

import math

import numpy as np

N_COEF = 8  # 7th-order pieces (src/optimizations/calculatingTrajectories.py:14)


# --------------------------------------------------------------------------- a1
def time_power_row(t: float, deriv: int) -> np.ndarray:
    """Row of the ``deriv``-th derivative of ``sum_k c_k t^k`` w.r.t. ``c``.

    Restates ``Polynomial([1]*8)`` differentiated ``deriv`` times, then
    ``pol_coeffs_at_t(t)``, then left-padded with ``deriv`` zeros
    (src/optimizations/uav_trajectory.py:25-36 and the ``np.pad`` calls at
    src/optimizations/calculatingTrajectories.py:68-71,93-97,105-109).

    Entry ``k`` is ``k!/(k-deriv)! * t**(k-deriv)`` for ``k >= deriv`` else 0.
    The reference multiplies an exact Python int by ``t**i`` evaluated with the
    Python float ``**`` operator (``0.0**0 == 1.0``); the same two operations are
    used here so the row is bit-identical.
    """
    t = float(t)
    if not t >= 0:  # uav_trajectory.py:30
        raise AssertionError("t must be >= 0")
    row = np.zeros(N_COEF)
    for k in range(deriv, N_COEF):
        falling = math.factorial(k) // math.factorial(k - deriv)  # exact int
        row[k] = falling * (t ** (k - deriv))
    return row


# --------------------------------------------------------------------------- a2
def assemble_system(values, times):
    """Square system ``A c = b`` for one axis.

    Restates the row layout of ``calculate_trajectory1D``
    (src/optimizations/calculatingTrajectories.py:48-131):

    * rows 0..3: derivatives 0..3 of piece 0 at local time ``times[0]``
      (the reference evaluates the start rows at ``t_0`` itself, quirk (i) of
      SURVEY §8a), rhs ``[v_0, 0, 0, 0]``                       (:65-73,82-85)
    * interior waypoint i (1..n-1), ``s = 4 + 8(i-1)``:
      rows s..s+5   derivative j=1..6 of piece i-1 at ``T_{i-1}`` minus the same
      derivative of piece i at 0                                (:115-121)
      row  s+6      piece i-1 at ``T_{i-1}`` equals ``v_i``      (:124,127)
      row  s+7      piece i at 0 equals ``v_i``                  (:125,128)
    * last 4 rows: derivatives 0..3 of piece n-1 at ``T_{n-1}``,
      rhs ``[v_n, 0, 0, 0]``                                    (:74-79,86-87)

    Returns ``(A, b, durations)`` with ``durations`` a list of ``n`` floats
    (the reference's ``time_points``, :59-61).
    """
    m = len(values)
    n = m - 1
    if n < 1:
        # the reference indexes ``A[j, 0:8]`` of a 0x0 matrix -> IndexError
        raise IndexError("need at least two waypoints")
    size = N_COEF * n
    A = np.zeros((size, size))
    b = np.zeros((size, 1))
    durations = []
    prev = 0.0
    for i in range(m):
        local = float(times[i]) - prev  # :58
        if i != 0:
            durations.append(local)
        if i == 0:
            for j in range(4):
                A[j, 0:N_COEF] = time_power_row(local, j)
            b[0, 0] = values[0]
        elif i == n:
            for j in range(4):
                A[size - 4 + j, N_COEF * (n - 1):N_COEF * n] = time_power_row(local, j)
            b[size - 4, 0] = values[n]
        else:
            s = 4 + N_COEF * (i - 1)
            left = slice(N_COEF * (i - 1), N_COEF * i)
            right = slice(N_COEF * i, N_COEF * (i + 1))
            for j in range(1, 7):
                A[s + j - 1, left] = time_power_row(local, j)
                A[s + j - 1, right] = -time_power_row(0.0, j)
            A[s + 6, left] = time_power_row(local, 0)
            A[s + 7, right] = time_power_row(0.0, 0)
            b[s + 6, 0] = values[i]
            b[s + 7, 0] = values[i]
        prev = float(times[i])
    return A, b, durations


# --------------------------------------------------------------------------- a2/a3
def solve_axis(values, times):
    """One axis: dense LU with partial pivoting, as ``np.linalg.solve`` at
    src/optimizations/calculatingTrajectories.py:137.  Returns
    ``(coef[n, 8], durations)``; raises ``numpy.linalg.LinAlgError`` on a
    singular system exactly as the reference does."""
    A, b, durations = assemble_system(values, times)
    x = np.linalg.solve(A, b)
    return x.reshape(len(durations), N_COEF), durations


# --------------------------------------------------------------------------- a4
def solve_waypoints(waypoints, times):
    """All axes of one trajectory: ``waypoints[m, K]``, ``times[m]`` ->
    ``(coef[n, K, 8], durations[n])``.  The reference solves every axis
    independently, re-assembling the same matrix each time
    (``calculate_trajectory4D``, calculatingTrajectories.py:200-213); so does
    this restatement, on purpose: it is also the timed CPU arm."""
    waypoints = np.asarray(waypoints, dtype=np.float64)
    m, K = waypoints.shape
    out = np.zeros((m - 1, K, N_COEF))
    durations = None
    for k in range(K):
        c, durations = solve_axis(waypoints[:, k], times)
        out[:, k, :] = c
    return out, np.asarray(durations, dtype=np.float64)


# --------------------------------------------------------------------------- a5
def horner(coefs, t):
    """``Polynomial.eval`` (src/optimizations/uav_trajectory.py:17-22): Horner
    from the highest power, multiply and add rounded separately (no FMA)."""
    t = float(t)
    if not t >= 0:
        raise AssertionError("t must be >= 0")
    x = 0.0
    for c in reversed([float(v) for v in coefs]):
        x = x * t + c
    return x


def derivative_coefs(coefs):
    """``Polynomial.derivative`` (uav_trajectory.py:25-26): ``(i+1) * p[i+1]``."""
    return [(i + 1) * float(coefs[i + 1]) for i in range(len(coefs) - 1)]


# --------------------------------------------------------------------------- a6
def piece_lookup(durations, t):
    """Piece index and local time under ``PiecewisePolynomial.eval`` semantics
    (uav_trajectory.py:154-169): strict ``t < acc + T_i`` with the running sum
    accumulated left to right; past the end the LAST piece is evaluated at
    ``t - sum(T[:-1])`` (extrapolation)."""
    t = float(t)
    if not t >= 0:
        raise AssertionError("t must be >= 0")
    acc = 0
    for i, T in enumerate(durations):
        if t < acc + T:
            return i, t - acc
        acc = acc + T
    return len(durations) - 1, t - sum(durations[:-1])


def piecewise_eval(coef, durations, t, deriv=0):
    """Value (or ``deriv``-th derivative) of one axis ``coef[n, 8]`` at ``t``."""
    i, local = piece_lookup([float(d) for d in durations], t)
    c = [float(v) for v in coef[i]]
    for _ in range(deriv):
        c = derivative_coefs(c)
    return horner(c, local)


def trajectory_lookup(durations, t):
    """Piece index / local time under ``Trajectory.eval`` semantics
    (uav_trajectory.py:119-127): asserts ``0 <= t <= sum(T)``, INCLUSIVE
    ``t <= acc + T_i``."""
    t = float(t)
    total = float(np.sum(np.asarray(durations, dtype=np.float64)))
    if not (t >= 0 and t <= total):
        raise AssertionError("t outside [0, duration]")
    acc = 0.0
    for i, T in enumerate(durations):
        if t <= acc + T:
            return i, t - acc
        acc = acc + T
    return None  # the reference falls off the loop and returns None


def uniform_sample_times(durations, S):
    """Sample times of the batched pipeline when the caller passes none:
    ``t_s = s * (total / S)``, ``s = 0..S-1``, ``total`` the left-to-right sum of
    the durations (the ``np.arange(0, duration, timestep)`` pattern of
    src/trajectory_visualising/visualization.py:53 with ``timestep = total/S``)."""
    total = 0.0
    for T in durations:
        total = total + float(T)
    dt = total / S
    return np.array([s * dt for s in range(S)], dtype=np.float64)


def sample_trajectory(coef, durations, ts, deriv=0):
    """``coef[n, K, 8]`` sampled at ``ts[S]`` -> ``out[S, K]``
    (PiecewisePolynomial semantics per axis)."""
    n, K, _ = coef.shape
    out = np.zeros((len(ts), K))
    for s, t in enumerate(ts):
        for k in range(K):
            out[s, k] = piecewise_eval(coef[:, k, :], durations, t, deriv)
    return out


# --------------------------------------------------------------------------- Polynomial4D
def flat_output(piece, t):
    """``Polynomial4D.eval`` (uav_trajectory.py:66-101): position, velocity,
    acceleration, yaw and the differential-flatness body rates ``omega`` of one
    piece ``piece[4, 8]`` (x, y, z, yaw) at local time ``t``."""
    c0 = [[float(v) for v in piece[k]] for k in range(4)]
    c1 = [derivative_coefs(c) for c in c0]
    c2 = [derivative_coefs(c) for c in c1]
    c3 = [derivative_coefs(c) for c in c2]
    pos = np.array([horner(c0[k], t) for k in range(3)])
    yaw = horner(c0[3], t)
    vel = np.array([horner(c1[k], t) for k in range(3)])
    dyaw = horner(c1[3], t)
    acc = np.array([horner(c2[k], t) for k in range(3)])
    jerk = np.array([horner(c3[k], t) for k in range(3)])
    thrust = acc + np.array([0, 0, 9.81])
    tn = np.linalg.norm(thrust)
    z_body = thrust / tn
    x_world = np.array([np.cos(yaw), np.sin(yaw), 0])
    yb = np.cross(z_body, x_world)
    y_body = yb / np.linalg.norm(yb)
    x_body = np.cross(y_body, z_body)
    h_w = (jerk - (np.dot(jerk, z_body) * z_body)) / tn
    omega = np.array([-np.dot(h_w, y_body), np.dot(h_w, x_body), z_body[2] * dyaw])
    return {"pos": pos, "vel": vel, "acc": acc, "yaw": yaw, "omega": omega}


# --------------------------------------------------------------------------- a8
def uniform_times(m, total_duration=10.0):
    """``path_to_pol`` time stamps (scripts/drones_pols_generator.py:44-56):
    ``t_i = (total/m) * i`` — note the step divides by the number of poses."""
    step = total_duration / m
    return np.array([step * i for i in range(m)], dtype=np.float64)


def pack_pol_matrix(coef, durations):
    """``(n, 1 + 8K)`` float32 matrix ``[T | x0..x7 | y0..y7 | ...]`` of
    ``path_to_pol`` (scripts/drones_pols_generator.py:63-77)."""
    n, K, _ = coef.shape
    mat = np.zeros((n, 1 + N_COEF * K), dtype=np.float32)
    for k in range(K):
        mat[:, 1 + N_COEF * k:1 + N_COEF * (k + 1)] = coef[:, k, :]
    mat[:, 0] = np.asarray(durations)
    return mat


# --------------------------------------------------------------------------- a9
def quat_rotate(q_xyzw, v):
    """Rotate ``v`` by the unit quaternion ``(x, y, z, w)`` — what
    ``tf2_geometry_msgs.do_transform_pose`` (KDL ``Rotation::Quaternion``) does
    to the drone offset at scripts/drones_traj_generator.py:77-82."""
    x, y, z, w = [float(c) for c in q_xyzw]
    R = np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ])
    return R @ np.asarray(v, dtype=np.float64)


def yaw_quat(yaw):
    """``quaternion_from_euler(0, 0, yaw)`` in xyzw order
    (src/RigidBodyPlanners/RB_planning_sep_coll_check.py:212)."""
    return np.array([0.0, 0.0, math.sin(yaw / 2.0), math.cos(yaw / 2.0)])


def formation_waypoints(rb_poses, offsets):
    """``transform(path)`` of scripts/drones_traj_generator.py:56-89 for poses
    ``rb_poses[m, 4] = (x, y, z, yaw)``: drone ``d`` waypoint
    ``R(q_rb) @ offset_d + t_rb`` with the rigid body's yaw carried over
    (the drone offset poses have identity orientation, :31-38).
    Returns ``[D, m, 4]``."""
    rb_poses = np.asarray(rb_poses, dtype=np.float64)
    offsets = np.asarray(offsets, dtype=np.float64)
    D, m = offsets.shape[0], rb_poses.shape[0]
    out = np.zeros((D, m, 4))
    for i in range(m):
        q = yaw_quat(rb_poses[i, 3])
        for d in range(D):
            out[d, i, :3] = quat_rotate(q, offsets[d]) + rb_poses[i, :3]
            out[d, i, 3] = rb_poses[i, 3]
    return out


# --------------------------------------------------------------------------- extension
def snap_cost(coef, durations):
    """Integral of the squared 4th derivative over all pieces and axes of ``coef[n, K, 8]`` -
    the objective whose optimality system the reference's square system is (the reference never
    evaluates it).  Exact polynomial integration with numpy.polynomial."""
    from numpy.polynomial import polynomial as P
    total = 0.0
    for i, T in enumerate(durations):
        for k in range(coef.shape[1]):
            d4 = P.polyder(np.asarray(coef[i, k], dtype=np.float64), 4)
            total += P.polyval(float(T), P.polyint(P.polymul(d4, d4)))
    return float(total)


def time_gradient(coef):
    """d(optimal snap cost)/d(duration of piece i), waypoints fixed, knot derivatives free: minus the
    Hamiltonian of the piece summed over axes, H = x4^2 - 2 x5 x3 + 2 x6 x2 - 2 x7 x1 with xk the k-th
    derivative at the start of the piece (constant along an optimal piece).  ``coef[n, K, 8]`` ->
    ``grad[n]``.  Checker of mst_time_gradient; itself checked against central differences of
    re-solves in tests/test_cpu_oracle.py."""
    coef = np.asarray(coef, dtype=np.float64)
    x = coef * np.array([1.0, 1.0, 2.0, 6.0, 24.0, 120.0, 720.0, 5040.0])
    H = x[..., 4] ** 2 - 2.0 * x[..., 5] * x[..., 3] + 2.0 * x[..., 6] * x[..., 2] - 2.0 * x[..., 7] * x[..., 1]
    return -H.sum(axis=-1)


def optimize_time_allocation(waypoints, times, iters=8, line_search=6, min_fraction=0.1):
    """Checker for ``drone_path_planning_python_b200.time_allocation`` (an extension: the reference
    never searches over stamps, so there is nothing to pin this against — parity UNPINNED, the
    two implementations are only checked against each other and against the properties of the
    method).  One problem: ``waypoints[m, K]``, ``times[m]`` -> ``(times_new[m], cost[iters + 1])``.
    Projected gradient descent on the durations, first stamp and total fixed: the gradient of one
    solve (``time_gradient``) projected on ``sum T = const``, candidates ``T - cap / 2^k * g``,
    best one kept when it lowers the snap cost."""
    waypoints = np.asarray(waypoints, dtype=np.float64)
    times = np.asarray(times, dtype=np.float64)
    n = len(times) - 1

    def solve(T):
        t = np.concatenate([[times[0]], times[0] + np.cumsum(T)])
        try:
            coef, dur = solve_waypoints(waypoints, t)
        except (np.linalg.LinAlgError, AssertionError):
            return np.inf, None
        c = snap_cost(coef, dur)
        return (c, coef) if np.isfinite(c) else (np.inf, None)

    T = np.diff(times)
    J, coef = solve(T)
    history = [J]
    if n < 2:
        return times.copy(), np.array(history * (iters + 1))
    total = T.sum()
    floor = min_fraction * total / n
    for _ in range(iters):
        J0 = history[-1]
        g = time_gradient(coef) if coef is not None else np.zeros(n)
        g = np.where(np.isfinite(g), g, 0.0)
        g = g - g.mean()
        cap = 0.5 * max((T - floor).min(), 0.0) / max(np.abs(g).max(), 1e-300)
        best_J, best_T, best_coef = np.inf, T, coef
        for k in range(line_search):
            cand = T - cap * 0.5 ** k * g
            Jk, ck = solve(cand)
            if Jk < best_J:
                best_J, best_T, best_coef = Jk, cand, ck
        if best_J < J0:
            T, coef = best_T, best_coef
            history.append(best_J)
        else:
            history.append(J0)
    out = np.concatenate([[times[0]], times[0] + np.cumsum(T)])
    out[-1] = times[-1]
    return out, np.array(history)

