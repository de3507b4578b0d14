#!/usr/bin/env python
"""B200 drop-in for ``optimizations.uav_trajectory`` of mjmyt/drone_path_planning_python.

Same names, constructor arguments, attributes, return shapes and exceptions as the reference
module (src/optimizations/uav_trajectory.py); every number is produced by the CUDA kernels of
``drone_path_planning_python_b200`` (libmst.so) — there is no host arithmetic fallback, so a
machine without the library or without a CUDA device raises on the first evaluation.

The object-per-call surface is kept for compatibility with the ROS nodes; code that wants
throughput should use the batched tensor API (``drone_path_planning_python_b200.batch``) or the
``*_many`` helpers added here.
"""
import numpy as np

import drone_path_planning_python_b200 as _mst

_NCOEF = 8


def _host(t):
    return t.cpu().numpy()


def _padded(p):
    """Coefficients as a float64 row of 8 (zero padded; higher padding does not change a
    Horner evaluation).  The kernels hold 7th-order pieces, which is all the reference uses."""
    flat = np.asarray(p, dtype=np.float64).reshape(-1)
    if flat.size > _NCOEF:
        raise NotImplementedError("polynomials above 7th order are not supported by the CUDA path")
    out = np.zeros(_NCOEF)
    out[:flat.size] = flat
    return out


def normalize(v):
    """Unit vector of ``v`` (asserts a non-zero norm); reference: uav_trajectory.py:6-9.  A plain
    helper exported for compatibility — the flatness maths that uses it runs inside the CUDA
    kernel behind ``Polynomial4D.eval``."""
    v = np.asarray(v, dtype=np.float64)
    length = float(np.sqrt(np.dot(v, v)))
    assert length > 0
    return v / length


class Polynomial:
    """reference: uav_trajectory.py:12-36"""

    def __init__(self, p):
        self.p = p

    def _column_shaped(self):
        return isinstance(self.p, np.ndarray) and self.p.ndim == 2

    # evaluate a polynomial using horner's rule (on the GPU: mst_sample_batch, one piece)
    def eval(self, t):
        assert t >= 0
        if len(self.p) == 0:
            return 0.0
        flat = np.asarray(self.p, dtype=np.float64).reshape(-1)
        if flat.size > _NCOEF:
            raise NotImplementedError("polynomials above 7th order are not supported by the CUDA path")
        # one launch + one stream synchronisation, arguments through pinned memory (sample_piece_now)
        val = _mst.sample_piece_now(flat, float(t))
        # with (8,1)-shaped coefficients the reference's Horner loop yields a shape-(1,) array
        return np.array([val]) if self._column_shaped() else val

    def eval_many(self, ts, deriv=0):
        """Values (or ``deriv``-th derivative) at many times in one launch."""
        ts = np.asarray(ts, dtype=np.float64).reshape(-1)
        assert (ts >= 0).all()
        coef = _padded(self.p).reshape(1, 1, 1, _NCOEF)
        return _host(_mst.sample_batch(coef, np.ones((1, 1)), ts=ts, deriv=deriv))[0, :, 0]

    def derivative(self):
        """d/dt as a new Polynomial (mst_poly_derivative)"""
        n = len(self.p)
        if n <= 1:
            return Polynomial([])
        flat = np.asarray(self.p, dtype=np.float64).reshape(1, -1)
        d = _host(_mst.poly_derivative(flat))[0]
        if self._column_shaped():
            return Polynomial([np.array([v]) for v in d])  # list of shape-(1,) arrays, as the reference builds
        return Polynomial([float(v) for v in d])

    def pol_coeffs_at_t(self, t):
        """The terms p[i] * t**i as an array (their sum is the value at t); mst_poly_terms_at_t."""
        assert t >= 0
        n = len(self.p)
        flat = np.asarray(self.p, dtype=np.float64).reshape(1, -1)
        coeffs = np.zeros(n)
        if n:
            coeffs[:] = _host(_mst.poly_terms_at_t(flat, np.array([float(t)])))[0]
        return coeffs


class TrajectoryOutput:
    """Flat outputs of one evaluation, all ``None`` until filled: ``pos`` / ``vel`` / ``acc`` (3-vectors in
    m, m/s, m/s^2), ``omega`` (body rates, rad/s) and ``yaw`` (rad).  reference: uav_trajectory.py:39-45"""

    def __init__(self):
        for field in ("pos", "vel", "acc", "omega", "yaw"):
            setattr(self, field, None)


def _flat_to_output(row):
    out = TrajectoryOutput()
    out.pos = row[0:3].copy()
    out.vel = row[3:6].copy()
    out.acc = row[6:9].copy()
    out.omega = row[9:12].copy()
    out.yaw = float(row[12])
    return out


class Polynomial4D:
    """One piece of an (x, y, z, yaw) trajectory with its duration.  reference: uav_trajectory.py:49-101"""

    def __init__(self, duration, px, py, pz, pyaw):
        self.duration = duration
        self.px, self.py, self.pz, self.pyaw = (Polynomial(axis) for axis in (px, py, pz, pyaw))

    def _coef(self):
        return np.stack([_padded(self.px.p), _padded(self.py.p), _padded(self.pz.p), _padded(self.pyaw.p)])

    def derivative(self):
        axes = (self.px, self.py, self.pz, self.pyaw)
        return Polynomial4D(self.duration, *(axis.derivative().p for axis in axes))

    def eval(self, t):
        assert t >= 0  # Polynomial.eval's assert in the reference
        coef = self._coef().reshape(1, 1, 4, _NCOEF)
        # a single piece evaluated at local time t, also past its nominal duration
        row = _host(_mst.flat_outputs(coef, np.ones((1, 1)), ts=np.array([float(t)]), mode="piecewise"))[0, 0]
        return _flat_to_output(row)


class Trajectory:
    """reference: uav_trajectory.py:104-127"""

    def __init__(self):
        self.polynomials = None
        self.duration = None
        self._dev = None

    def n_pieces(self):
        return len(self.polynomials)

    def loadcsv(self, filename):
        # skiprows=1 although path_to_pol writes no header: kept
        table = np.atleast_2d(np.loadtxt(filename, delimiter=",", skiprows=1, usecols=range(33)))
        # a row is [duration | 8 x | 8 y | 8 z | 8 yaw]
        self.polynomials = [Polynomial4D(r[0], *(r[1 + 8 * a:9 + 8 * a] for a in range(4))) for r in table]
        self.duration = np.sum(table[:, 0])
        self._dev = None

    def _device_arrays(self):
        if self._dev is None:
            coef = np.stack([p._coef() for p in self.polynomials])[None]          # [1, n, 4, 8]
            dur = np.array([float(p.duration) for p in self.polynomials])[None]   # [1, n]
            self._dev = (_mst.batch._f64(coef, _mst._abi.require_cuda()),
                         _mst.batch._f64(dur, _mst._abi.require_cuda()))
        return self._dev

    def eval(self, t):
        assert t >= 0
        assert t <= self.duration
        coef, dur = self._device_arrays()
        row, status = _mst.sample_now(coef, dur, float(t), mode="trajectory", flat=True)
        if status != 0:
            return None  # the reference falls off its loop (rounding of the running sum)
        return _flat_to_output(row)

    def eval_many(self, ts):
        """``[S, 13]`` rows ``pos vel acc omega yaw`` for many times in one launch."""
        ts = np.asarray(ts, dtype=np.float64).reshape(-1)
        assert (ts >= 0).all() and (ts <= self.duration).all()
        coef, dur = self._device_arrays()
        return _host(_mst.flat_outputs(coef, dur, ts=ts, mode="trajectory"))[0]


class PiecewisePolynomial():
    """
    Piece-wise polynomial (reference: uav_trajectory.py:130-169): ``pols`` is the list of
    ``Polynomial`` pieces in order, ``time_durations`` the list of their durations in seconds.
    """

    def __init__(self, pols: list, time_durations: list):
        self.pols = pols
        self.nOfPols = len(pols)
        self.time_durations = time_durations

    def _arrays(self):
        coef = np.stack([_padded(p.p) for p in self.pols]).reshape(1, self.nOfPols, 1, _NCOEF)
        dur = np.asarray([float(d) for d in self.time_durations], dtype=np.float64).reshape(1, -1)
        return coef, dur

    def _device_arrays(self):
        """Device copies of the pieces, rebuilt when the piece list or a duration changes."""
        key = (tuple(id(p.p) for p in self.pols), tuple(self.time_durations))
        cached = getattr(self, "_dev", None)
        if cached is None or cached[0] != key:
            coef, dur = self._arrays()
            dev = _mst._abi.require_cuda()
            cached = self._dev = (key, _mst.batch._f64(coef, dev), _mst.batch._f64(dur, dev))
        return cached[1], cached[2]

    def eval(self, t):
        assert t >= 0
        coef, dur = self._device_arrays()
        val = float(_mst.sample_now(coef, dur, float(t), mode="piecewise")[0][0])
        column = self.nOfPols > 0 and isinstance(self.pols[0].p, np.ndarray) and self.pols[0].p.ndim == 2
        return np.array([val]) if column else val

    def eval_many(self, ts, deriv=0):
        ts = np.asarray(ts, dtype=np.float64).reshape(-1)
        assert (ts >= 0).all()
        coef, dur = self._arrays()
        return _host(_mst.sample_batch(coef, dur, ts=ts, deriv=deriv))[0, :, 0]


class Waypoint():
    """One waypoint (x, y, z, yaw); reference: uav_trajectory.py:172-195."""
    WP_TYPE_X = 0
    WP_TYPE_Y = 1
    WP_TYPE_Z = 2
    WP_TYPE_YAW = 3
    _FIELDS = ("x", "y", "z", "yaw")

    def __init__(self, x, y, z, yaw):
        self.x, self.y, self.z, self.yaw = x, y, z, yaw

    def getType(self, type):
        if type in (0, 1, 2, 3):
            return getattr(self, Waypoint._FIELDS[type])
        print("Sorry, invalid type")  # the reference prints and returns None


class Point_time():
    """Waypoint + time stamp for trajectory generation; reference: uav_trajectory.py:198-202."""

    def __init__(self, wp: Waypoint, t: float):
        self.wp, self.t = wp, t


class Point_time1D():
    """Scalar waypoint + time stamp; reference: uav_trajectory.py:205-209."""

    def __init__(self, wp: float, t: float):
        self.wp, self.t = wp, t
