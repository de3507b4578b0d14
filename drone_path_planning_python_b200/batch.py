"""Tensor-level API over libmst.so: the throughput path.

Every function takes torch tensors (CUDA, or host tensors / numpy arrays which are copied
to the current CUDA device), launches on ``torch.cuda.current_stream()`` and returns CUDA
tensors.  PyTorch is only the allocator / stream provider here; all arithmetic happens in
the hand-written kernels behind the C ABI (include/mst.h).  There is no CPU fallback.

The legacy, object-per-trajectory surface of the reference (``optimizations``,
``RigidBodyPlanners.fcl_checker``) lives in ``dropin/`` and is a thin adapter over this
module.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import numpy as np
import torch

from . import _abi

_SOLVERS = {"auto": _abi.SOLVER_AUTO, "banded_lu": _abi.SOLVER_BANDED_LU, "condensed": _abi.SOLVER_CONDENSED,
            "auto_one_pass": _abi.SOLVER_AUTO_ONE_PASS}   # the last one: pipeline() only
_MODES = {"piecewise": _abi.SAMPLE_PIECEWISE, "trajectory": _abi.SAMPLE_TRAJECTORY}


def _f64(x, device) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.float64).contiguous()
    return torch.as_tensor(np.ascontiguousarray(np.asarray(x, dtype=np.float64)), device=device)


def _check_out(name: str, x, shape, dtype, device) -> None:
    """Caller-supplied output buffers go to the kernels as raw pointers: refuse anything the
    kernels would write out of bounds or mis-stride."""
    if not isinstance(x, torch.Tensor) or not x.is_cuda or x.device != device:
        raise ValueError("%s must be a CUDA tensor on %s" % (name, device))
    if x.dtype != dtype or tuple(x.shape) != tuple(shape) or not x.is_contiguous():
        raise ValueError("%s must be a contiguous %s tensor of shape %s (got %s %s)"
                         % (name, dtype, tuple(shape), x.dtype, tuple(x.shape)))


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# --------------------------------------------------------------------------- a1
def time_power_rows(t) -> torch.Tensor:
    """``rows[count, 8, 8]``: derivative j, power k -> ``k!/(k-j)! * t**(k-j)``
    (reference: Polynomial.pol_coeffs_at_t, src/optimizations/uav_trajectory.py:28-36)."""
    dev = _abi.require_cuda()
    lib = _abi.load()
    tt = _f64(t, dev).reshape(-1)
    rows = torch.empty((tt.numel(), 8, 8), dtype=torch.float64, device=dev)
    _abi.check(lib.mst_time_power_rows(_ptr(tt), tt.numel(), _ptr(rows), _stream_ptr()), "mst_time_power_rows")
    return rows


def poly_derivative(p) -> torch.Tensor:
    """``p[count, len]`` -> ``[count, len-1]``: ``(i+1) * p[i+1]``
    (reference: Polynomial.derivative, src/optimizations/uav_trajectory.py:25-26)."""
    dev = _abi.require_cuda()
    lib = _abi.load()
    pp = _f64(p, dev)
    if pp.dim() == 1:
        pp = pp[None]
    count, ln = pp.shape
    out = torch.empty((count, max(ln - 1, 0)), dtype=torch.float64, device=dev)
    _abi.check(lib.mst_poly_derivative(_ptr(pp), count, ln, _ptr(out), _stream_ptr()), "mst_poly_derivative")
    return out


def poly_terms_at_t(p, t) -> torch.Tensor:
    """``p[count, len]``, ``t[count]`` -> ``p[c, i] * t[c]**i``
    (reference: Polynomial.pol_coeffs_at_t, src/optimizations/uav_trajectory.py:28-36)."""
    dev = _abi.require_cuda()
    lib = _abi.load()
    pp = _f64(p, dev)
    if pp.dim() == 1:
        pp = pp[None]
    tt = _f64(t, dev).reshape(-1)
    count, ln = pp.shape
    out = torch.empty((count, ln), dtype=torch.float64, device=dev)
    _abi.check(lib.mst_poly_terms_at_t(_ptr(pp), _ptr(tt), count, ln, _ptr(out), _stream_ptr()), "mst_poly_terms_at_t")
    return out


# --------------------------------------------------------------------------- a2-a4
def solve_batch(wp, t, share_time_group: int = 1, solver: str = "auto"
                ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Minimum-snap solve of ``B`` trajectories.

    wp ``[B, n+1, K]`` waypoints, t ``[B // share_time_group, n+1]`` time stamps.
    Returns ``coef[B, n, K, 8]`` (ascending powers), ``dur[B, n]``, ``info[B]`` (int32;
    0 ok, >0 singular at that column, <0 bad input — see include/mst.h).
    Reference: calculate_trajectory1D/4D, src/optimizations/calculatingTrajectories.py:37-213.
    """
    dev = _abi.require_cuda()
    lib = _abi.load()
    wp = _f64(wp, dev)
    t = _f64(t, dev)
    if wp.dim() != 3 or t.dim() != 2:
        raise ValueError("wp must be [B, n+1, K] and t [B/G, n+1]")
    B, m, K = wp.shape
    n = m - 1
    G = int(share_time_group)
    if n < 1:
        raise IndexError("need at least two waypoints")  # the reference raises IndexError
    if G < 1 or B % G != 0 or t.shape[0] != B // G or t.shape[1] != m:
        raise ValueError("t must be [B/share_time_group, n+1]")
    coef = torch.empty((B, n, K, 8), dtype=torch.float64, device=dev)
    dur = torch.empty((B, n), dtype=torch.float64, device=dev)
    info = torch.empty((B,), dtype=torch.int32, device=dev)
    ws = torch.empty((max(1, lib.mst_solve_workspace_bytes(B, n, K, G)),), dtype=torch.uint8, device=dev)
    rc = lib.mst_solve_batch(_ptr(wp), _ptr(t), B, n, K, G, _SOLVERS[solver], _ptr(coef), _ptr(dur),
                             _ptr(info), _ptr(ws), _stream_ptr())
    _abi.check(rc, "mst_solve_batch")
    return coef, dur, info


def snap_cost(coef, dur) -> torch.Tensor:
    """``cost[B]`` = sum over pieces and axes of the integral of the squared 4th derivative
    (the objective whose optimality system the reference solves; an extension for
    time-allocation searches)."""
    dev = _abi.require_cuda()
    lib = _abi.load()
    coef = _f64(coef, dev)
    dur = _f64(dur, dev)
    B, n, K, _ = coef.shape
    cost = torch.empty((B,), dtype=torch.float64, device=dev)
    _abi.check(lib.mst_snap_cost(_ptr(coef), _ptr(dur), B, n, K, _ptr(cost), _stream_ptr()), "mst_snap_cost")
    return cost


def time_gradient(coef) -> torch.Tensor:
    """``grad[B, n]`` = d(optimal snap cost)/d(duration of piece i) with the waypoints fixed, from the
    coefficients of one solve (minus the piece's Hamiltonian; an extension, see include/mst.h)."""
    dev = _abi.require_cuda()
    lib = _abi.load()
    coef = _f64(coef, dev)
    B, n, K, _ = coef.shape
    grad = torch.empty((B, n), dtype=torch.float64, device=dev)
    _abi.check(lib.mst_time_gradient(_ptr(coef), B, n, K, _ptr(grad), _stream_ptr()), "mst_time_gradient")
    return grad


# --------------------------------------------------------------------------- a8
def pack_pol_matrix(coef, dur, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``coef[B, n, K, 8]``, ``dur[B, n]`` -> float32 ``[B, n, 1 + 8K]`` rows
    ``[T | x0..x7 | y0..y7 | ...]`` (reference: path_to_pol's matrix,
    scripts/drones_pols_generator.py:63-77)."""
    dev = _abi.require_cuda()
    lib = _abi.load()
    coef = _f64(coef, dev)
    dur = _f64(dur, dev)
    B, n, K, _ = coef.shape
    if out is None:
        out = torch.empty((B, n, 1 + 8 * K), dtype=torch.float32, device=dev)
    else:
        _check_out("out", out, (B, n, 1 + 8 * K), torch.float32, dev)
    _abi.check(lib.mst_pack_pol_matrix(_ptr(coef), _ptr(dur), B, n, K, _ptr(out), _stream_ptr()), "mst_pack_pol_matrix")
    return out


def pol_matrix_csv(matrix) -> list:
    """float32 ``matrix[B, n, 1 + 8K]`` (``pack_pol_matrix``) -> list of ``B`` ``bytes`` objects, each the
    file ``np.savetxt(f, matrix[b], delimiter=",")`` writes (scripts/drones_pols_generator.py:79-81),
    byte for byte; all matrices formatted in one launch."""
    dev = _abi.require_cuda()
    lib = _abi.load()
    m = matrix if isinstance(matrix, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(matrix, dtype=np.float32))
    m = m.to(device=dev, dtype=torch.float32).contiguous()
    if m.dim() == 2:
        m = m[None]
    B, n, width = m.shape
    stride = lib.mst_csv_stride(n, width)
    text = torch.empty((B, stride), dtype=torch.uint8, device=dev)
    length = torch.empty((B,), dtype=torch.int32, device=dev)
    _abi.check(lib.mst_format_pol_matrix_csv(_ptr(m), B, n, width, _ptr(text), stride, _ptr(length), _stream_ptr()),
               "mst_format_pol_matrix_csv")
    text_h, len_h = text.cpu().numpy(), length.cpu().numpy()
    return [text_h[b, :len_h[b]].tobytes() for b in range(B)]


# --------------------------------------------------------------------------- a5/a6
def sample_batch(coef, dur, ts=None, S: Optional[int] = None, mode: str = "piecewise", deriv: int = 0,
                 return_status: bool = False):
    """Evaluate ``coef[B, n, K, 8]`` / ``dur[B, n]`` at sample times.

    ``ts`` may be ``[S]`` (shared), ``[B, S]`` (per trajectory) or None with ``S`` given
    (uniform ``t_s = s * sum(dur)/S``).  Returns ``out[B, S, K]`` (+ ``status[B, S]``).
    Reference: Polynomial.eval / PiecewisePolynomial.eval / Trajectory.eval,
    src/optimizations/uav_trajectory.py:17-26,119-127,154-169.
    """
    dev = _abi.require_cuda()
    lib = _abi.load()
    coef = _f64(coef, dev)
    dur = _f64(dur, dev)
    B, n, K, _ = coef.shape
    per = 0
    if ts is None:
        if S is None:
            raise ValueError("give ts or S")
        tsd = None
    else:
        tsd = _f64(ts, dev)
        per = 1 if tsd.dim() == 2 else 0
        S = tsd.shape[-1]
        if per and tsd.shape[0] != B:
            raise ValueError("per-trajectory ts must be [B, S]")
    out = torch.empty((B, S, K), dtype=torch.float64, device=dev)
    status = torch.empty((B, S), dtype=torch.uint8, device=dev)
    rc = lib.mst_sample_batch(_ptr(coef), _ptr(dur), B, n, K, _ptr(tsd), per, S, _MODES[mode], int(deriv),
                              _ptr(out), _ptr(status), _stream_ptr())
    _abi.check(rc, "mst_sample_batch")
    return (out, status) if return_status else out


def flat_outputs(coef, dur, ts=None, S: Optional[int] = None, mode: str = "trajectory",
                 return_status: bool = False):
    """Differential-flatness outputs ``[B, S, 13] = pos vel acc omega yaw`` of 4-axis
    trajectories (reference: Polynomial4D.eval, src/optimizations/uav_trajectory.py:66-101)."""
    dev = _abi.require_cuda()
    lib = _abi.load()
    coef = _f64(coef, dev)
    dur = _f64(dur, dev)
    B, n, K, _ = coef.shape
    if K != 4:
        raise ValueError("flat_outputs needs x, y, z, yaw (K = 4)")
    per = 0
    if ts is None:
        if S is None:
            raise ValueError("give ts or S")
        tsd = None
    else:
        tsd = _f64(ts, dev)
        per = 1 if tsd.dim() == 2 else 0
        S = tsd.shape[-1]
        if per and tsd.shape[0] != B:
            raise ValueError("per-trajectory ts must be [B, S]")
    out = torch.empty((B, S, 13), dtype=torch.float64, device=dev)
    status = torch.empty((B, S), dtype=torch.uint8, device=dev)
    rc = lib.mst_flat_outputs(_ptr(coef), _ptr(dur), B, n, _ptr(tsd), per, S, _MODES[mode], _ptr(out),
                              _ptr(status), _stream_ptr())
    _abi.check(rc, "mst_flat_outputs")
    return (out, status) if return_status else out


# --------------------------------------------------------------------------- a9
def formation_waypoints(rb_poses, offsets, K: int = 4) -> torch.Tensor:
    """``rb_poses[F, m, 4|7]`` x ``offsets[D, 3]`` -> ``wp[F*D, m, K]``
    (reference: transform(path), scripts/drones_traj_generator.py:56-89)."""
    dev = _abi.require_cuda()
    lib = _abi.load()
    rb = _f64(rb_poses, dev)
    off = _f64(offsets, dev)
    F, m, pd = rb.shape
    D = off.shape[0]
    wp = torch.empty((F * D, m, K), dtype=torch.float64, device=dev)
    rc = lib.mst_formation_waypoints(_ptr(rb), F, m, pd, _ptr(off), D, K, _ptr(wp), _stream_ptr())
    _abi.check(rc, "mst_formation_waypoints")
    return wp


# --------------------------------------------------------------------------- a10-a12
class Mesh:
    """Device-resident triangle mesh (reference: Fcl_mesh,
    src/RigidBodyPlanners/fcl_checker.py:13-59).  ``triangles`` is ``[T, 3, 3]`` float64,
    already rounded the way ``load_stl`` rounds (see ``meshio.ingest_mesh``)."""

    def __init__(self, triangles):
        _abi.require_cuda()
        lib = _abi.load()
        tri = np.ascontiguousarray(np.asarray(triangles, dtype=np.float64).reshape(-1, 3, 3))
        self.triangles = tri
        handle = ctypes.c_void_p()
        rc = lib.mst_mesh_create(tri.ctypes.data_as(ctypes.c_void_p), tri.shape[0], ctypes.byref(handle))
        _abi.check(rc, "mst_mesh_create")
        self._handle = handle

    @property
    def handle(self):
        return self._handle

    def __len__(self):
        return self.triangles.shape[0]

    def close(self):
        if getattr(self, "_handle", None):
            _abi.load().mst_mesh_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def collide_poses(robot: Mesh, env: Mesh, poses) -> torch.Tensor:
    """``poses[P, 3|4|7]`` -> ``hit[P]`` uint8 (reference: Fcl_checker.check_collision via
    isStateValid, fcl_checker.py:93-103, RB_planning_sep_coll_check.py:208-215)."""
    dev = _abi.require_cuda()
    lib = _abi.load()
    ps = _f64(poses, dev)
    if ps.dim() != 2:
        raise ValueError("poses must be [P, pose_dim]")
    P, pd = ps.shape
    hit = torch.empty((P,), dtype=torch.uint8, device=dev)
    rc = lib.mst_collide_poses(robot.handle, env.handle, _ptr(ps), P, pd, _ptr(hit), _stream_ptr())
    _abi.check(rc, "mst_collide_poses")
    return hit


def collide_pose_now(robot: Mesh, env: Mesh, pose) -> int:
    """One pose (3 | 4 | 7 doubles, host) -> 0 / 1, synchronously: the low-latency single query
    behind ``Fcl_checker.check_collision`` (one launch + one stream synchronisation, no tensors)."""
    _abi.require_cuda()
    lib = _abi.load()
    p = np.ascontiguousarray(pose, dtype=np.float64).reshape(-1)
    if p.size not in (3, 4, 7):
        raise ValueError("pose must have 3, 4 or 7 entries")
    out = ctypes.c_int(-1)
    rc = lib.mst_collide_pose_sync(robot.handle, env.handle, p.ctypes.data_as(ctypes.c_void_p), int(p.size),
                                   ctypes.byref(out))
    _abi.check(rc, "mst_collide_pose_sync")
    return int(out.value)


def collide_motions(robot: Mesh, env: Mesh, state_a, state_b, steps: int) -> torch.Tensor:
    """``state_a[M, 4]``, ``state_b[M, 4]`` (x, y, z, yaw) -> ``invalid[M]`` uint8: 1 iff one of the
    ``steps`` states interpolated at fractions ``j/steps`` (j = 1..steps) collides.  All
    interpolated states of all candidate motions in one launch."""
    dev = _abi.require_cuda()
    lib = _abi.load()
    a = _f64(state_a, dev)
    b = _f64(state_b, dev)
    if a.shape != b.shape or a.dim() != 2 or a.shape[1] != 4:
        raise ValueError("states must be [M, 4] (x, y, z, yaw)")
    M = a.shape[0]
    invalid = torch.empty((M,), dtype=torch.uint8, device=dev)
    rc = lib.mst_collide_motions(robot.handle, env.handle, _ptr(a), _ptr(b), M, int(steps), _ptr(invalid), _stream_ptr())
    _abi.check(rc, "mst_collide_motions")
    return invalid


def collide_trajectories(coef, dur, S: int, robot: Mesh, env: Mesh):
    """Sample ``S`` uniform times of every trajectory and collision-check the robot mesh placed
    there; returns ``hit[B, S]``, ``any_hit[B]`` (the pipeline's second kernel on its own)."""
    dev = _abi.require_cuda()
    lib = _abi.load()
    coef = _f64(coef, dev)
    dur = _f64(dur, dev)
    B, n, K, _ = coef.shape
    hit = torch.empty((B, S), dtype=torch.uint8, device=dev)
    any_hit = torch.empty((B,), dtype=torch.uint8, device=dev)
    rc = lib.mst_collide_trajectories(_ptr(coef), _ptr(dur), B, n, K, S, robot.handle, env.handle, _ptr(hit),
                                      _ptr(any_hit), _stream_ptr())
    _abi.check(rc, "mst_collide_trajectories")
    return hit, any_hit


# --------------------------------------------------------------------------- fused pipeline
class PipelineResult:
    __slots__ = ("coef", "dur", "info", "hit", "any_hit")

    def __init__(self, coef, dur, info, hit, any_hit):
        self.coef, self.dur, self.info, self.hit, self.any_hit = coef, dur, info, hit, any_hit


def pipeline(wp, t, S: int, robot: Mesh, env: Mesh, share_time_group: int = 1, solver: str = "auto",
             out: Optional[PipelineResult] = None, pol_matrix: Optional[torch.Tensor] = None) -> PipelineResult:
    """Solve, sample ``S`` uniform times per trajectory, place the robot mesh at every
    sample and collision-check it against ``env``.  Returns coefficients, durations,
    solver status, ``hit[B, S]`` and ``any_hit[B]``.  ``pol_matrix`` (float32 ``[B, n, 1 + 8K]``,
    optional): also filled with the reference's wire format (``pack_pol_matrix`` of the results),
    written by the solver kernel itself."""
    dev = _abi.require_cuda()
    lib = _abi.load()
    wp = _f64(wp, dev)
    t = _f64(t, dev)
    B, m, K = wp.shape
    n = m - 1
    G = int(share_time_group)
    if n < 1:
        raise IndexError("need at least two waypoints")
    if G < 1 or B % G != 0 or t.shape[0] != B // G or t.shape[1] != m:
        raise ValueError("t must be [B/share_time_group, n+1]")
    if out is not None:
        _check_out("out.coef", out.coef, (B, n, K, 8), torch.float64, dev)
        _check_out("out.dur", out.dur, (B, n), torch.float64, dev)
        _check_out("out.info", out.info, (B,), torch.int32, dev)
        _check_out("out.hit", out.hit, (B, S), torch.uint8, dev)
        _check_out("out.any_hit", out.any_hit, (B,), torch.uint8, dev)
    if out is None:
        out = PipelineResult(torch.empty((B, n, K, 8), dtype=torch.float64, device=dev),
                             torch.empty((B, n), dtype=torch.float64, device=dev),
                             torch.empty((B,), dtype=torch.int32, device=dev),
                             torch.empty((B, S), dtype=torch.uint8, device=dev),
                             torch.empty((B,), dtype=torch.uint8, device=dev))
    ws = torch.empty((max(1, lib.mst_pipeline_workspace_bytes(B, n, K, G, S)),), dtype=torch.uint8, device=dev)
    if pol_matrix is not None:
        _check_out("pol_matrix", pol_matrix, (B, n, 1 + 8 * K), torch.float32, dev)
        rc = lib.mst_pipeline_packed(_ptr(wp), _ptr(t), B, n, K, G, _SOLVERS[solver], S, robot.handle, env.handle,
                                     _ptr(out.coef), _ptr(out.dur), _ptr(out.info), _ptr(out.hit), _ptr(out.any_hit),
                                     _ptr(pol_matrix), _ptr(ws), _stream_ptr())
        _abi.check(rc, "mst_pipeline_packed")
        return out
    rc = lib.mst_pipeline(_ptr(wp), _ptr(t), B, n, K, G, _SOLVERS[solver], S, robot.handle, env.handle,
                          _ptr(out.coef), _ptr(out.dur), _ptr(out.info), _ptr(out.hit), _ptr(out.any_hit),
                          _ptr(ws), _stream_ptr())
    _abi.check(rc, "mst_pipeline")
    return out


def make_wire_targets(pol_matrix=None, hit=None, any_hit=None, row_offset: int = 0):
    """Build the ``mst_wire_targets`` block of ``pipeline_wire``: lists of CUDA tensors (this rank's
    gather buffer and the peer mappings of the other ranks' buffers, e.g. from
    ``torch.distributed._symmetric_memory``), one entry per destination.  Keeps the pointer arrays
    alive on the returned object."""
    lists = [x for x in (pol_matrix, hit, any_hit) if x is not None]
    if not lists or len({len(x) for x in lists}) != 1:
        raise ValueError("give pol_matrix and / or hit + any_hit, one tensor per destination each")
    if (hit is None) != (any_hit is None):
        raise ValueError("hit and any_hit travel together")
    count = len(lists[0])
    block = _abi.WireTargets()
    block.count = count
    block.row_offset = int(row_offset)
    keep = []
    for name, tensors in (("pol_matrix", pol_matrix), ("hit", hit), ("any_hit", any_hit)):
        if tensors is None:
            setattr(block, name, None)
            continue
        arr = (ctypes.c_void_p * count)(*[t.data_ptr() for t in tensors])
        keep.append((arr, tensors))
        setattr(block, name, ctypes.cast(arr, ctypes.POINTER(ctypes.c_void_p)))
    block._keep = keep
    return block


def pipeline_wire(wp, t, S: int, robot: Mesh, env: Mesh, wire, share_time_group: int = 1,
                  out: Optional[PipelineResult] = None) -> PipelineResult:
    """``pipeline`` that ALSO stores the float32 polynomial matrix (the reference's wire format,
    scripts/drones_pols_generator.py:63-77) and / or the flags through the destination pointers of
    ``wire`` (``make_wire_targets``) from inside the kernel — the multi-GPU gather by NVLink peer
    stores.  The caller closes the step with a cross-rank barrier."""
    dev = _abi.require_cuda()
    lib = _abi.load()
    wp = _f64(wp, dev)
    t = _f64(t, dev)
    B, m, K = wp.shape
    n = m - 1
    G = int(share_time_group)
    if n < 1:
        raise IndexError("need at least two waypoints")
    if G < 1 or B % G != 0 or t.shape[0] != B // G or t.shape[1] != m:
        raise ValueError("t must be [B/share_time_group, n+1]")
    if out is not None:
        _check_out("out.coef", out.coef, (B, n, K, 8), torch.float64, dev)
        _check_out("out.dur", out.dur, (B, n), torch.float64, dev)
        _check_out("out.info", out.info, (B,), torch.int32, dev)
        _check_out("out.hit", out.hit, (B, S), torch.uint8, dev)
        _check_out("out.any_hit", out.any_hit, (B,), torch.uint8, dev)
    else:
        out = PipelineResult(torch.empty((B, n, K, 8), dtype=torch.float64, device=dev),
                             torch.empty((B, n), dtype=torch.float64, device=dev),
                             torch.empty((B,), dtype=torch.int32, device=dev),
                             torch.empty((B, S), dtype=torch.uint8, device=dev),
                             torch.empty((B,), dtype=torch.uint8, device=dev))
    ws = torch.empty((max(1, lib.mst_pipeline_workspace_bytes(B, n, K, G, S)),), dtype=torch.uint8, device=dev)
    rc = lib.mst_pipeline_wire(_ptr(wp), _ptr(t), B, n, K, G, S, robot.handle, env.handle, _ptr(out.coef), _ptr(out.dur),
                               _ptr(out.info), _ptr(out.hit), _ptr(out.any_hit), ctypes.byref(wire), _ptr(ws),
                               _stream_ptr())
    _abi.check(rc, "mst_pipeline_wire")
    return out


# --------------------------------------------------------------------------- single evaluations
class _SyncSlot:
    """Per-thread staging for SINGLE evaluations through the drop-in surface (``Polynomial.eval``,
    ``PiecewisePolynomial.eval``, ``Trajectory.eval`` called one time value at a time): inputs and
    outputs live in pinned host memory, which the kernels read and write in place (unified
    addressing), so a call is one launch and one stream synchronisation — no tensor allocation,
    no staging copies."""

    def __init__(self):
        dev = _abi.require_cuda()
        self.stream = torch.cuda.Stream(device=dev)
        self.inp = torch.zeros(64, dtype=torch.float64).pin_memory()    # [0:32] coefficients, [32] duration, [40] t
        self.out = torch.zeros(16, dtype=torch.float64).pin_memory()
        self.status = torch.zeros(8, dtype=torch.uint8).pin_memory()
        self.inp_np, self.out_np, self.status_np = self.inp.numpy(), self.out.numpy(), self.status.numpy()
        self.p_coef = self.inp.data_ptr()
        self.p_dur = self.inp.data_ptr() + 32 * 8
        self.p_t = self.inp.data_ptr() + 40 * 8
        self.p_out, self.p_status = self.out.data_ptr(), self.status.data_ptr()
        self.inp_np[32] = 1.0


_sync_local = __import__("threading").local()


def _sync_slot() -> _SyncSlot:
    slot = getattr(_sync_local, "slot", None)
    if slot is None:
        slot = _sync_local.slot = _SyncSlot()
    return slot


def sample_piece_now(coefs, t: float, deriv: int = 0) -> float:
    """One polynomial (<= 8 ascending coefficients, host) at one time: ``Polynomial.eval``."""
    lib = _abi.load()
    slot = _sync_slot()
    c = slot.inp_np
    c[:8] = 0.0
    c[:len(coefs)] = coefs
    c[40] = t
    rc = lib.mst_sample_batch(slot.p_coef, slot.p_dur, 1, 1, 1, slot.p_t, 0, 1, _abi.SAMPLE_PIECEWISE, int(deriv),
                              slot.p_out, slot.p_status, slot.stream.cuda_stream)
    _abi.check(rc, "mst_sample_batch")
    slot.stream.synchronize()
    return float(slot.out_np[0])


def sample_now(coef_dev: torch.Tensor, dur_dev: torch.Tensor, t: float, mode: str = "piecewise", deriv: int = 0,
               flat: bool = False):
    """One time value against device-resident ``coef[1, n, K, 8]`` / ``dur[1, n]``: returns
    ``(values[K] or flat outputs[13], status)`` as host numbers (``PiecewisePolynomial.eval`` /
    ``Trajectory.eval`` one call at a time)."""
    lib = _abi.load()
    slot = _sync_slot()
    slot.inp_np[40] = t
    _, n, K, _ = coef_dev.shape
    # the arrays may have been produced on another stream: order this stream behind it
    slot.stream.wait_stream(torch.cuda.current_stream())
    if flat:
        rc = lib.mst_flat_outputs(coef_dev.data_ptr(), dur_dev.data_ptr(), 1, n, slot.p_t, 0, 1, _MODES[mode], slot.p_out,
                                  slot.p_status, slot.stream.cuda_stream)
        _abi.check(rc, "mst_flat_outputs")
    else:
        rc = lib.mst_sample_batch(coef_dev.data_ptr(), dur_dev.data_ptr(), 1, n, K, slot.p_t, 0, 1, _MODES[mode], int(deriv),
                                  slot.p_out, slot.p_status, slot.stream.cuda_stream)
        _abi.check(rc, "mst_sample_batch")
    slot.stream.synchronize()
    return slot.out_np[:13 if flat else K].copy(), int(slot.status_np[0])
