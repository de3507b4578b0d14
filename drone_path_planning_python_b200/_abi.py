"""ctypes binding of libmst.so (include/mst.h).

There is no CPU fallback: if the library is missing, cannot be loaded, or a CUDA device
is not available, every compute call raises.  ``load()`` only dlopens the library — that
works on a machine without a GPU, which is what the CPU test tier checks (symbols present,
argument validation) — compute entry points need a CUDA device.
"""
from __future__ import annotations

import ctypes
import os
import threading

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libmst.so")

c_double_p = ctypes.c_void_p  # device pointers travel as integers
c_void_p = ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/mst.h one to one
PROTOTYPES = {
    "mst_version": (ctypes.c_int, []),
    "mst_strerror": (ctypes.c_char_p, [ctypes.c_int]),
    "mst_last_cuda_error": (ctypes.c_char_p, []),
    "mst_time_power_rows": (ctypes.c_int, [c_void_p, ctypes.c_int, c_void_p, c_void_p]),
    "mst_poly_derivative": (ctypes.c_int, [c_void_p, ctypes.c_int, ctypes.c_int, c_void_p, c_void_p]),
    "mst_poly_terms_at_t": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int, ctypes.c_int, c_void_p, c_void_p]),
    "mst_solve_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int] * 4),
    "mst_solve_batch": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                       ctypes.c_int, ctypes.c_int, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p]),
    "mst_snap_cost": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_void_p, c_void_p]),
    "mst_time_gradient": (ctypes.c_int, [c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_void_p, c_void_p]),
    "mst_pack_pol_matrix": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           c_void_p, c_void_p]),
    "mst_csv_stride": (ctypes.c_size_t, [ctypes.c_int, ctypes.c_int]),
    "mst_format_pol_matrix_csv": (ctypes.c_int, [c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_void_p,
                                                 ctypes.c_longlong, c_void_p, c_void_p]),
    "mst_sample_batch": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        c_void_p, c_void_p, c_void_p]),
    "mst_flat_outputs": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int, ctypes.c_int, c_void_p,
                                        ctypes.c_int, ctypes.c_int, ctypes.c_int, c_void_p, c_void_p,
                                        c_void_p]),
    "mst_formation_waypoints": (ctypes.c_int, [c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               c_void_p, ctypes.c_int, ctypes.c_int, c_void_p, c_void_p]),
    "mst_mesh_create": (ctypes.c_int, [c_void_p, ctypes.c_int, ctypes.POINTER(c_void_p)]),
    "mst_mesh_destroy": (ctypes.c_int, [c_void_p]),
    "mst_mesh_triangle_count": (ctypes.c_int, [c_void_p]),
    "mst_collide_poses": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_int, ctypes.c_int,
                                         c_void_p, c_void_p]),
    "mst_collide_pose_sync": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]),
    "mst_collide_motions": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_int, ctypes.c_int,
                                           c_void_p, c_void_p]),
    "mst_collide_trajectories": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mst_pipeline_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int] * 5),
    "mst_pipeline_launch_count": (ctypes.c_int, [ctypes.c_int] * 6),
    "mst_pipeline": (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_int, ctypes.c_int, ctypes.c_int, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p]),
}


class WireTargets(ctypes.Structure):
    """``mst_wire_targets`` of include/mst.h: host arrays of device base pointers."""
    _fields_ = [("count", ctypes.c_int),
                ("pol_matrix", ctypes.POINTER(ctypes.c_void_p)),
                ("hit", ctypes.POINTER(ctypes.c_void_p)),
                ("any_hit", ctypes.POINTER(ctypes.c_void_p)),
                ("row_offset", ctypes.c_longlong)]


PROTOTYPES["mst_pipeline_packed"] = (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                    ctypes.c_int, ctypes.c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p])
PROTOTYPES["mst_pipeline_stage"] = (ctypes.c_int, [ctypes.c_int, c_void_p, c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                   ctypes.c_int, ctypes.c_int, ctypes.c_int, c_void_p, c_void_p,
                                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                                   c_void_p])
PROTOTYPES["mst_pipeline_wire"] = (ctypes.c_int, [c_void_p, c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                  ctypes.c_int, ctypes.c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                                  c_void_p, c_void_p, c_void_p, ctypes.POINTER(WireTargets), c_void_p,
                                                  c_void_p])

MST_OK = 0
SOLVER_AUTO, SOLVER_BANDED_LU, SOLVER_CONDENSED, SOLVER_AUTO_ONE_PASS = 0, 1, 2, 3
SAMPLE_PIECEWISE, SAMPLE_TRAJECTORY = 0, 1
INFO_DECREASING, INFO_NONFINITE, INFO_DECLINED = -1, -2, -3

_lib = None
_lock = threading.Lock()


class MstError(RuntimeError):
    """A libmst call returned a non-zero status."""


def load() -> ctypes.CDLL:
    """dlopen libmst.so and attach the prototypes.  Raises if the library is absent —
    build it with ``python -m drone_path_planning_python_b200.build`` (or
    ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libmst.so not found at %s: the CUDA library is the only compute path of this "
                "package (no CPU fallback). Build it with "
                "`python -m drone_path_planning_python_b200.build`." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in PROTOTYPES.items():
            fn = getattr(lib, name)  # AttributeError here means header and library disagree
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
        return lib


def check(rc: int, what: str) -> None:
    if rc == MST_OK:
        return
    lib = load()
    msg = lib.mst_strerror(rc).decode()
    detail = lib.mst_last_cuda_error().decode()
    raise MstError("%s failed: %s (%d)%s" % (what, msg, rc, (": " + detail) if detail and rc == -3 else ""))


def require_cuda():
    """The device every call runs on; raises when there is none (no CPU fallback)."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("drone_path_planning_python_b200 needs a CUDA device (B200, sm_100a); "
                           "there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())
