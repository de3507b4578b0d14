"""B200-native batched minimum-snap trajectory generation + mesh collision checking.

A from-scratch sm_100a implementation of the data-parallel hot path of
mjmyt/drone_path_planning_python: ``batch`` is the tensor-level
throughput API over the C ABI in ``include/mst.h``; ``dropin/`` mirrors the reference's
own Python call surface (``optimizations``, ``RigidBodyPlanners.fcl_checker``,
``scripts``).  The CUDA library is the only compute path: without ``libmst.so`` and a CUDA
device every call raises.
"""
from . import _abi  # noqa: F401
from .batch import (Mesh, PipelineResult, make_wire_targets, pipeline_wire, collide_motions, collide_pose_now, collide_poses, collide_trajectories, flat_outputs, formation_waypoints,  # noqa: F401
                    pack_pol_matrix, pipeline, pol_matrix_csv, sample_now, sample_piece_now, poly_derivative, poly_terms_at_t, sample_batch, snap_cost, solve_batch,
                    time_gradient, time_power_rows)
from .time_allocation import optimize_time_allocation  # noqa: F401

__version__ = "0.1.0"


def dropin_path() -> str:
    """Directory to put in front of ``sys.path`` (or PYTHONPATH) so that
    ``import optimizations`` / ``import RigidBodyPlanners`` resolve to the B200 drop-ins."""
    import os
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "dropin")
