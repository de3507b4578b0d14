"""B200 drop-in for ``trajectory_visualising.uav_trajectory``
(reference: src/trajectory_visualising/uav_trajectory.py — the crazyswarm original that
``optimizations.uav_trajectory`` extends: ``normalize``, ``Polynomial``, ``TrajectoryOutput``,
``Polynomial4D``, ``Trajectory`` with the same ``skiprows=1`` loader, :96-109).  The classes are
the CUDA-backed ones of the ``optimizations`` drop-in; nothing is re-implemented here."""
import importlib.util as _ilu
import os as _os

import numpy as np  # noqa: F401  (the reference module exposes it through its star import)

_path = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "optimizations",
                      "uav_trajectory.py")
_spec = _ilu.spec_from_file_location("_mst_dropin_uav_trajectory", _path)
_mod = _ilu.module_from_spec(_spec)
_spec.loader.exec_module(_mod)

normalize = _mod.normalize
Polynomial = _mod.Polynomial
TrajectoryOutput = _mod.TrajectoryOutput
Polynomial4D = _mod.Polynomial4D
Trajectory = _mod.Trajectory
