"""B200 drop-in for the reference package ``trajectory_visualising``
(src/trajectory_visualising/__init__.py:1-2 exports ``Trajectory``, ``TrajectoryOutput`` and
``get_nav_path_msg``; consumed by scripts/path_vis.py:6-7)."""
from .uav_trajectory import Trajectory, TrajectoryOutput  # noqa: F401
from .visualization import get_nav_path_msg  # noqa: F401
