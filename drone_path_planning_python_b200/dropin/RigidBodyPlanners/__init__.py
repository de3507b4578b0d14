"""B200 drop-in slice of the reference package ``RigidBodyPlanners``.

Only ``fcl_checker`` (the collision half of the hot path) is replaced.  ``__path__`` is
extended over every other ``RigidBodyPlanners`` directory on ``sys.path`` so that, with this
directory placed in front of the reference's ``src``, ``RigidBodyPlanners.fcl_checker`` is the
CUDA-backed module while ``RigidBodyPlanners.RB_planning_sep_coll_check`` and
``RigidBodyPlanners.frameTransforms`` remain the reference's own, unmodified files (they need
OMPL / ROS and are out of scope).  The reference's ``__init__`` re-exports those two modules'
names (src/RigidBodyPlanners/__init__.py:1-2); that is repeated here when they import.
"""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)

try:  # present only next to a reference checkout with OMPL + ROS installed
    from .RB_planning_sep_coll_check import *  # noqa: F401,F403
    from .frameTransforms import *  # noqa: F401,F403
except ImportError:
    pass
