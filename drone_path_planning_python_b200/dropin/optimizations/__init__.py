"""B200 drop-in for the reference package ``optimizations``
(src/optimizations/__init__.py:1-2 exports every name of ``uav_trajectory`` — ``np`` included,
there is no ``__all__`` — plus ``calculate_trajectory4D``)."""
from .uav_trajectory import *  # noqa: F401,F403
from .uav_trajectory import np  # noqa: F401  (the reference leaks it; drones_pols_generator.py relies on star-import)
from .calculatingTrajectories import calculate_trajectory4D  # noqa: F401
