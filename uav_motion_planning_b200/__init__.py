"""uav_motion_planning_b200 — B200-native batched kinodynamic A* + minimum-jerk/snap QP.

Only what the hot path needs: `csrc/` (CUDA kernels + the C-ABI of include/uavmp.h) and thin host mirrors of the
reference's two classes (KinoAstar, MinimumControl).  The CUDA library is the only implementation.
"""
from ._lib import Context, KinoParams, OsqpSettings, UavmpError, load  # noqa: F401
from .kino_astar import NO_PATH_FOUND, REACH_END, KinoAstar  # noqa: F401
from .a_star import Astar  # noqa: F401
from .rrt_star import RRTStar  # noqa: F401
from .mapgen import make_world, sample_queries  # noqa: F401

try:  # lands with the QP kernel
    from .minimum_control import MinimumControl  # noqa: F401
except ImportError:  # pragma: no cover
    pass
