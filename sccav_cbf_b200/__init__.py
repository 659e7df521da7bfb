"""sccav_cbf_b200 -- B200-native batched CBF-QP safety filter (hand-written sm_100a CUDA behind a
C-ABI, include/sccav_cbf.h) with the Python class API of Safety-Critical-Control-WIRIN/sccav_cbf.

    from sccav_cbf_b200 import DBM_CBF_2DS, CollisionCone2D, Ellipse2D, PolyLane, LateralStanley, PID1

mirrors ``from cbf.cbf import ...`` / ``from cbf.obstacles import ...`` / ``from cbf.controllers import ...``
of the reference.  ``ops`` holds the tensor-level operators, ``rollout.ClosedLoopRollout`` the
persistent closed-loop kernel front end.  There is no CPU compute path: importing works anywhere
(so host logic can be tested), every operator raises without the CUDA library and a GPU.
"""
from .cbf import DBM_CBF_2DS, DUM_CBF_2DS, KBM_VC_CBF2D, SADBM_CBF_2DS
from .controllers import PID1, LateralStanley
from .euclid import Point2, Point3, Vector2, Vector3
from .geometry import Rotation, Transform
from .obstacles import (BatchedObstacleList2D, BoundingBox, CollisionCone2D, Ellipse2D, Obstacle2DBase, Obstacle2DTypes,
                        ObstacleList2D, PolyLane)
from .utils import ZERO_TOL, Timer, TimerError, normalize_angle, saturation, sigmoid, vec_norm

__all__ = [
    "DBM_CBF_2DS", "DUM_CBF_2DS", "KBM_VC_CBF2D", "SADBM_CBF_2DS", "LateralStanley", "PID1", "Vector2", "Point2", "Vector3", "Point3", "Rotation",
    "Transform", "BatchedObstacleList2D", "BoundingBox", "CollisionCone2D", "Ellipse2D", "Obstacle2DBase", "Obstacle2DTypes", "ObstacleList2D",
    "PolyLane", "ZERO_TOL", "Timer", "TimerError", "normalize_angle", "saturation", "sigmoid", "vec_norm",
]
__version__ = "0.1.0"
