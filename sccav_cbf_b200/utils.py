"""Helpers of the reference's cbf/utils.py with the same names and behaviour, accepting Python
numbers or torch tensors (one value per vehicle).

ZERO_TOL (cbf/utils.py:27) is part of the barrier arithmetic, not a tunable tolerance: the CUDA
path hard-codes the same 1e-3 (csrc/path.cuh, Real<T>::zero_tol).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .euclid import Point3, Vector3
from .geometry import Rotation, Transform  # noqa: F401  (re-exported like cbf/utils.py:22)

ZERO_TOL = 1e-3                                                    # cbf/utils.py:27


class TimerError(Exception):
    """Custom Exception for Timer related errors (cbf/utils.py:29-31)."""


class Timer:
    """Monotone timestamp guard (cbf/utils.py:33-48).  The reference's setter reads the property
    before it exists; here the first assignment is accepted and later ones must not decrease."""

    def __init__(self, timestamp=0.0):
        self.timestamp_ = timestamp

    @property
    def timestamp(self):
        return self.timestamp_

    @timestamp.setter
    def timestamp(self, value):
        if self.timestamp_ > value:
            raise TimerError("Negative time or Decreasing Timestamp trend detected. Please make sure that the "
                             "Timestamp is monotonically increasing when manually set.")
        self.timestamp_ = value


def convert_LH_to_RH(flipped_axis="y", *args):
    """cbf/utils.py:51-91 (returns the conversion of the FIRST argument, like the reference)."""
    if flipped_axis not in ("x", "y", "z"):
        raise ValueError("Invalid input to the flipped_axis argument. Expected values from ['x', 'y', 'z']. Received "
                         + str(flipped_axis))
    for arg in args:
        if isinstance(arg, Rotation):
            return Rotation(arg.roll, -arg.pitch, -arg.yaw)
        if isinstance(arg, Vector3):
            sx, sy, sz = {"x": (-1, 1, 1), "y": (1, -1, 1), "z": (1, 1, -1)}[flipped_axis]
            cls = Point3 if isinstance(arg, Point3) else Vector3
            return cls(sx * arg.x, sy * arg.y, sz * arg.z)
        raise TypeError("Invalid input. Expected euclid.Vector3, euclid.Point3 or cbf.geometry.Rotation objects. "
                        "Received " + type(arg).__name__)


def normalize_angle(angle):
    """Normalize an angle to [-pi, pi] (cbf/utils.py:93-106: subtract / add 2 pi while outside,
    strict inequalities).  Tensors: the same rule element-wise; |angle| > 1e4 has whole turns
    removed first, as in the CUDA path."""
    if isinstance(angle, torch.Tensor):
        two_pi = 2.0 * math.pi
        a = torch.where(angle.abs() <= 1e4, angle, angle - two_pi * torch.round(angle / two_pi))
        for _ in range(4):                       # |a| <= 1e4 needs at most ~1600 turns: remove them in bulk first
            k = torch.where(a > math.pi, torch.ceil((a - math.pi) / two_pi), torch.zeros_like(a))
            k = torch.where(a < -math.pi, -torch.ceil((-a - math.pi) / two_pi), k)
            a = a - two_pi * k
        return a
    if not (abs(angle) <= 1e4):
        angle = angle - (2.0 * np.pi) * float(np.rint(angle / (2.0 * np.pi)))
    while angle > np.pi:
        angle -= 2.0 * np.pi
    while angle < -np.pi:
        angle += 2.0 * np.pi
    return angle


def sigmoid(x):
    if isinstance(x, torch.Tensor):
        return torch.sigmoid(x)
    return 1 / (1 + np.exp(-x))


def saturation(x, x_min, x_max):
    """cbf/utils.py:108-114."""
    if isinstance(x, torch.Tensor):
        return torch.clamp(x, min=x_min, max=x_max)
    if x > x_max:
        return x_max
    elif x < x_min:
        return x_min
    return x


def get_closest_idx(x, x_list):
    """cbf/utils.py:116-118."""
    dx = [abs(x - ix) for ix in x_list]
    return int(np.argmin(dx))


def vec_norm(x):
    """sqrt(x^T x) (cbf/utils.py:120-121) for a sequence / array / tensor of components."""
    if isinstance(x, torch.Tensor):
        return torch.sqrt((x * x).sum(dim=0))
    return math.sqrt(sum(float(v) * float(v) for v in np.asarray(x, dtype=np.float64).reshape(-1)))
