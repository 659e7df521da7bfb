"""Host-side course generation (once per scenario set; not on the per-step path).

The reference builds its target course with the PythonRobotics natural cubic spline
(test_scripts/PathPlanning/CubicSpline/cubic_spline_planner.py:12-190, called at
stanley_controller_ellipse.py:584-588): two 1-D natural splines x(s), y(s) over the cumulative
chord length s, sampled every ``ds``; yaw = atan2(y'(s), x'(s)).  This module produces the same
samples (bit-for-bit on the reference's way-points; see tests/test_course.py) so that obstacle
placement "at 75 % of the course" means the same point.
"""
from __future__ import annotations

import math
from typing import Sequence, Tuple

import numpy as np

CONFIG1_WAYPOINTS = ([0.0, 100.0, 100.0, 50.0, 60.0], [0.0, 0.0, -30.0, -20.0, 0.0])   # sce.py:584-585


def _natural_spline(knots: Sequence[float], vals: Sequence[float]):
    """Coefficients (a, b, c, d) per segment of the natural cubic spline through (knots, vals)."""
    n = len(knots)
    a = [float(v) for v in vals]
    h = [float(knots[i + 1]) - float(knots[i]) for i in range(n - 1)]
    A = np.zeros((n, n))
    rhs = np.zeros(n)
    A[0, 0] = 1.0
    A[n - 1, n - 1] = 1.0
    for i in range(1, n - 1):
        A[i, i - 1] = h[i - 1]
        A[i, i] = 2.0 * (h[i - 1] + h[i])
        A[i, i + 1] = h[i]
        rhs[i] = 3.0 * (a[i + 1] - a[i]) / h[i] - 3.0 * (a[i] - a[i - 1]) / h[i - 1]
    c = np.linalg.solve(A, rhs)
    b = [(a[i + 1] - a[i]) / h[i] - h[i] * (c[i + 1] + 2.0 * c[i]) / 3.0 for i in range(n - 1)]
    d = [(c[i + 1] - c[i]) / (3.0 * h[i]) for i in range(n - 1)]
    return a, b, [float(v) for v in c], d


def spline_course(wx: Sequence[float], wy: Sequence[float], ds: float = 0.1) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(cx, cy, cyaw) float64 arrays sampled every ``ds`` along the spline through the way-points."""
    seg = np.hypot(np.diff(wx), np.diff(wy))
    s = [0.0]
    s.extend(float(v) for v in np.cumsum(seg))
    ax_, bx, cx_, dx_ = _natural_spline(s, wx)
    ay_, by, cy_, dy_ = _natural_spline(s, wy)
    ts = np.arange(0, s[-1], ds)
    nseg = len(s) - 1
    cx = np.empty(len(ts)); cy = np.empty(len(ts)); cyaw = np.empty(len(ts))
    i = 0
    for j, t in enumerate(ts):
        t = float(t)
        while i + 1 < nseg and t >= s[i + 1]:
            i += 1
        u = t - s[i]
        u2 = u ** 2.0
        u3 = u ** 3.0
        cx[j] = ax_[i] + bx[i] * u + cx_[i] * u2 + dx_[i] * u3
        cy[j] = ay_[i] + by[i] * u + cy_[i] * u2 + dy_[i] * u3
        gx = bx[i] + 2.0 * cx_[i] * u + 3.0 * dx_[i] * u2
        gy = by[i] + 2.0 * cy_[i] * u + 3.0 * dy_[i] * u2
        cyaw[j] = math.atan2(gy, gx)
    return cx, cy, cyaw


def config1_course(ds: float = 0.1):
    """The course of stanley_controller_ellipse.py main(): P = 2034 points."""
    return spline_course(CONFIG1_WAYPOINTS[0], CONFIG1_WAYPOINTS[1], ds)
