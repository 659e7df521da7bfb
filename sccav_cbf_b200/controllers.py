"""Nominal controllers with the class API of the reference's cbf/controllers.py.

``LateralStanley.control`` is ONE launch of the Stanley kernel (csrc/kernels.cuh stanley_kernel):
exact global nearest way-point over the whole trajectory, front-axle error, monotone index clamp,
``delta = normalize(yaw_des[idx] - yaw) + atan2(k e, v + ks)`` (cbf/controllers.py:66-151).  The
reference also computes -- and discards -- a second cross-track error and a "manual" reference yaw
(:121-138); those have no effect on the returned values and are not computed.

``PID1`` is three multiply-adds per call and stays element-wise on whatever its inputs are.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from ._batch import as_state, as_vec, cuda_device, to_output
from .euclid import Vector2


class LateralStanley:
    """Stanley steering control (cbf/controllers.py:29-151)."""

    def __init__(self, lr=2.0, lf=2.0, k=0.5, ks=0.01):
        self.__lr = lr
        self.__lf = lf
        self.__last_target_idx = None          # int32 [N] on the GPU once the batch size is known
        self.__k = k
        self.__ks = ks
        self.__state = None
        self.__course = None
        self.__trajectory = None
        self.fx = None
        self.fy = None

    def update_state(self, x, y, yaw, v):
        self.__state = (x, y, yaw, v)

    def set_gains(self, k, ks):
        self.__k = k
        self.__ks = ks

    def set_trajectory(self, trajectory):
        """trajectory: sequence of (x, y, yaw, v) way-points (cbf/controllers.py:61-67), or a [P, 4] array."""
        self.__trajectory = trajectory
        t = np.asarray([[float(p[0]), float(p[1]), float(p[2]), float(p[3])] for p in trajectory], dtype=np.float64)
        self.__course_host = t
        self.__course = None

    def reset(self):
        """Forget the last target index (a new run over the same trajectory)."""
        self.__last_target_idx = None

    def __device_course(self, dtype, device):
        if self.__course is None or self.__course[0].dtype != dtype or self.__course[0].device != device:
            t = torch.from_numpy(self.__course_host).to(device=device, dtype=dtype)
            self.__course = (t[:, 0].contiguous(), t[:, 1].contiguous(), t[:, 2].contiguous())
        return self.__course

    def control(self, trajectory=None, front_coords=None, initial_yaw=0):
        """-> (delta, target_idx).  ``front_coords``: externally measured front-axle position (a
        ``Vector2``; TypeError otherwise, cbf/controllers.py:105-108)."""
        if front_coords is not None and not isinstance(front_coords, Vector2):
            raise TypeError("Coordinates of the front wheel must be specified as an euclid.Vector2() object.")
        if trajectory is not None:
            self.set_trajectory(trajectory)
        if self.__trajectory is None:
            raise AttributeError("set_trajectory() has not been called")
        if self.__state is None:
            raise AttributeError("update_state() has not been called")
        state, scalar = as_state(self.__state)
        N = state.shape[1]
        dev, dt = state.device, state.dtype
        front = None
        if front_coords is not None:
            front = torch.stack([as_vec(front_coords.x, N, dt, dev), as_vec(front_coords.y, N, dt, dev)]).contiguous()
        if self.__last_target_idx is None or self.__last_target_idx.shape[0] != N or self.__last_target_idx.device != dev:
            self.__last_target_idx = torch.zeros((N,), dtype=torch.int32, device=dev)
        prm = ops.make_params(L=float(self.__lf), k_stanley=float(self.__k), ks_stanley=float(self.__ks))
        delta = ops.stanley_control(prm, state, self.__device_course(dt, dev), self.__last_target_idx, front=front)
        if scalar:
            return to_output(delta, True), int(self.__last_target_idx[0].item())
        return delta, self.__last_target_idx.clone()


class PID1:
    """cbf/controllers.py:153-180: e = xref - x; de = (e - e_prev)/dt; ie += dt e; u = kp e + ki ie + kd de."""

    def __init__(self, kp=1.0, kd=0.0, ki=0.0):
        self.__kp = kp
        self.__kd = kd
        self.__ki = ki
        self.__e = 0
        self.__eprev = 0
        self.__de = 0
        self.__ie = 0
        self.__dt = 0.1

    def set_gains(self, kp, kd, ki):
        self.__kp = kp
        self.__kd = kd
        self.__ki = ki

    def set_dt(self, dt):
        self.__dt = dt

    def control(self, x, xref):
        self.__e = xref - x
        self.__de = (self.__e - self.__eprev) / self.__dt
        self.__ie = self.__ie + self.__dt * self.__e
        u = self.__kp * self.__e + self.__ki * self.__ie + self.__kd * self.__de
        self.__eprev = self.__e
        return u
