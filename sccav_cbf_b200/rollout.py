"""Closed-loop rollout front end: the batched counterpart of the reference's per-tick loops
(test_scripts/stanley_controller_ellipse.py:630-830, radial_dynamic_obstacles.py:427-507).

``ClosedLoopRollout`` owns the device buffers of one scenario shard (one process per GPU; shards
are independent, there is no collective on the path) and exposes

* ``run()``          -- inputs already resident in HBM (what ``bench.py`` reports as ``value``);
* ``run_from_host()``-- the end-to-end call: pinned host inputs -> H2D -> kernel -> D2H of the
                        per-vehicle results (what ``bench.py`` reports as ``e2e``).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import ops
from .scenarios import ScenarioBatch


class ClosedLoopRollout:
    def __init__(self, batch: ScenarioBatch, dtype: torch.dtype = torch.float64, device: Optional[torch.device] = None,
                 pin: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError("ClosedLoopRollout needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.dtype = dtype
        self.batch = batch
        self.params = ops.make_params(**batch.params)
        self.slot_desc = list(batch.slot_desc)
        self.T = batch.T

        def host(a):
            if a is None:
                return None
            t = torch.from_numpy(np.ascontiguousarray(a)).to(dtype)
            return t.pin_memory() if pin else t

        # host copies (pinned): the source of the e2e path
        self.h_state = host(batch.state)
        self.h_obst = host(batch.obst)
        self.h_alpha = host(batch.alpha)
        self.h_R = host(batch.R)
        self.h_tspeed = host(batch.target_speed)
        self.course = None
        self.h_course = None
        if batch.course is not None:
            self.h_course = tuple(host(c) for c in batch.course)
            self.course = tuple(c.to(self.device) for c in self.h_course)
        # device-resident inputs
        self.d_state = self.h_state.to(self.device)
        self.d_obst0 = None if self.h_obst is None else self.h_obst.to(self.device)
        self.d_obst = None if self.d_obst0 is None else self.d_obst0.clone()
        self.d_alpha = None if self.h_alpha is None else self.h_alpha.to(self.device)
        self.d_R = None if self.h_R is None else self.h_R.to(self.device)
        self.d_tspeed = None if self.h_tspeed is None else self.h_tspeed.to(self.device)
        self.out: Dict[str, torch.Tensor] = {}
        self._host_out: Dict[str, torch.Tensor] = {}
        self._h_obst0 = None

    @property
    def N(self) -> int:
        return self.h_state.shape[1]

    @property
    def M(self) -> int:
        return len(self.slot_desc)

    def d2h_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self._host_out.values())

    def reset(self) -> None:
        """Restore the moving-obstacle buffer (rollouts with seekers update it in place)."""
        if self.d_obst is not None and self.params.seeker:
            self.d_obst.copy_(self.d_obst0)

    def run(self, T: Optional[int] = None, record_stride: int = 0) -> Dict[str, torch.Tensor]:
        self.reset()
        return self.launch(T, record_stride)

    def launch(self, T: Optional[int] = None, record_stride: int = 0) -> Dict[str, torch.Tensor]:
        """The rollout launch alone (callers that time it call reset() themselves, outside the timed region)."""
        return ops.rollout(self.params, self.slot_desc, self.d_state, self.d_obst, self.course,
                           self.T if T is None else T, alpha=self.d_alpha, R=self.d_R, target_speed=self.d_tspeed,
                           record_stride=record_stride, out=self.out)

    def run_from_host(self, T: Optional[int] = None) -> Dict[str, torch.Tensor]:
        """The end-to-end call: ONE C-ABI call with HOST pointers (sccav_rollout_host_*): the inputs are
        copied from pinned host memory, the rollout kernel runs, the per-vehicle results are copied
        back, and the stream is synchronised inside the call.  Returns pinned host tensors."""
        if self.h_obst is not None and self.params.seeker:
            # the host entry point writes the moved obstacles back into the buffer it was given: start every call
            # from the pristine scenario, not from the previous call's final obstacle positions
            if self._h_obst0 is None:
                self._h_obst0 = self.h_obst.clone()
            else:
                self.h_obst.copy_(self._h_obst0)
        with torch.cuda.device(self.device):
            res = ops.rollout(self.params, self.slot_desc, self.h_state, self.h_obst, self.h_course,
                              self.T if T is None else T, alpha=self.h_alpha, R=self.h_R, target_speed=self.h_tspeed,
                              record_stride=0, out=self._host_out)
        self._host_out = res
        return res

    def h2d_bytes(self) -> int:
        n = self.h_state.numel() * self.h_state.element_size()
        for t in (self.h_obst, self.h_alpha, self.h_R, self.h_tspeed) + (self.h_course or ()):
            if t is not None:
                n += t.numel() * t.element_size()
        return n
