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

import ctypes as C

from . import _native as nv
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

    def pipeline(self, depth: int = 2, resident_obstacles: bool = False) -> RolloutPipeline:
        """A pipelined host-API handle for this batch shape (course resident; obstacles resident on request)."""
        return RolloutPipeline(self.params, self.slot_desc, self.N, self.T, self.h_course, self.dtype,
                               obst_resident=self.h_obst if (resident_obstacles and not self.params.seeker) else None,
                               depth=depth, device=self.device)

    def run_pipelined(self, pipe: RolloutPipeline, steps: int):
        """`steps` end-to-end rollouts of this batch through the pipeline: every submission uploads the batch's inputs
        from pinned host memory and downloads its results.  Returns the results of the last one."""
        tickets = []
        res = None
        for i in range(steps):
            tickets.append(pipe.submit(self.h_state, None if pipe.resident_obstacles else self.h_obst, alpha=self.h_alpha, R=self.h_R,
                                       target_speed=self.h_tspeed))
            if len(tickets) > pipe.depth - 1:
                res = pipe.wait(tickets.pop(0))
        while tickets:
            res = pipe.wait(tickets.pop(0))
        return res

    def h2d_bytes(self) -> int:
        n = self.h_state.numel() * self.h_state.element_size()
        for t in (self.h_obst, self.h_alpha, self.h_R, self.h_tspeed) + (self.h_course or ()):
            if t is not None:
                n += t.numel() * t.element_size()
        return n


class RolloutPipeline:
    """Handle of the pipelined host API (sccav_pipeline_*, include/sccav_cbf.h): the course -- and static obstacles --
    are uploaded once and stay resident; ``submit`` enqueues H2D of a batch's inputs, the rollout and D2H of its
    per-vehicle summaries on three streams and returns at once; ``wait`` blocks until a submission's results are in
    (pinned) host memory.  With ``depth`` >= 2 the copies of neighbouring submissions hide behind the kernels."""

    SUMMARY = ("steps", "target_idx", "n_active", "n_infeasible", "h_min", "beta_min", "beta_max", "beta_int", "n_evals")

    def __init__(self, params, slot_desc, N: int, T: int, course=None, dtype: torch.dtype = torch.float64,
                 obst_resident: Optional[torch.Tensor] = None, depth: int = 2, device: Optional[torch.device] = None):
        nv.require_cuda()
        self.L = nv.lib()
        self.sfx = "f64" if dtype == torch.float64 else "f32"
        self.dtype, self.N, self.T, self.depth = dtype, int(N), int(T), int(depth)
        self.sd = ops.slot_bytes(slot_desc)
        self.M = len(self.sd)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        import copy
        self.params = copy.copy(params)
        self.params.record_stride = 0
        cx = cy = cyaw = None
        P = 0
        if course is not None:
            cx, cy, cyaw = (c.to(dtype).contiguous() for c in course)
            P = cx.shape[0]
        ob = None if obst_resident is None else obst_resident.to(dtype).contiguous()
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            nv.check(getattr(self.L, "sccav_pipeline_create_" + self.sfx)(
                C.byref(self.params), self.sd, self.M, self.N, self.T, ops._ptr(cx), ops._ptr(cy), ops._ptr(cyaw), P, ops._ptr(ob),
                self.depth, C.byref(h)))
        self.handle = h
        self.resident_obstacles = ob is not None
        # one set of pinned result buffers per in-flight submission
        idt = torch.int32
        self.results = [dict(state=torch.empty((4, self.N), dtype=dtype).pin_memory(),
                             **{k: torch.empty((self.N,), dtype=(idt if k in ("steps", "target_idx", "n_active", "n_infeasible", "n_evals") else dtype)).pin_memory()
                                for k in self.SUMMARY}) for _ in range(self.depth)]
        self._keep = [None] * self.depth

    def submit(self, state: torch.Tensor, obst: Optional[torch.Tensor] = None, alpha=None, R=None, target_speed=None, count=None) -> int:
        """HOST tensors (pinned for true overlap).  Returns the ticket; its results live in ``self.results[ticket % depth]``."""
        if state.is_cuda or (obst is not None and obst.is_cuda):
            raise ValueError("the pipelined API takes host tensors")
        slot = None
        pv = nv.PerVehicle()
        keep = [state, obst, alpha, R, target_speed, count]
        for name, t in (("alpha", alpha), ("R", R), ("target_speed", target_speed)):
            if t is not None:
                setattr(pv, name, t.data_ptr())
        if count is not None:
            pv.count = count.data_ptr()
        ro = nv.RolloutOut()
        ticket = C.c_int64(-1)
        # the slot this submission will use is (next % depth): results buffer of that slot
        nxt = getattr(self, "_next", 0)
        slot = nxt % self.depth
        res = self.results[slot]
        ro.state = res["state"].data_ptr()
        for k in self.SUMMARY:
            setattr(ro, k, res[k].data_ptr())
        with torch.cuda.device(self.device):
            nv.check(getattr(self.L, "sccav_pipeline_submit_" + self.sfx)(self.handle, ops._ptr(state), ops._ptr(obst), C.byref(pv), C.byref(ro),
                                                                         C.byref(ticket)))
        self._keep[slot] = keep
        self._next = nxt + 1
        return int(ticket.value)

    def wait(self, ticket: int) -> Dict[str, torch.Tensor]:
        with torch.cuda.device(self.device):
            nv.check(getattr(self.L, "sccav_pipeline_wait_" + self.sfx)(self.handle, int(ticket)))
        return self.results[ticket % self.depth]

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            with torch.cuda.device(self.device):
                getattr(self.L, "sccav_pipeline_destroy_" + self.sfx)(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
