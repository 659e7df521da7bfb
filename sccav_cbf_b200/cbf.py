"""CBF-QP safety filters with the class API of the reference's cbf/cbf.py:

    DBM_CBF_2DS   acceleration-controlled dynamic bicycle model, u = (a, beta)   (cbf/cbf.py:112-220)
    KBM_VC_CBF2D  velocity-controlled kinematic bicycle model,   u = (v, omega)  (cbf/cbf.py:33-110)

``solve_cbf`` is ONE launch of the fused CUDA kernel K1+K2 (csrc/kernels.cuh filter_step_kernel:
barrier rows -> exact 2-variable QP -> control conversion) instead of a cvxopt interior-point run;
it returns the exact KKT optimum that ``cvxopt.solvers.cp`` approximates to ~1e-5.

Scalars (one scenario, Python floats in / a length-2 CPU tensor out, as the reference's 2x1 cvxopt
matrix) or batches ([4, N] state, [2, N] references in / [2, N] CUDA tensor out) -- see
``sccav_cbf_b200._batch``.  There is no CPU fallback.

``DUM_CBF_2DS`` (cbf/cbf.py:222-298) is provided with the 4-vector ``fc`` its code lists (the reference declares
it 5 x 1 and raises).  ``SADBM_CBF_2DS`` (cbf/cbf.py:300-437) is provided with its fixed ``dt`` (default 0.001) and,
for ``dt=None``, the reference's wall-clock step measured on the host; its debug prints are dropped.
"""
from __future__ import annotations

import time
from typing import Optional

import numpy as np
import torch

from . import _native as nv
from . import ops
from ._batch import as_state, as_vec, batch_size, cuda_device, is_scalar, state_scalar
from .obstacles import ObstacleList2D
from .utils import ZERO_TOL

EMPTY_MSG = ("Cannot solve CBF for an empty obstacle list. Update the obstacle list so that it is non-empty in order "
             "to move forward.")


def _as_uref(u_ref, N, dtype, device):
    """u_ref = [u0, u1] numbers, or [2, N] tensor / array, or a pair of [N] tensors -> [2, N]."""
    if isinstance(u_ref, torch.Tensor) and u_ref.dim() == 2:
        t = u_ref.to(device=device, dtype=dtype)
        if t.shape[1] == 1 and N > 1:
            t = t.expand(2, N)
        return t.contiguous(), False
    if isinstance(u_ref, np.ndarray) and u_ref.ndim == 2 and u_ref.shape[1] > 1:
        return torch.from_numpy(np.ascontiguousarray(u_ref, dtype=np.float64)).to(device=device, dtype=dtype), False
    c = [u_ref[0], u_ref[1]]
    scalar = all(is_scalar(v) for v in c)
    n = max([N] + [batch_size(v) for v in c])
    c = [np.asarray(v).reshape(-1)[0] if (is_scalar(v) and not isinstance(v, torch.Tensor)) else v for v in c]
    return torch.stack([as_vec(v, n, dtype, device) for v in c]).contiguous(), scalar


class _FilterBase:
    MODEL = nv.MODEL_DBM

    def __init__(self, alpha=1.0):
        self.obstacle_list2d = ObstacleList2D()
        self._alpha = alpha
        self._R = np.eye(2)                                          # cbf.py:52,134
        self.s = None
        self.last_info = None
        self._host_scalar = True          # one-scenario calls with Python numbers go through the host entry point

    def set_alpha(self, alpha=1.0):
        self._alpha = alpha

    def set_qp_cost_weight(self, R):
        """2x2 cost weight of (u - u_ref)^T R (u - u_ref) (cbf.py:154-157).  A constant 2x2, or a
        [2, 2, N] / [4, N] tensor of per-vehicle weights."""
        if isinstance(R, torch.Tensor) and R.dim() == 3:
            if tuple(R.shape[:2]) != (2, 2):
                raise ValueError("Expected a symmetrix matrix of size 2 as input. Please check the value of the matrix R you are using.")
            self._R = R.reshape(4, -1)
            return
        if isinstance(R, torch.Tensor) and R.dim() == 2 and R.shape[0] == 4 and R.shape[1] != 4:
            self._R = R
            return
        R = np.asarray(R.cpu() if isinstance(R, torch.Tensor) else R, dtype=np.float64)
        if R.shape != (2, 2):
            raise ValueError("Expected a symmetrix matrix of size 2 as input. Please check the value of the matrix R you are using.")
        self._R = R

    # ---- shared solve -------------------------------------------------------------------------------
    def _aug(self, N, state):
        """carried per-vehicle state of the model (SADBM only)"""
        return None

    def _params(self, **kw):
        R = self._R if isinstance(self._R, np.ndarray) else np.eye(2)
        return ops.make_params(model=self.MODEL, R=R.reshape(-1).tolist(), **kw)

    def _state(self):
        """([4, N] CUDA tensor, False) for a batch; (numpy [4], True) for one scenario given as plain numbers when the
        host entry point will take the call (no device tensor is made for it)."""
        if self._host_scalar and self.MODEL != nv.MODEL_SADBM:
            v = state_scalar(self.s)
            if v is not None:
                return v, True
        return as_state(self.s)

    def _solve_scalar_host(self, state, u_ref, params, return_solver):
        """One scenario given as Python numbers (the reference's own call): everything is marshalled on the host
        and handed to the library's host entry point in ONE call -- no torch device op per field."""
        descs, obst = self.obstacle_list2d.pack_host()
        M = obst.shape[0]
        st = torch.from_numpy(np.ascontiguousarray(state, dtype=np.float64).reshape(4, 1))
        ob = torch.from_numpy(obst.reshape(M, nv.NFIELD, 1))
        ur = torch.tensor([[float(np.asarray(u_ref[0]).reshape(-1)[0])], [float(np.asarray(u_ref[1]).reshape(-1)[0])]], dtype=torch.float64)
        params.alpha = float(self._alpha)
        aug = self._aug(1, st)
        u, mask, status, hmin = ops.filter_step(params, descs, st, ob, ur, aug=aug)
        self.last_info = {"status": status, "active_mask": mask, "h_min": hmin, "u_ref": ur}
        out = u[:, 0].clone()
        if not return_solver:
            return out
        info = {"status": int(status[0]), "active_mask": int(mask[0]) & 0xFFFFFFFF, "h_min": float(hmin[0]), "x": out}
        return info, out

    def _scalar_call(self, u_ref) -> bool:
        if isinstance(self._alpha, torch.Tensor) and self._alpha.dim() > 0:
            return False
        if isinstance(self._R, torch.Tensor):
            return False
        if getattr(self.obstacle_list2d, "count", None) is not None or not hasattr(self.obstacle_list2d, "pack_host"):
            return False
        try:
            if len(u_ref) != 2 or not all(is_scalar(v) for v in (u_ref[0], u_ref[1])):
                return False
        except TypeError:
            return False
        if isinstance(u_ref, torch.Tensor) and u_ref.dim() == 2:
            return False
        return self.obstacle_list2d.all_scalar()

    def _solve(self, state, scalar_state, u_ref, params, return_solver):
        if len(self.obstacle_list2d) < 1:
            raise ValueError(EMPTY_MSG)                               # cbf.py:77-80,177-180
        if isinstance(state, np.ndarray):
            if self._scalar_call(u_ref):
                return self._solve_scalar_host(state, u_ref, params, return_solver)
            state = torch.from_numpy(state.reshape(4, 1)).to(cuda_device())
        prepared = bool(getattr(self.obstacle_list2d, "use_prepared", False)) and self.MODEL != nv.MODEL_SADBM
        if prepared:
            slot_desc, obst, scalar_obs = self.obstacle_list2d.pack(state, prepared=True)
        else:
            slot_desc, obst, scalar_obs = self.obstacle_list2d.pack(state)
        N = obst.shape[2]
        if state.shape[1] != N:
            state = state.expand(4, N).contiguous()
        ur, scalar_u = _as_uref(u_ref, N, state.dtype, state.device)
        if ur.shape[1] != N:
            if N == 1:
                N = ur.shape[1]
                state = state.expand(4, N).contiguous()
                obst = obst.expand(obst.shape[0], nv.NFIELD, N).contiguous()
            else:
                raise ValueError("u_ref has %d columns but the batch has %d vehicles" % (ur.shape[1], N))
        alpha = None
        if isinstance(self._alpha, torch.Tensor) and self._alpha.dim() > 0:
            alpha = as_vec(self._alpha, N, state.dtype, state.device)
        else:
            params.alpha = float(self._alpha)
        Rv = None
        if isinstance(self._R, torch.Tensor):
            Rv = self._R.to(device=state.device, dtype=state.dtype).contiguous()
        u, mask, status, hmin = ops.filter_step(params, slot_desc, state, obst, ur, alpha=alpha, R=Rv,
                                                count=getattr(self.obstacle_list2d, "count", None), aug=self._aug(N, state))
        scalar = scalar_state and scalar_obs and scalar_u and N == 1
        info = {"status": status, "active_mask": mask, "h_min": hmin, "u_ref": ur}
        self.last_info = info
        out = u[:, 0].cpu() if scalar else u
        if scalar:
            info = {"status": int(status[0].item()), "active_mask": int(mask[0].item()) & 0xFFFFFFFF, "h_min": float(hmin[0].item()),
                    "x": out}
        return (info, out) if return_solver else out


class DBM_CBF_2DS(_FilterBase):
    """Control barrier function filter for the acceleration-controlled dynamic bicycle model with the
    small-side-slip approximation (cbf/cbf.py:112-220).  State s = [x, y, theta, v], control
    u = [a, delta] in / out (delta <-> beta conversions of cbf.py:175,216 happen on the GPU)."""
    MODEL = nv.MODEL_DBM

    def __init__(self, alpha=1.0):
        super().__init__(alpha)
        self._lr = None
        self._lf = None

    def update_state(self, s, s_obs_dict: Optional[dict] = None, buffer: Optional[float] = None, **kwargs):
        """cbf.py:137-145: stores the ego state and pushes it into every obstacle of the list."""
        self.s = s
        self.s_obs_dict = s_obs_dict
        self.obstacle_list2d.update_state(s=s, s_obs_dict=s_obs_dict, buffer=buffer)

    def set_model_params(self, lr, lf):
        self._lr = lr
        self._lf = lf

    def solve_cbf(self, u_ref, return_solver=False):
        """u = argmin (u - u_ref)^T R (u - u_ref)  s.t.  Lf h_k + Lg h_k u + alpha h_k + dh_k/dt >= 0 for every
        obstacle k (cbf.py:166-220).  ``u_ref`` = [a_ref, delta_ref]; returns u = [a, delta], or
        (info, u) with ``return_solver`` (info: status, active_mask, h_min)."""
        if len(self.obstacle_list2d) < 1:
            raise ValueError(EMPTY_MSG)
        if self._lr is None or self._lf is None:
            raise AttributeError("set_model_params(lr, lf) has not been called")
        if self.s is None:
            raise AttributeError("update_state(s) has not been called")
        state, scalar = self._state()
        params = self._params(lr=float(self._lr), lf=float(self._lf), L=float(self._lr) + float(self._lf))
        return self._solve(state, scalar, u_ref, params, return_solver)


class KBM_VC_CBF2D(_FilterBase):
    """Velocity-controlled kinematic bicycle model filter (cbf/cbf.py:33-110; functional twin ``CBF()``,
    stanley_controller_ellipse.py:214-238).  State (p, theta); control u = [v, delta] in / out.
    D7 of the survey: the ego position / heading ARE forwarded to the obstacles, and the cost uses R."""
    MODEL = nv.MODEL_KBM

    def __init__(self, alpha=1.0):
        super().__init__(alpha)
        self._L = None
        self._p = None
        self._theta = None

    def update_state(self, p, theta):
        """p: Point2 / Vector2 / (x, y); theta: heading (cbf.py:54-56)."""
        self._p = p
        self._theta = theta
        x, y = (p.x, p.y) if hasattr(p, "x") else (p[0], p[1])
        self.s = [x, y, theta, 0.0]
        self.obstacle_list2d.update_state(s=self.s)

    def set_model_params(self, L):
        self._L = L

    def solve_cbf(self, u_ref):
        """``u_ref`` = [v_ref, delta_ref] -> (info, u) with u = [v, delta] (cbf.py:67-110; always returns
        the pair, like the reference)."""
        if len(self.obstacle_list2d) < 1:
            raise ValueError(EMPTY_MSG)
        if self._L is None:
            raise AttributeError("set_model_params(L) has not been called")
        if self.s is None:
            raise AttributeError("update_state(p, theta) has not been called")
        state, scalar = self._state()
        params = self._params(L=float(self._L))
        return self._solve(state, scalar, u_ref, params, True)


class DUM_CBF_2DS(DBM_CBF_2DS):
    """Dynamic unicycle model filter (cbf/cbf.py:222-298): state s = [x, y, theta, v], control u = [a, omega]
    in and out -- g_c selects (v_dot, theta_dot), so Lg h = [h_v, h_theta]; no delta <-> beta conversion.
    Ellipse / lane barriers have h_v = h_theta = 0 and cannot be influenced by this model's controls;
    it is meant for the collision cone."""
    MODEL = nv.MODEL_DUM

    def solve_cbf(self, u_ref, return_solver=False):
        if len(self.obstacle_list2d) < 1:
            raise ValueError(EMPTY_MSG)
        if self.s is None:
            raise AttributeError("update_state(s) has not been called")
        state, scalar = self._state()
        return self._solve(state, scalar, u_ref, self._params(), return_solver)


class SADBM_CBF_2DS(DBM_CBF_2DS):
    """State-augmented, steer-rate controlled dynamic bicycle model (cbf/cbf.py:300-437): beta joins the state,
    its time derivative is the second control, so the small-angle approximation of DBM_CBF_2DS is gone and beta
    stays continuous.  ``solve_cbf([a_ref, delta_ref])`` converts delta_ref to beta_ref, differentiates it
    against the previous call (``dt``), solves for (a, d(beta)/dt), integrates beta and returns [a, delta].
    The object is stateful exactly like the reference's: ``_beta`` and ``beta_ref_last`` ([N] tensors here)
    carry over between calls; every CollisionCone2D of the list sees the vehicle's beta (cbf.py:424-426).

    ``dt`` is the class's fixed step (default 0.001, cbf.py:323).  ``dt=None`` selects the reference's
    wall-clock mode (cbf.py:326-328,363-365): the time since the previous call, floored at ZERO_TOL, measured on
    the host and handed to the kernel -- by nature not reproducible.  The reference's prints are dropped."""
    MODEL = nv.MODEL_SADBM

    def __init__(self, alpha=1.0, dt=0.001):
        super().__init__(alpha)
        self._DT_MODE_AUTO = 1 if dt is None else 0
        self._dt = 1e-6 if dt is None else dt
        self.t_last = time.time()
        self._beta = None              # [N] tensor after the first solve (0.0 before, cbf.py:335-336)
        self.beta_ref_last = None
        self._augbuf = None

    def _aug(self, N, state):
        if self._augbuf is None or self._augbuf.shape[1] != N or self._augbuf.device != state.device or self._augbuf.dtype != state.dtype:
            self._augbuf = torch.zeros((2, N), dtype=state.dtype, device=state.device)
        self._beta, self.beta_ref_last = self._augbuf[0], self._augbuf[1]
        return self._augbuf

    def solve_cbf(self, u_ref, return_solver=False):
        if len(self.obstacle_list2d) < 1:
            raise ValueError(EMPTY_MSG)
        if self._lr is None or self._lf is None:
            raise AttributeError("set_model_params(lr, lf) has not been called")
        if self.s is None:
            raise AttributeError("update_state(s) has not been called")
        t_current = time.time()
        if self._DT_MODE_AUTO:
            self._dt = max(t_current - self.t_last, ZERO_TOL)                      # cbf.py:363-365
        state, scalar = self._state()
        params = self._params(lr=float(self._lr), lf=float(self._lf), L=float(self._lr) + float(self._lf),
                              sadbm_dt=float(self._dt))
        out = self._solve(state, scalar, u_ref, params, return_solver)
        self.t_last = time.time()                                                  # cbf.py:433
        return out
