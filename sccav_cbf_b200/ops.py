"""Tensor-level operators over the C-ABI (libsccav_cbf.so): thin, no arithmetic in Python.

All arrays are structure-of-arrays torch tensors with the vehicle index last (fastest):

    state [4, N]   obst [M, 8, N]   u_ref / u [2, N]   A [2, M, N]   b [M, N]

CUDA tensors go through the device-pointer entry points on the current torch stream; CPU tensors
go through the ``*_host_*`` entry points (H2D + kernel + D2H inside the call).  dtype float64
(default, parity-checked) or float32 (reported variant).  There is no CPU compute path.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import torch

from . import _native as nv
from ._native import Params

_SFX = {torch.float64: "f64", torch.float32: "f32"}


def make_params(**kw) -> Params:
    """sccav_params with the reference defaults (stanley_controller_ellipse.py:52-58,590),
    overridden by keyword.  ``R`` may be a 2x2 nested sequence / tensor or 4 numbers row-major."""
    p = nv.default_params()
    for k, v in kw.items():
        if k == "R":
            flat = torch.as_tensor(v, dtype=torch.float64).reshape(-1).tolist()
            if len(flat) != 4:
                raise ValueError("Expected a symmetrix matrix of size 2 as input.")     # cbf.py:156-157
            for i in range(4):
                p.R[i] = flat[i]
        elif hasattr(p, k):
            setattr(p, k, v)
        else:
            raise TypeError("unknown parameter %r" % k)
    return p


def slot_bytes(slot_desc: Sequence[int]) -> bytes:
    b = bytes(int(d) & 0xFF for d in slot_desc)
    if len(b) > nv.MAX_ROWS:
        raise ValueError("at most %d obstacle slots are supported" % nv.MAX_ROWS)
    return b


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _chk(t: torch.Tensor, shape, dtype, device, name: str) -> torch.Tensor:
    if t.dtype != dtype:
        raise TypeError("%s: expected dtype %s, got %s" % (name, dtype, t.dtype))
    if t.device != device:
        raise ValueError("%s: expected device %s, got %s" % (name, device, t.device))
    if tuple(t.shape) != tuple(shape):
        raise ValueError("%s: expected shape %s, got %s" % (name, tuple(shape), tuple(t.shape)))
    return t if t.is_contiguous() else t.contiguous()


def _stream(device: torch.device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _pv(N, dtype, device, alpha=None, R=None, target_speed=None, count=None, aug=None):
    pv = nv.PerVehicle()
    keep = []
    if aug is not None:
        # checked BEFORE _chk (which would hand back a contiguous copy: the kernel would advance the copy's beta /
        # beta_ref_last and the caller's SADBM state would never move)
        if not aug.is_contiguous():
            raise ValueError("aug must be contiguous (it is updated in place)")
        aug = _chk(aug, (2, N), dtype, device, "aug")
        keep.append(aug); pv.aug = aug.data_ptr()
    if count is not None:
        count = _chk(count, (N,), torch.int32, device, "count"); keep.append(count); pv.count = count.data_ptr()
    if alpha is not None:
        alpha = _chk(alpha, (N,), dtype, device, "alpha"); keep.append(alpha); pv.alpha = alpha.data_ptr()
    if R is not None:
        R = _chk(R, (4, N), dtype, device, "R"); keep.append(R); pv.R = R.data_ptr()
    if target_speed is not None:
        target_speed = _chk(target_speed, (N,), dtype, device, "target_speed"); keep.append(target_speed)
        pv.target_speed = target_speed.data_ptr()
    return pv, keep


def _need_cuda_lib():
    nv.require_cuda()
    return nv.lib()


def barrier_rows(params: Params, slot_desc, state: torch.Tensor, obst: torch.Tensor, alpha=None, count=None):
    """K1: rows (A [2,M,N], b [M,N]) and barrier values h [M,N] -- ObstacleList2D.f/dx/.. +
    the assembly of cbf/cbf.py:194-207."""
    L = _need_cuda_lib()
    sd = slot_bytes(slot_desc)
    M, N = len(sd), state.shape[-1]
    dt, dev = state.dtype, state.device
    if not state.is_cuda:
        raise ValueError("barrier_rows takes CUDA tensors")
    state = _chk(state, (4, N), dt, dev, "state")
    obst = _chk(obst, (M, nv.NFIELD, N), dt, dev, "obst")
    pv, keep = _pv(N, dt, dev, alpha=alpha, count=count)
    A = torch.empty((2, M, N), dtype=dt, device=dev)
    b = torch.empty((M, N), dtype=dt, device=dev)
    h = torch.empty((M, N), dtype=dt, device=dev)
    with torch.cuda.device(dev):
        nv.check(getattr(L, "sccav_barrier_rows_" + _SFX[dt])(
            C.byref(params), sd, M, N, _ptr(state), _ptr(obst), C.byref(pv), _ptr(A), _ptr(b), _ptr(h), _stream(dev)))
    return A, b, h


def barrier_partials(slot_desc, state: torch.Tensor, obst: torch.Tensor) -> torch.Tensor:
    """K0: barrier values and partials [M, 6, N] = (h, h_x, h_y, h_theta, h_v, h_t) per slot --
    the obstacle getters f/dx/dy/dtheta/dv/dt of cbf/obstacles.py."""
    L = _need_cuda_lib()
    sd = slot_bytes(slot_desc)
    M, N = len(sd), state.shape[-1]
    dt, dev = state.dtype, state.device
    if not state.is_cuda:
        raise ValueError("barrier_partials takes CUDA tensors")
    state = _chk(state, (4, N), dt, dev, "state")
    obst = _chk(obst, (M, nv.NFIELD, N), dt, dev, "obst")
    out = torch.empty((M, 6, N), dtype=dt, device=dev)
    with torch.cuda.device(dev):
        nv.check(getattr(L, "sccav_barrier_partials_" + _SFX[dt])(sd, M, N, _ptr(state), _ptr(obst), _ptr(out), _stream(dev)))
    return out


def stanley_control(params: Params, state: torch.Tensor, course, target_idx: torch.Tensor,
                    front: Optional[torch.Tensor] = None, want_error: bool = False):
    """KS: one Stanley steering call (LateralStanley.control / stanley_control).  ``target_idx``
    (int32 [N]) is updated IN PLACE to the new target index.  Returns delta [N] (and the
    front-axle error [N] with ``want_error``)."""
    L = _need_cuda_lib()
    N = state.shape[-1]
    dt, dev = state.dtype, state.device
    if not state.is_cuda:
        raise ValueError("stanley_control takes CUDA tensors")
    state = _chk(state, (4, N), dt, dev, "state")
    cx, cy, cyaw = course
    P = cx.shape[0]
    cx = _chk(cx, (P,), dt, dev, "course_x"); cy = _chk(cy, (P,), dt, dev, "course_y"); cyaw = _chk(cyaw, (P,), dt, dev, "course_yaw")
    if target_idx.dtype != torch.int32 or tuple(target_idx.shape) != (N,) or target_idx.device != dev or not target_idx.is_contiguous():
        raise ValueError("target_idx must be a contiguous int32 [N] tensor on the state's device")
    if front is not None:
        front = _chk(front, (2, N), dt, dev, "front")
    delta = torch.empty((N,), dtype=dt, device=dev)
    err = torch.empty((N,), dtype=dt, device=dev) if want_error else None
    with torch.cuda.device(dev):
        nv.check(getattr(L, "sccav_stanley_control_" + _SFX[dt])(
            C.byref(params), N, _ptr(state), _ptr(front), _ptr(cx), _ptr(cy), _ptr(cyaw), P, _ptr(target_idx), _ptr(delta),
            _ptr(err), _stream(dev)))
    return (delta, err) if want_error else delta


def qp2_solve(params: Params, A: torch.Tensor, b: torch.Tensor, r: torch.Tensor, R=None, warp_per_problem=False, count=None):
    """K2: exact optimum of min (u-r)^T R (u-r) s.t. A u >= b (what cvxopt.solvers.cp approximates
    at cbf/cbf.py:213).  Returns u [2,N], active mask int32 [N] (bit k = row k), status uint8 [N]."""
    L = _need_cuda_lib()
    M, N = b.shape
    dt, dev = b.dtype, b.device
    if not b.is_cuda:
        raise ValueError("qp2_solve takes CUDA tensors")
    A = _chk(A, (2, M, N), dt, dev, "A")
    b = _chk(b, (M, N), dt, dev, "b")
    r = _chk(r, (2, N), dt, dev, "r")
    pv, keep = _pv(N, dt, dev, R=R, count=count)
    u = torch.empty((2, N), dtype=dt, device=dev)
    mask = torch.empty((N,), dtype=torch.int32, device=dev)
    status = torch.empty((N,), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        nv.check(getattr(L, "sccav_qp2_solve_" + _SFX[dt])(
            C.byref(params), M, N, _ptr(A), _ptr(b), _ptr(r), C.byref(pv), _ptr(u), _ptr(mask), _ptr(status),
            1 if warp_per_problem else 0, _stream(dev)))
    return u, mask, status


def filter_step(params: Params, slot_desc, state: torch.Tensor, obst: torch.Tensor, u_ref: torch.Tensor,
                alpha=None, R=None, count=None, aug=None):
    """K1+K2 fused: one batched ``solve_cbf(u_ref)`` (cbf/cbf.py:166-220 / :67-110).
    CUDA tensors -> device entry point on the current stream; CPU tensors -> host entry point.
    ``aug`` [2,N] (model SADBM only, cbf/cbf.py:300-437): beta and the last beta_ref, updated IN PLACE.
    Returns u [2,N] = (a|v, delta), active mask int32 [N], status uint8 [N], h_min [N]."""
    L = nv.lib()
    nv.require_cuda()
    sd = slot_bytes(slot_desc)
    M, N = len(sd), state.shape[-1]
    dt, dev = state.dtype, state.device
    state = _chk(state, (4, N), dt, dev, "state")
    obst = _chk(obst, (M, nv.NFIELD, N), dt, dev, "obst")
    u_ref = _chk(u_ref, (2, N), dt, dev, "u_ref")
    pv, keep = _pv(N, dt, dev, alpha=alpha, R=R, count=count, aug=aug)
    u = torch.empty((2, N), dtype=dt, device=dev)
    mask = torch.empty((N,), dtype=torch.int32, device=dev)
    status = torch.empty((N,), dtype=torch.uint8, device=dev)
    hmin = torch.empty((N,), dtype=dt, device=dev)
    if state.is_cuda:
        with torch.cuda.device(dev):
            nv.check(getattr(L, "sccav_filter_step_" + _SFX[dt])(
                C.byref(params), sd, M, N, _ptr(state), _ptr(obst), _ptr(u_ref), C.byref(pv), _ptr(u), _ptr(mask),
                _ptr(status), _ptr(hmin), _stream(dev)))
    else:
        cur = torch.device("cuda", torch.cuda.current_device())
        nv.check(getattr(L, "sccav_filter_step_host_" + _SFX[dt])(
            C.byref(params), sd, M, N, _ptr(state), _ptr(obst), _ptr(u_ref), C.byref(pv), _ptr(u), _ptr(mask),
            _ptr(status), _ptr(hmin), _stream(cur)))
    return u, mask, status, hmin


ROLLOUT_SUMMARY = ("steps", "target_idx", "n_active", "n_infeasible", "h_min", "beta_min", "beta_max", "beta_int",
                   "n_evals")


def rollout(params: Params, slot_desc, state: torch.Tensor, obst: Optional[torch.Tensor], course, T: int,
            alpha=None, R=None, target_speed=None, count=None, record_stride: int = 0, summary: bool = True,
            out: Optional[Dict[str, torch.Tensor]] = None, course_np: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """K3: persistent closed-loop rollout of T steps (stanley_controller_ellipse.py:630-830 /
    radial_dynamic_obstacles.py:427-507).

    ``course`` = (cx, cy, cyaw) tensors [P] (ignored for NOMINAL_CONST; may be None then), or [C, P_max] tensors with
    ``course_np`` int32 [C] for C roads in ONE launch (vehicles [c N/C, (c+1) N/C) drive road c) -- bit-identical to C
    launches with one road each.
    ``obst`` is updated in place when ``params.seeker``.  Returns a dict: ``state`` [4,N] final,
    the summary arrays of ROLLOUT_SUMMARY, and with ``record_stride`` > 0 ``traj`` [T_rec,7,N]
    (NaN-initialised), ``traj_idx`` / ``traj_mask`` [T_rec,N] (-1 / 0 initialised).
    ``out`` may carry preallocated tensors to reuse (keys as returned).
    """
    L = nv.lib()
    nv.require_cuda()
    sd = slot_bytes(slot_desc)
    M, N = len(sd), state.shape[-1]
    dt, dev = state.dtype, state.device
    state = _chk(state, (4, N), dt, dev, "state")
    if M > 0:
        if obst is None:
            raise ValueError("obst is required when there are obstacle slots")
        if not obst.is_contiguous():
            raise ValueError("obst must be contiguous (it is updated in place for moving obstacles)")
        obst = _chk(obst, (M, nv.NFIELD, N), dt, dev, "obst")
    roads = 0
    if course is not None:
        cx, cy, cyaw = course[:3]
        if cx.dim() == 2:
            # several roads [C, P_max] (the output of spline_course): ONE launch, vehicles grouped by road
            if not state.is_cuda:
                raise ValueError("several roads per launch take CUDA tensors")
            roads, P = int(cx.shape[0]), int(cx.shape[1])
            if course_np is None:
                raise ValueError("course_np (int32 [C]: points of every road) is required with [C, P_max] courses")
            cx = _chk(cx, (roads, P), dt, dev, "course_x"); cy = _chk(cy, (roads, P), dt, dev, "course_y")
            cyaw = _chk(cyaw, (roads, P), dt, dev, "course_yaw")
            course_np = _chk(course_np, (roads,), torch.int32, dev, "course_np")
            if N % roads:
                raise ValueError("%d vehicles do not split evenly over %d roads" % (N, roads))
        else:
            P = cx.shape[0]
            cx = _chk(cx, (P,), dt, dev, "course_x"); cy = _chk(cy, (P,), dt, dev, "course_y")
            cyaw = _chk(cyaw, (P,), dt, dev, "course_yaw")
    else:
        cx = cy = cyaw = None
        P = 0
    pv, keep = _pv(N, dt, dev, alpha=alpha, R=R, target_speed=target_speed, count=count)
    import copy
    prm = copy.copy(params)
    prm.record_stride = int(record_stride)
    res: Dict[str, torch.Tensor] = {} if out is None else out

    def buf(name, shape, dtype, fill=None):
        t = res.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype or t.device != dev:
            # host path: results land in pinned memory (plain D2H DMA, no staging copy)
            t = torch.empty(shape, dtype=dtype, pin_memory=True) if dev.type == "cpu" else torch.empty(shape, dtype=dtype, device=dev)
            res[name] = t
        if fill is not None:
            t.fill_(fill)
        return t

    ro = nv.RolloutOut()
    ro.state = buf("state", (4, N), dt).data_ptr()
    if summary:
        ro.steps = buf("steps", (N,), torch.int32).data_ptr()
        ro.target_idx = buf("target_idx", (N,), torch.int32).data_ptr()
        ro.n_active = buf("n_active", (N,), torch.int32).data_ptr()
        ro.n_infeasible = buf("n_infeasible", (N,), torch.int32).data_ptr()
        ro.h_min = buf("h_min", (N,), dt).data_ptr()
        ro.beta_min = buf("beta_min", (N,), dt).data_ptr()
        ro.beta_max = buf("beta_max", (N,), dt).data_ptr()
        ro.beta_int = buf("beta_int", (N,), dt).data_ptr()
        ro.n_evals = buf("n_evals", (N,), torch.int32).data_ptr()
    if record_stride > 0:
        trec = (T + record_stride - 1) // record_stride
        ro.traj = buf("traj", (trec, nv.TRAJ_FIELDS, N), dt, float("nan")).data_ptr()
        ro.traj_idx = buf("traj_idx", (trec, N), torch.int32, -1).data_ptr()
        ro.traj_mask = buf("traj_mask", (trec, N), torch.int32, 0).data_ptr()
    args = (C.byref(prm), sd, M, N, int(T), _ptr(state), _ptr(obst), _ptr(cx), _ptr(cy), _ptr(cyaw), P,
            C.byref(pv), C.byref(ro))
    if roads:
        with torch.cuda.device(dev):
            nv.check(getattr(L, "sccav_rollout_roads_" + _SFX[dt])(
                C.byref(prm), sd, M, N, int(T), _ptr(state), _ptr(obst), _ptr(cx), _ptr(cy), _ptr(cyaw), P, roads,
                _ptr(course_np), C.byref(pv), C.byref(ro), _stream(dev)))
    elif state.is_cuda:
        with torch.cuda.device(dev):
            nv.check(getattr(L, "sccav_rollout_" + _SFX[dt])(*args, _stream(dev)))
    else:
        cur = torch.device("cuda", torch.cuda.current_device())
        nv.check(getattr(L, "sccav_rollout_host_" + _SFX[dt])(*args, _stream(cur)))
    return res


def measure_fma_peak(dtype=torch.float64) -> float:
    """Achieved TFLOP/s of an unrolled FMA-chain kernel (the CUDA-core roofline denominator)."""
    L = _need_cuda_lib()
    v = C.c_double(0.0)
    nv.check(L.sccav_measure_fma_peak(64 if dtype == torch.float64 else 32, C.byref(v)))
    return v.value


def launch_count() -> int:
    return int(nv.lib().sccav_launch_count())


def prepare_obstacles(slot_desc, obst: torch.Tensor, out: Optional[torch.Tensor] = None):
    """KP: obstacle ingest for repeated solves -- every ELLIPSE slot of ``obst`` [M,8,N] becomes an
    ELLIPSE_PREP slot (the batched counterpart of constructing the Ellipse2D objects once,
    cbf/obstacles.py:146-165, and calling solve_cbf every tick).  Returns (new slot_desc list,
    prepared obst); ``out`` may be ``obst`` itself (in place)."""
    L = nv.lib()
    nv.require_cuda()
    sd = slot_bytes(slot_desc)
    M, N = len(sd), obst.shape[-1]
    dt, dev = obst.dtype, obst.device
    if dev.type != "cuda":
        raise ValueError("prepare_obstacles works on device tensors")
    obst = _chk(obst, (M, nv.NFIELD, N), dt, dev, "obst")
    if not obst.is_contiguous():
        raise ValueError("obst must be contiguous")
    if out is None:
        out = torch.empty_like(obst)
    out = _chk(out, (M, nv.NFIELD, N), dt, dev, "out")
    sd_out = (C.c_uint8 * max(M, 1))()
    with torch.cuda.device(dev):
        nv.check(getattr(L, "sccav_prepare_obstacles_" + _SFX[dt])(sd, M, N, _ptr(obst), _ptr(out), C.addressof(sd_out), _stream(dev)))
    return [int(sd_out[m]) for m in range(M)], out


def ingest_boxes(obs_type: int, box_id: torch.Tensor, box: torch.Tensor, slot_id: torch.Tensor, obst: torch.Tensor,
                 count: torch.Tensor, buffer: float = 0.5, mode: int = nv.INGEST_UPDATE,
                 dropped: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """KB: batched ``ObstacleList2D.update_by_bounding_box`` (cbf/obstacles.py:833-858) -- N vehicles,
    each with its own list of up to M obstacles held in ``slot_id`` [M,N] int32 (-1 = empty) / ``obst``
    [M,8,N] / ``count`` [N] int32, fed with up to K boxes ``box_id`` [K,N] int32 (< 0 = none) / ``box`` [K,6,N]
    (extent.x, extent.y, location.x, location.y, yaw, velocity).  All three list tensors are updated in
    place; ``count`` is what the filter kernels take as their per-vehicle ``count``."""
    L = nv.lib()
    nv.require_cuda()
    dt, dev = obst.dtype, obst.device
    if dev.type != "cuda":
        raise ValueError("ingest_boxes works on device tensors")
    M, N = obst.shape[0], obst.shape[-1]
    K = box_id.shape[0]
    for t, name in ((obst, "obst"), (slot_id, "slot_id"), (count, "count")):
        if not t.is_contiguous():
            raise ValueError("%s must be contiguous (it is updated in place)" % name)
    obst = _chk(obst, (M, nv.NFIELD, N), dt, dev, "obst")
    box = _chk(box, (K, nv.BOX_FIELDS, N), dt, dev, "box")
    box_id = _chk(box_id, (K, N), torch.int32, dev, "box_id")
    slot_id = _chk(slot_id, (M, N), torch.int32, dev, "slot_id")
    count = _chk(count, (N,), torch.int32, dev, "count")
    if dropped is not None:
        dropped = _chk(dropped, (N,), torch.int32, dev, "dropped")
    with torch.cuda.device(dev):
        nv.check(getattr(L, "sccav_ingest_boxes_" + _SFX[dt])(int(obs_type), int(mode), float(buffer), M, K, N, _ptr(box_id), _ptr(box),
                                                             _ptr(slot_id), _ptr(obst), _ptr(count), _ptr(dropped), _stream(dev)))
    return dropped


def spline_courses(wx: torch.Tensor, wy: torch.Tensor, ds: float = 0.1, P_max: Optional[int] = None, curvature: bool = False):
    """KC: C courses at once from way-points ``wx``, ``wy`` [C, K] (calc_spline_course,
    cubic_spline_planner.py:178-190).  Returns (cx, cy, cyaw [C, P_max], np [C] int32[, ck]); course c has
    ``np[c]`` samples, rows are padded with NaN beyond.  ``P_max`` defaults to a bound from the way-point
    polygon length (the spline is longer than its chords only through rounding of ceil)."""
    L = nv.lib()
    nv.require_cuda()
    dt, dev = wx.dtype, wx.device
    if dev.type != "cuda":
        raise ValueError("spline_courses works on device tensors")
    C_, K = wx.shape
    wx = _chk(wx, (C_, K), dt, dev, "wx")
    wy = _chk(wy, (C_, K), dt, dev, "wy")
    if P_max is None:
        seg = torch.hypot(wx[:, 1:] - wx[:, :-1], wy[:, 1:] - wy[:, :-1]).sum(dim=1)
        P_max = int(torch.ceil(seg.max() / ds).item()) + 1
    out = torch.full((4 if curvature else 3, C_, int(P_max)), float("nan"), dtype=dt, device=dev)
    npts = torch.zeros((C_,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        nv.check(getattr(L, "sccav_spline_course_" + _SFX[dt])(C_, K, _ptr(wx), _ptr(wy), float(ds), int(P_max), _ptr(out[0]), _ptr(out[1]),
                                                              _ptr(out[2]), _ptr(out[3]) if curvature else None, _ptr(npts), _stream(dev)))
    return (out[0], out[1], out[2], npts) + ((out[3],) if curvature else ())


def fit_lanes(x: torch.Tensor, y: torch.Tensor, n: int = 3, sigma: Optional[torch.Tensor] = None,
              count: Optional[torch.Tensor] = None):
    """KL: weighted polynomial fits of C lanes at once -- the batched ``PolyLane.fit_polynomial_curve``
    (cbf/obstacles.py:715-773).  ``x``, ``y`` [K, C] points (fixed points appended by the caller with their small
    sigma, as the reference does), ``sigma`` [K, C] or None (= 10), ``count`` [C] int32 points per lane or None.
    Returns (coeffs [6, C] = c0..c5 of y = sum c_i x^i -- the coefficient fields of a LANE slot, status [C] int32)."""
    L = nv.lib()
    nv.require_cuda()
    dt, dev = x.dtype, x.device
    if dev.type != "cuda":
        raise ValueError("fit_lanes works on device tensors")
    K, C_ = x.shape
    x = _chk(x, (K, C_), dt, dev, "x")
    y = _chk(y, (K, C_), dt, dev, "y")
    if sigma is not None:
        sigma = _chk(sigma, (K, C_), dt, dev, "sigma")
    if count is not None:
        count = _chk(count, (C_,), torch.int32, dev, "count")
    coeffs = torch.empty((6, C_), dtype=dt, device=dev)
    status = torch.empty((C_,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        nv.check(getattr(L, "sccav_fit_lanes_" + _SFX[dt])(C_, K, _ptr(x), _ptr(y), _ptr(sigma), _ptr(count), int(n), _ptr(coeffs),
                                                         _ptr(status), _stream(dev)))
    return coeffs, status


def actuator_shaping(u: torch.Tensor, throttle_prev: torch.Tensor, brake_prev: torch.Tensor, max_steer: float = 1.0,
                     rate: float = 0.1, reset_brake: bool = False):
    """KA: (a, delta) -> (throttle, brake, steer) as the CARLA drivers do after ``solve_cbf``
    (multi_obstacle_CBF_local_with_lanes.py:955-980): tanh map, saturation to [0, 1], increase limited to
    ``rate`` per tick, steering clamp.  ``throttle_prev`` / ``brake_prev`` [N] are updated in place.
    ``reset_brake`` zeroes the brake while throttling (the driver keeps its last value)."""
    L = nv.lib()
    nv.require_cuda()
    N = u.shape[-1]
    dt, dev = u.dtype, u.device
    if dev.type != "cuda":
        raise ValueError("actuator_shaping works on device tensors")
    for t, name in ((throttle_prev, "throttle_prev"), (brake_prev, "brake_prev")):
        if not t.is_contiguous():
            raise ValueError("%s must be contiguous (it is updated in place)" % name)
    u = _chk(u, (2, N), dt, dev, "u")
    throttle_prev = _chk(throttle_prev, (N,), dt, dev, "throttle_prev")
    brake_prev = _chk(brake_prev, (N,), dt, dev, "brake_prev")
    out = torch.empty((3, N), dtype=dt, device=dev)
    with torch.cuda.device(dev):
        nv.check(getattr(L, "sccav_actuator_shaping_" + _SFX[dt])(N, _ptr(u), float(max_steer), float(rate),
                                                                 nv.ACT_RESET_BRAKE if reset_brake else 0, _ptr(throttle_prev),
                                                                 _ptr(brake_prev), _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _stream(dev)))
    return out[0], out[1], out[2]


def rollout_launch_info(slot_desc, N: int, P: int, dtype=torch.float64) -> Dict[str, int]:
    """Launch shape of the rollout kernel for (slot_desc, N, P) on the current device (no launch).
    `slot_desc`: the slot descriptors of the batch, or an int M meaning M per-vehicle ellipses."""
    L = _need_cuda_lib()
    info = (C.c_int32 * 8)()
    if isinstance(slot_desc, int):
        slot_desc = [0] * slot_desc
    sd = bytes(bytearray(int(d) for d in slot_desc))
    nv.check(getattr(L, "sccav_rollout_launch_info_" + _SFX[dtype])(sd, len(sd), int(N), int(P), info))
    keys = ("grid", "block", "smem_bytes", "registers", "max_threads_per_block", "ctas_per_sm", "course_in_smem", "sms")
    return dict(zip(keys, [int(v) for v in info]))


def drive_ticks(params: Params, slot_desc, n_fixed: int, obst: torch.Tensor, traj, target_idx: torch.Tensor, carry: torch.Tensor,
                T: int, state0: Optional[torch.Tensor] = None, ego: Optional[torch.Tensor] = None, box_id: Optional[torch.Tensor] = None,
                box: Optional[torch.Tensor] = None, dt: Optional[torch.Tensor] = None, count: Optional[torch.Tensor] = None,
                kp: float = 1.0, ki: float = 0.01, kd: float = 0.01, rad_to_steer: float = 1.0, max_steer_cmd: float = 1.0,
                rate: float = 0.1, cone_buffer: float = 1.5, act_flags: int = 0, want_u: bool = True, want_state: bool = True):
    """KD: T ticks of the CARLA driver's loop (multi_obstacle_CBF_local_with_lanes.py:861-983) for N egos in ONE launch:
    LateralStanley.control (class form) -> PID1 with the tick's dt -> lanes + one fresh CollisionCone2D per box -> solve_cbf
    -> throttle / brake / steer.  ``ego`` [T,4,N] = the simulator's ego states (or None: stand-in plant update_com from
    ``state0`` [4,N]); ``box_id`` [T,K,N] int32 / ``box`` [T,K,6,N]; ``dt`` [T]; ``traj`` = (x, y, yaw, v) [P];
    ``obst`` [M,8,N] whose first ``n_fixed`` slots are the caller's (lanes); ``target_idx`` int32 [N] and ``carry`` [4,N]
    (PID e_prev, integral, throttle_prev, brake_prev) are updated IN PLACE.  Returns a dict: act [T,3,N], u [T,2,N],
    active_mask [T,N], target_idx [T,N], state [4,N]."""
    L = _need_cuda_lib()
    sd = slot_bytes(slot_desc)
    M = len(sd)
    N = obst.shape[-1] if M > 0 else target_idx.shape[0]
    dtp, dev = carry.dtype, carry.device
    if not carry.is_cuda:
        raise ValueError("drive_ticks takes CUDA tensors")
    for t, name in ((obst, "obst"), (target_idx, "target_idx"), (carry, "carry")):
        if not t.is_contiguous():
            raise ValueError("%s must be contiguous (it is updated in place)" % name)
    obst = _chk(obst, (M, nv.NFIELD, N), dtp, dev, "obst")
    carry = _chk(carry, (4, N), dtp, dev, "carry")
    target_idx = _chk(target_idx, (N,), torch.int32, dev, "target_idx")
    tx, ty, tyaw, tv = traj
    P = tx.shape[0]
    tx = _chk(tx, (P,), dtp, dev, "traj_x"); ty = _chk(ty, (P,), dtp, dev, "traj_y")
    tyaw = _chk(tyaw, (P,), dtp, dev, "traj_yaw"); tv = _chk(tv, (P,), dtp, dev, "traj_v")
    K = 0
    if box_id is not None:
        K = box_id.shape[1]
        box_id = _chk(box_id, (T, K, N), torch.int32, dev, "box_id")
        box = _chk(box, (T, K, nv.BOX_FIELDS, N), dtp, dev, "box")
    if ego is not None:
        ego = _chk(ego, (T, 4, N), dtp, dev, "ego")
    if state0 is not None:
        state0 = _chk(state0, (4, N), dtp, dev, "state0")
    if dt is not None:
        dt = _chk(dt, (T,), dtp, dev, "dt")
    pv, keep = _pv(N, dtp, dev, count=count)
    dp = nv.DriveParams(kp, ki, kd, rad_to_steer, max_steer_cmd, rate, cone_buffer, int(act_flags), 0)
    act = torch.empty((T, 3, N), dtype=dtp, device=dev)
    u = torch.empty((T, 2, N), dtype=dtp, device=dev) if want_u else None
    mask = torch.empty((T, N), dtype=torch.int32, device=dev)
    tidx = torch.empty((T, N), dtype=torch.int32, device=dev)
    st = torch.empty((4, N), dtype=dtp, device=dev) if want_state else None
    with torch.cuda.device(dev):
        nv.check(getattr(L, "sccav_drive_ticks_" + _SFX[dtp])(
            C.byref(params), C.byref(dp), sd, M, int(n_fixed), N, int(T), _ptr(state0), _ptr(ego), K, _ptr(box_id), _ptr(box), _ptr(dt),
            _ptr(obst), _ptr(tx), _ptr(ty), _ptr(tyaw), _ptr(tv), P, C.byref(pv), _ptr(target_idx), _ptr(carry), _ptr(act), _ptr(u),
            _ptr(mask), _ptr(tidx), _ptr(st), _stream(dev)))
    return {"act": act, "u": u, "active_mask": mask, "target_idx": tidx, "state": st}
