"""ctypes wrapper of oracle/oracle.c (TEST INFRASTRUCTURE -- see the header of oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this.  Arrays are numpy float64 in the layouts of include/sccav_cbf.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")


class Params(C.Structure):
    _fields_ = [
        ("model", C.c_int32), ("nominal", C.c_int32), ("terminate", C.c_int32), ("seeker", C.c_int32),
        ("kbm_driver_delta", C.c_int32), ("record_stride", C.c_int32), ("flags", C.c_int32), ("reserved1", C.c_int32),
        ("alpha", C.c_double), ("lr", C.c_double), ("lf", C.c_double), ("L", C.c_double),
        ("max_steer", C.c_double), ("dt", C.c_double), ("k_stanley", C.c_double), ("ks_stanley", C.c_double),
        ("Kp", C.c_double), ("target_speed", C.c_double), ("t_max", C.c_double),
        ("R", C.c_double * 4), ("seeker_k", C.c_double), ("seeker_vmin", C.c_double),
        ("uref0", C.c_double), ("uref1", C.c_double), ("sadbm_dt", C.c_double),
    ]


class PerVehicle(C.Structure):
    _fields_ = [("alpha", C.c_void_p), ("R", C.c_void_p), ("target_speed", C.c_void_p), ("count", C.c_void_p),
                ("aug", C.c_void_p)]


class RolloutOut(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("state", "steps", "target_idx", "n_active", "n_infeasible", "h_min",
                                          "beta_min", "beta_max", "beta_int", "traj", "traj_idx", "traj_mask", "n_evals")]


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE] + (["-B"] if force else []), check=True, stdout=subprocess.DEVNULL)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def default_params(**kw) -> Params:
    """Defaults of stanley_controller_ellipse.py:52-58,590 (same values as sccav_default_params)."""
    p = Params()
    p.model = 0; p.nominal = 0
    p.alpha = 1.0; p.L = 2.9; p.lr = 2.9 / 2; p.lf = 2.9 - 2.9 / 2
    p.max_steer = float(np.radians(30.0)); p.dt = 0.1; p.k_stanley = 0.5; p.ks_stanley = 0.0
    p.Kp = 1.0; p.target_speed = 30.0 / 3.6; p.t_max = 30.0
    p.R[0] = 1.0; p.R[3] = 1.0
    p.seeker_k = 0.2; p.seeker_vmin = 3.0
    for k, v in kw.items():
        if k == "R":
            flat = np.asarray(v, dtype=np.float64).reshape(-1)
            for i in range(4):
                p.R[i] = float(flat[i])
        else:
            if not hasattr(p, k):
                raise TypeError(k)
            setattr(p, k, v)
    return p


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _pv(alpha, R, target_speed, count=None):
    pv = PerVehicle()
    keep = [_f64(alpha), _f64(R), _f64(target_speed), None if count is None else np.ascontiguousarray(count, dtype=np.int32)]
    if keep[3] is not None:
        pv.count = keep[3].ctypes.data
    if keep[0] is not None:
        pv.alpha = keep[0].ctypes.data
    if keep[1] is not None:
        pv.R = keep[1].ctypes.data
    if keep[2] is not None:
        pv.target_speed = keep[2].ctypes.data
    return pv, keep


def filter_step(params: Params, slot_desc, state, obst, u_ref, alpha=None, R=None, rows=False, nthreads=0, count=None):
    sd = bytes(int(d) & 0xFF for d in slot_desc)
    state, obst, u_ref = _f64(state), _f64(obst), _f64(u_ref)
    M, N = len(sd), state.shape[1]
    u = np.empty((2, N)); mask = np.empty(N, dtype=np.uint32); status = np.empty(N, dtype=np.uint8); hmin = np.empty(N)
    A = np.empty((2, M, N)) if rows else None
    b = np.empty((M, N)) if rows else None
    pv, keep = _pv(alpha, R, None, count)
    rc = lib().orc_filter_step(C.byref(params), sd, C.c_int32(M), C.c_int64(N), _p(state), _p(obst), _p(u_ref), C.byref(pv),
                               _p(u), _p(mask), _p(status), _p(hmin), _p(A), _p(b), C.c_int(nthreads))
    if rc != 0:
        raise ValueError("orc_filter_step rc=%d" % rc)
    out = dict(u=u, mask=mask, status=status, h_min=hmin)
    if rows:
        out.update(A=A, b=b)
    return out


def rollout(params: Params, slot_desc, state, obst, course, T, alpha=None, R=None, target_speed=None,
            record_stride=0, nthreads=0, count=None):
    import copy
    sd = bytes(int(d) & 0xFF for d in slot_desc)
    state = _f64(state)
    M, N = len(sd), state.shape[1]
    obst = None if obst is None else np.array(obst, dtype=np.float64, order="C", copy=True)
    if course is not None:
        cx, cy, cyaw = (_f64(c) for c in course)
        P = len(cx)
    else:
        cx = cy = cyaw = None
        P = 0
    prm = copy.copy(params)
    prm.record_stride = int(record_stride)
    res = dict(state=np.empty((4, N)), steps=np.empty(N, np.int32), target_idx=np.empty(N, np.int32),
               n_active=np.empty(N, np.int32), n_infeasible=np.empty(N, np.int32), h_min=np.empty(N),
               beta_min=np.empty(N), beta_max=np.empty(N), beta_int=np.empty(N))
    if record_stride > 0:
        trec = (T + record_stride - 1) // record_stride
        res["traj"] = np.full((trec, 7, N), np.nan)
        res["traj_idx"] = np.full((trec, N), -1, np.int32)
        res["traj_mask"] = np.zeros((trec, N), np.uint32)
    ro = RolloutOut()
    for k, v in res.items():
        setattr(ro, k, v.ctypes.data)
    pv, keep = _pv(alpha, R, target_speed, count)
    rc = lib().orc_rollout(C.byref(prm), sd, C.c_int32(M), C.c_int64(N), C.c_int32(int(T)), _p(state), _p(obst), _p(cx), _p(cy),
                           _p(cyaw), C.c_int32(P), C.byref(pv), C.byref(ro), C.c_int(nthreads))
    if rc != 0:
        raise ValueError("orc_rollout rc=%d" % rc)
    res["obst"] = obst
    return res


def num_threads() -> int:
    return int(lib().orc_num_threads())
