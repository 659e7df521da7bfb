"""ctypes binding of libsccav_cbf.so (the C-ABI declared in include/sccav_cbf.h).

There is no CPU fallback: if the library is missing this module raises, and every op raises
when no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# SCCAV_CBF_LIB: developer override (scripts/exp_variants.py times alternative builds of the same ABI)
LIB_PATH = os.environ.get("SCCAV_CBF_LIB") or os.path.join(HERE, "libsccav_cbf.so")

NFIELD = 8
MAX_ROWS = 32
TRAJ_FIELDS = 7

SLOT_ELLIPSE, SLOT_CONE, SLOT_LANE, SLOT_RADIAL, SLOT_DISTANCE, SLOT_ELLIPSE_PREP, SLOT_LANE_SQRT = 0, 1, 2, 3, 4, 5, 6
SLOT_TYPE_MASK = 0x3F
SLOT_STATIC = 0x40
SLOT_SHARED = 0x80
FLAG_PREPARED_ROWS = 1
FLAG_QP_ENUMERATE = 2
FLAG_FUSED_STEER = 4
FLAG_BETA_IO = 8
FLAG_SEEKER_DIRECT = 16
BOX_FIELDS = 6
INGEST_UPDATE, INGEST_REBUILD = 0, 1
ACT_RESET_BRAKE = 1
MODEL_DBM, MODEL_KBM, MODEL_NONE, MODEL_DUM, MODEL_SADBM = 0, 1, 2, 3, 4
NOMINAL_STANLEY, NOMINAL_CONST = 0, 1
STATUS_INACTIVE, STATUS_ACTIVE, STATUS_INFEASIBLE = 0, 1, 2
OK, EINVAL, ECUDA, ENOMEM = 0, -1, -2, -3


class Params(C.Structure):
    """struct sccav_params (include/sccav_cbf.h)."""
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


class DriveParams(C.Structure):
    _fields_ = [("kp", C.c_double), ("ki", C.c_double), ("kd", C.c_double), ("rad_to_steer", C.c_double),
                ("max_steer_cmd", C.c_double), ("rate", C.c_double), ("cone_buffer", C.c_double),
                ("act_flags", C.c_int32), ("reserved", C.c_int32)]


class RolloutOut(C.Structure):
    _fields_ = [
        ("state", C.c_void_p), ("steps", C.c_void_p), ("target_idx", C.c_void_p), ("n_active", C.c_void_p),
        ("n_infeasible", C.c_void_p), ("h_min", C.c_void_p), ("beta_min", C.c_void_p), ("beta_max", C.c_void_p),
        ("beta_int", C.c_void_p), ("traj", C.c_void_p), ("traj_idx", C.c_void_p), ("traj_mask", C.c_void_p),
        ("n_evals", C.c_void_p),
    ]


class SccavError(RuntimeError):
    pass


_lib = None

# every symbol include/sccav_cbf.h declares (tests check the library exports all of them)
SYMBOLS = [
    "sccav_version", "sccav_last_error", "sccav_device_ok", "sccav_default_params",
    "sccav_barrier_rows_f64", "sccav_barrier_rows_f32", "sccav_qp2_solve_f64", "sccav_qp2_solve_f32",
    "sccav_filter_step_f64", "sccav_filter_step_f32", "sccav_rollout_f64", "sccav_rollout_f32",
    "sccav_filter_step_host_f64", "sccav_filter_step_host_f32", "sccav_rollout_host_f64", "sccav_rollout_host_f32",
    "sccav_measure_fma_peak", "sccav_launch_count", "sccav_debug_course_index_host", "sccav_debug_cover_host",
    "sccav_rollout_launch_info_f64", "sccav_rollout_launch_info_f32",
    "sccav_barrier_partials_f64", "sccav_barrier_partials_f32", "sccav_stanley_control_f64", "sccav_stanley_control_f32",
    "sccav_prepare_obstacles_f64", "sccav_prepare_obstacles_f32", "sccav_ingest_boxes_f64", "sccav_ingest_boxes_f32",
    "sccav_actuator_shaping_f64", "sccav_actuator_shaping_f32", "sccav_spline_course_f64", "sccav_spline_course_f32",
    "sccav_fit_lanes_f64", "sccav_fit_lanes_f32", "sccav_rollout_roads_f64", "sccav_rollout_roads_f32",
    "sccav_trim_pool", "sccav_drive_ticks_f64", "sccav_drive_ticks_f32",
    "sccav_pipeline_create_f64", "sccav_pipeline_create_f32", "sccav_pipeline_submit_f64", "sccav_pipeline_submit_f32",
    "sccav_pipeline_wait_f64", "sccav_pipeline_wait_f32", "sccav_pipeline_destroy_f64", "sccav_pipeline_destroy_f32",
]


def lib() -> C.CDLL:
    """Load libsccav_cbf.so (built in-tree by sccav_cbf_b200.build); raise if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SccavError(
            "libsccav_cbf.so not found at %s -- build it with `python -m sccav_cbf_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    PP, PV, RO = C.POINTER(Params), C.POINTER(PerVehicle), C.POINTER(RolloutOut)
    L.sccav_version.restype = C.c_int
    L.sccav_last_error.restype = C.c_char_p
    L.sccav_device_ok.restype = C.c_int
    L.sccav_default_params.argtypes = [PP]
    L.sccav_default_params.restype = None
    L.sccav_measure_fma_peak.argtypes = [i32, C.POINTER(C.c_double)]
    L.sccav_launch_count.restype = i64
    L.sccav_debug_course_index_host.argtypes = [vp, vp, i32, vp, vp, vp, i64, i32, vp, vp, vp]
    L.sccav_debug_cover_host.argtypes = [i32, vp, vp, vp, vp, vp, vp]
    for sfx in ("f64", "f32"):
        f = getattr(L, "sccav_barrier_rows_" + sfx)
        f.argtypes = [PP, C.c_char_p, i32, i64, vp, vp, PV, vp, vp, vp, vp]
        f = getattr(L, "sccav_barrier_partials_" + sfx)
        f.argtypes = [C.c_char_p, i32, i64, vp, vp, vp, vp]
        f = getattr(L, "sccav_prepare_obstacles_" + sfx)
        f.argtypes = [C.c_char_p, i32, i64, vp, vp, vp, vp]
        f = getattr(L, "sccav_spline_course_" + sfx)
        f.argtypes = [i32, i32, vp, vp, C.c_double, i32, vp, vp, vp, vp, vp, vp]
        f = getattr(L, "sccav_fit_lanes_" + sfx)
        f.argtypes = [i32, i32, vp, vp, vp, vp, i32, vp, vp, vp]
        f = getattr(L, "sccav_actuator_shaping_" + sfx)
        f.argtypes = [i64, vp, C.c_double, C.c_double, i32, vp, vp, vp, vp, vp, vp]
        f = getattr(L, "sccav_ingest_boxes_" + sfx)
        f.argtypes = [i32, i32, C.c_double, i32, i32, i64, vp, vp, vp, vp, vp, vp, vp]
        f = getattr(L, "sccav_stanley_control_" + sfx)
        f.argtypes = [PP, i64, vp, vp, vp, vp, vp, i32, vp, vp, vp, vp]
        f = getattr(L, "sccav_rollout_launch_info_" + sfx)
        f.argtypes = [C.c_char_p, i32, i64, i32, vp]
        f = getattr(L, "sccav_rollout_roads_" + sfx)
        f.argtypes = [PP, C.c_char_p, i32, i64, i32, vp, vp, vp, vp, vp, i32, i32, vp, PV, RO, vp]
        getattr(L, "sccav_drive_ticks_" + sfx).argtypes = [PP, C.POINTER(DriveParams), C.c_char_p, i32, i32, i64, i32, vp, vp, i32, vp, vp, vp, vp,
                                                            vp, vp, vp, vp, i32, PV, vp, vp, vp, vp, vp, vp, vp, vp]
        getattr(L, "sccav_pipeline_create_" + sfx).argtypes = [PP, C.c_char_p, i32, i64, i32, vp, vp, vp, i32, vp, i32, C.POINTER(vp)]
        getattr(L, "sccav_pipeline_submit_" + sfx).argtypes = [vp, vp, vp, PV, RO, C.POINTER(i64)]
        getattr(L, "sccav_pipeline_wait_" + sfx).argtypes = [vp, i64]
        getattr(L, "sccav_pipeline_destroy_" + sfx).argtypes = [vp]
        f = getattr(L, "sccav_qp2_solve_" + sfx)
        f.argtypes = [PP, i32, i64, vp, vp, vp, PV, vp, vp, vp, i32, vp]
        for host in ("", "host_"):
            f = getattr(L, "sccav_filter_step_" + host + sfx)
            f.argtypes = [PP, C.c_char_p, i32, i64, vp, vp, vp, PV, vp, vp, vp, vp, vp]
            f = getattr(L, "sccav_rollout_" + host + sfx)
            f.argtypes = [PP, C.c_char_p, i32, i64, i32, vp, vp, vp, vp, vp, i32, PV, RO, vp]
    for name in SYMBOLS:
        fn = getattr(L, name)
        if fn.restype is C.c_int and name not in ("sccav_version", "sccav_device_ok"):
            fn.restype = C.c_int
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != OK:
        msg = lib().sccav_last_error().decode("utf-8", "replace")
        if rc == EINVAL:
            raise ValueError(msg)
        raise SccavError("sccav error %d: %s" % (rc, msg))


def default_params() -> Params:
    p = Params()
    lib().sccav_default_params(C.byref(p))
    return p


def require_cuda() -> None:
    if not lib().sccav_device_ok():
        raise SccavError("no CUDA device: sccav_cbf_b200 has no CPU fallback")
