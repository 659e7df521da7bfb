"""Synthetic scenario batches for the BASELINE.json configurations.

Everything is generated on the host in float64 with ``numpy.random.default_rng([seed, block])``
per block of 65,536 scenarios, so a scenario's inputs depend only on its GLOBAL index -- the
batch a rank receives is independent of how many GPUs share the job.

Layouts follow include/sccav_cbf.h:  state [4, N],  obst [M, 8, N].
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _native as nv
from .course import config1_course

BLOCK = 65536


@dataclass
class ScenarioBatch:
    name: str
    state: np.ndarray                      # [4, N]
    slot_desc: List[int]                   # M entries
    obst: Optional[np.ndarray]             # [M, 8, N]
    course: Optional[Tuple[np.ndarray, np.ndarray, np.ndarray]]
    params: Dict[str, object] = field(default_factory=dict)   # overrides for ops.make_params
    T: int = 1000
    alpha: Optional[np.ndarray] = None     # per-vehicle overrides
    R: Optional[np.ndarray] = None
    target_speed: Optional[np.ndarray] = None

    @property
    def N(self) -> int:
        return self.state.shape[1]

    @property
    def M(self) -> int:
        return len(self.slot_desc)


def _blocks(lo: int, hi: int):
    b = lo // BLOCK
    while b * BLOCK < hi:
        s, e = max(lo, b * BLOCK), min(hi, (b + 1) * BLOCK)
        yield b, s - b * BLOCK, e - b * BLOCK
        b += 1


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block partition of the scenario axis."""
    per = (n_total + world - 1) // world
    lo = min(n_total, rank * per)
    return lo, min(n_total, lo + per)


def _init_states(rng, n):
    """x~U(-2,2), y~U(3,7), yaw~U(10deg,30deg), v~U(8,12) around the reference's start (sce.py:595)."""
    s = np.empty((4, n))
    s[0] = rng.uniform(-2.0, 2.0, n)
    s[1] = rng.uniform(3.0, 7.0, n)
    s[2] = np.radians(rng.uniform(10.0, 30.0, n))
    s[3] = rng.uniform(8.0, 12.0, n)
    return s


def _ellipses_on_course(rng, n, M, course, buffer=0.5):
    """M static ellipses per vehicle near the course: centre = course point ~U{200..1900} shifted
    ~U(-6,6) m along the course normal; a~U(2,6), b~U(1,3), theta~U(-pi,pi); buffer added to a, b
    as Ellipse2D.__init__ does (cbf/obstacles.py:159-160)."""
    cx, cy, cyaw = course
    o = np.zeros((M, nv.NFIELD, n))
    idx = rng.integers(200, 1901, size=(M, n))
    off = rng.uniform(-6.0, 6.0, (M, n))
    o[:, 0] = cx[idx] - off * np.sin(cyaw[idx])
    o[:, 1] = cy[idx] + off * np.cos(cyaw[idx])
    o[:, 2] = rng.uniform(2.0, 6.0, (M, n)) + buffer
    o[:, 3] = rng.uniform(1.0, 3.0, (M, n)) + buffer
    o[:, 4] = rng.uniform(-np.pi, np.pi, (M, n))
    return o


def config1(kind: str = "cone") -> ScenarioBatch:
    """BASELINE config #1: the reference's single-vehicle run (stanley_controller_ellipse.py main()).
    kind: 'cone' (CBF_TYPE 4, the committed default, golden beta_vs_time.mat), 'ellipse_dbm'
    (CBF_A, CBF_TYPE 2) or 'ellipse_kbm' (CBF, CBF_TYPE 0)."""
    course = config1_course()
    cx, cy, _ = course
    oi = int((len(cx) - 1) * 0.75)                                  # sce.py:610
    state = np.array([[-0.0], [5.0], [np.radians(20.0)], [10.0]])  # sce.py:595
    o = np.zeros((1, nv.NFIELD, 1))
    prm: Dict[str, object] = dict(terminate=1)
    if kind == "cone":
        slot = [nv.SLOT_CONE]
        o[0, :, 0] = [cx[oi], cy[oi], 0.0, 0.0, np.hypot(20, 10) / 2 + 1.5, 0.0, 0.0, 0.0]   # sce.py:645,721,737
        prm.update(R=[0.5, 0.0, 0.0, 0.5])                          # sce.py:741
    else:
        slot = [nv.SLOT_ELLIPSE]
        o[0, :, 0] = [cx[oi], cy[oi], 20.0, 10.0, 0.0, 0.0, 0.0, 0.0]                      # sce.py:608-609
        if kind == "ellipse_kbm":
            prm.update(model=nv.MODEL_KBM, kbm_driver_delta=1)
    return ScenarioBatch("config1_" + kind, state, slot, o, course, prm, T=10 ** 6)


def config2(n_total: int = 65536, M: int = 8, T: int = 1000, seed: int = 0, lo: int = 0, hi: Optional[int] = None) -> ScenarioBatch:
    """BASELINE config #2: N vehicles x M static ellipses, T-step closed loop, DBM + Stanley + update_com."""
    hi = n_total if hi is None else hi
    course = config1_course()
    st, ob = [], []
    for b, s, e in _blocks(lo, hi):
        rng = np.random.default_rng([seed, b])
        st.append(_init_states(rng, BLOCK)[:, s:e])
        ob.append(_ellipses_on_course(rng, BLOCK, M, course)[:, :, s:e])
    state = np.ascontiguousarray(np.concatenate(st, axis=1))
    obst = np.ascontiguousarray(np.concatenate(ob, axis=2))
    # "static ellipses" (BASELINE config 2): the slots say so -- their velocity fields are not read
    return ScenarioBatch("config2", state, [nv.SLOT_ELLIPSE | nv.SLOT_STATIC] * M, obst, course, dict(), T=T)


def config3(n_total: int = 262144, M: int = 16, T: int = 600, seed: int = 1, lo: int = 0, hi: Optional[int] = None,
            stanley: bool = False) -> ScenarioBatch:
    """BASELINE config #3: radial-dynamic obstacles (test_scripts/radial_dynamic_obstacles.py).
    M seeker circles per vehicle: r~U(1.5,2.0) (rdo.py:55,151), spawn distance ~U(10,20) (:56,165),
    angle ~U(0,2pi) (:158) around the ego start; initial seeker speed = ego speed (:186);
    ego starts at rest at the origin (:330) with u_ref = (0,0) (:444), dt = 1/30, gamma = kv = 1."""
    hi = n_total if hi is None else hi
    st, ob = [], []
    for b, s, e in _blocks(lo, hi):
        rng = np.random.default_rng([seed, b])
        n = BLOCK
        if stanley:
            s0 = _init_states(rng, n)
        else:
            s0 = np.zeros((4, n))
        o = np.zeros((M, nv.NFIELD, n))
        r = rng.uniform(1.5, 2.0, (M, n))
        ang = rng.uniform(0.0, 2 * np.pi, (M, n))
        dist = rng.uniform(10.0, 20.0, (M, n))
        o[:, 0] = s0[0] + dist * np.cos(ang)
        o[:, 1] = s0[1] + dist * np.sin(ang)
        o[:, 2] = r
        o[:, 3] = r
        o[:, 4] = 1.0                                            # kv (rdo.py:463)
        yaw = np.arctan2(s0[1] - o[:, 1], s0[0] - o[:, 0])       # rdo.py:174
        o[:, 5] = s0[3] * np.cos(yaw)                            # update_velocity_by_magnitude(state.v)
        o[:, 6] = s0[3] * np.sin(yaw)
        st.append(s0[:, s:e]); ob.append(o[:, :, s:e])
    state = np.ascontiguousarray(np.concatenate(st, axis=1))
    obst = np.ascontiguousarray(np.concatenate(ob, axis=2))
    prm: Dict[str, object] = dict(seeker=1, dt=1.0 / 30.0, alpha=1.0,
                                  nominal=nv.NOMINAL_STANLEY if stanley else nv.NOMINAL_CONST, uref0=0.0, uref1=0.0)
    return ScenarioBatch("config3", state, [nv.SLOT_RADIAL] * M, obst, config1_course() if stanley else None, prm, T=T)


# two global lane boundaries for config #4 (PolyLane default buffer 1.5, cbf/obstacles.py:551):
# a cubic above the first straight of the course and a straight line below the whole course
LANE_CUBIC = [10.5, 0.012, -2.0e-4, 1.0e-6, 0.0, 0.0]
LANE_STRAIGHT = [-46.0, 0.02, 0.0, 0.0, 0.0, 0.0]


def config4(n_total: int = 1048576, M: int = 8, T: int = 1000, seed: int = 2, lo: int = 0, hi: Optional[int] = None) -> ScenarioBatch:
    """BASELINE config #4: config-2 ellipses + 2 shared lane barriers (slots M, M+1)."""
    base = config2(n_total, M, T, seed, lo, hi)
    n = base.N
    lanes = np.zeros((2, nv.NFIELD, n))
    lanes[0, 0, :] = 1.5
    lanes[0, 1:7, :] = np.array(LANE_CUBIC)[:, None]
    lanes[1, 0, :] = 1.5
    lanes[1, 1:7, :] = np.array(LANE_STRAIGHT)[:, None]
    obst = np.ascontiguousarray(np.concatenate([base.obst, lanes], axis=0))
    slots = base.slot_desc + [nv.SLOT_LANE | nv.SLOT_SHARED, nv.SLOT_LANE | nv.SLOT_SHARED]
    return ScenarioBatch("config4", base.state, slots, obst, base.course, dict(), T=T)


def config5(n_total: int = 16777216, T: int = 300, seed: int = 3, lo: int = 0, hi: Optional[int] = None) -> ScenarioBatch:
    """BASELINE config #5: Monte-Carlo sweep of the beta_vs_time experiment (config-1 cone scenario)
    over the CBF gain alpha in logspace(-1,1), the steering weight R[1,1] in logspace(-1,2) and the
    initial lateral offset y0~U(3,7): one scenario per global index, grid index = mixed radix."""
    hi = n_total if hi is None else hi
    c1 = config1("cone")
    n = hi - lo
    gi = np.arange(lo, hi, dtype=np.int64)
    na, nr = 256, 256
    ia = gi % na
    ir = (gi // na) % nr
    alpha = 10.0 ** (-1.0 + 2.0 * ia / (na - 1))
    r11 = 10.0 ** (-1.0 + 3.0 * ir / (nr - 1))
    y0 = np.empty(n)
    pos = 0
    for b, s, e in _blocks(lo, hi):
        rng = np.random.default_rng([seed, b])
        y0[pos:pos + (e - s)] = rng.uniform(3.0, 7.0, BLOCK)[s:e]
        pos += e - s
    state = np.repeat(c1.state, n, axis=1)
    state[1] = y0
    obst = np.ascontiguousarray(np.repeat(c1.obst, n, axis=2))
    R = np.zeros((4, n))
    R[0] = 0.5
    R[3] = r11
    return ScenarioBatch("config5", np.ascontiguousarray(state), c1.slot_desc, obst, c1.course,
                         dict(terminate=1), T=T, alpha=alpha, R=R)


def roads(n_roads: int, per_road: int, M: int = 8, seed: int = 4, dtype=None, device=None):
    """Monte-Carlo over ROADS: every road is the config-1 course with its inner way-points moved by
    U(-8, 8) m -- so the roads differ in shape and length -- generated on the device by the course kernel
    (sccav_spline_course_*); per road, config-2 style vehicles and M static ellipses placed along THAT road.
    Returns ((cx, cy, cyaw [C, P_max], np [C]) device tensors, np host array, state [4, N], obst [M, 8, N]) with
    N = n_roads * per_road and the vehicles grouped by road -- the input of ONE sccav_rollout_roads_* launch."""
    import torch
    from . import ops
    dtype = torch.float64 if dtype is None else dtype
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
    rng = np.random.default_rng([seed, n_roads])
    wx = np.tile(np.array([0.0, 100.0, 100.0, 50.0, 60.0]), (n_roads, 1))
    wy = np.tile(np.array([0.0, 0.0, -30.0, -20.0, 0.0]), (n_roads, 1))
    wx[:, 1:] += rng.uniform(-8, 8, (n_roads, 4)); wy[:, 1:] += rng.uniform(-8, 8, (n_roads, 4))
    if n_roads > 1:
        wx[1, 4] += 25.0                                             # one road clearly longer than the others
    cx, cy, cyaw, npts = ops.spline_courses(torch.from_numpy(wx).to(device=device, dtype=dtype),
                                            torch.from_numpy(wy).to(device=device, dtype=dtype), ds=0.1)
    nph = npts.cpu().numpy()
    cxh, cyh, cyawh = cx.double().cpu().numpy(), cy.double().cpu().numpy(), cyaw.double().cpu().numpy()
    states, obsts = [], []
    for c in range(n_roads):
        n = int(nph[c])
        states.append(_init_states(rng, per_road))
        o = np.zeros((M, nv.NFIELD, per_road))
        idx = rng.integers(150, n - 150, size=(M, per_road))
        off = rng.uniform(-6.0, 6.0, (M, per_road))
        o[:, 0] = cxh[c][idx] - off * np.sin(cyawh[c][idx]); o[:, 1] = cyh[c][idx] + off * np.cos(cyawh[c][idx])
        o[:, 2] = rng.uniform(2.0, 6.0, (M, per_road)) + 0.5; o[:, 3] = rng.uniform(1.0, 3.0, (M, per_road)) + 0.5
        o[:, 4] = rng.uniform(-np.pi, np.pi, (M, per_road))
        obsts.append(o)
    return (cx, cy, cyaw, npts), nph, np.ascontiguousarray(np.concatenate(states, axis=1)), np.ascontiguousarray(np.concatenate(obsts, axis=2))
