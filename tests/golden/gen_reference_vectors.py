"""Generate tests/golden/reference_vectors.npz by EXECUTING THE REFERENCE'S OWN CODE.

Run once in the build container (``python tests/golden/gen_reference_vectors.py``); the GPU box
has no /root/reference, so the resulting vectors are committed.  The reference cannot be imported
as-is here (``euclid`` and ``cvxopt`` are absent), so tests/golden/shims supplies stand-ins for
those two packages only (see tests/golden/shims/README.md) -- every barrier value, partial, row
and plant/nominal-controller number below is produced by unmodified reference code:

* ``cbf/obstacles.py``: CollisionCone2D / Ellipse2D / PolyLane methods (PolyLane runs the real
  scipy Newton-CG);
* ``cbf/cbf.py``: DBM_CBF_2DS.solve_cbf row assembly (rows captured inside the shimmed
  ``solvers.cp``) and the delta<->beta conversions;
* ``test_scripts/stanley_controller_ellipse.py``: functions ``CBF``, ``D_CBF``, ``CBF_A``,
  ``CBF_cone``, ``State``, ``pid_control``, ``stanley_control``, ``calc_target_index``,
  ``normalize_angle`` pulled out of the file by AST (the module itself needs matplotlib/imageio/
  cvxpy at import) and driven by a loop that mirrors ``main()`` lines 630-830 without plotting;
* ``test_scripts/radial_dynamic_obstacles.py``: ``single_obstacle_CBF1`` (same AST route).
"""
import ast
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "test_scripts", "PathPlanning", "CubicSpline"))

import cubic_spline_planner  # noqa: E402  (numpy-only; imports cleanly)
from cvxopt import matrix, solvers, sqrt  # noqa: E402  (the shim)
from euclid import Point2, Vector2  # noqa: E402  (the shim)

from cbf.cbf import DBM_CBF_2DS  # noqa: E402
from cbf.obstacles import CollisionCone2D, Ellipse2D, PolyLane  # noqa: E402


def extract(path, names, consts):
    """exec selected top-level defs / constant assignments of a reference script."""
    src = open(path).read()
    tree = ast.parse(src)
    keep = []
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            keep.append(node)
        elif isinstance(node, ast.Assign) and all(isinstance(t, ast.Name) and t.id in consts for t in node.targets):
            keep.append(node)
    mod = ast.Module(body=keep, type_ignores=[])
    ns = dict(np=np, matrix=matrix, solvers=solvers, sqrt=sqrt, Point2=Point2, Vector2=Vector2,
              _NARRAY=object)
    exec(compile(mod, path, "exec"), ns)
    return ns


out = {}
rng = np.random.default_rng(20261017)

# ---------------------------------------------------------------- 1. cone partials (class)
N = 96
cone_in = np.zeros((N, 10))
cone_out = np.zeros((N, 6))
for i in range(N):
    s = np.array([rng.uniform(-50, 50), rng.uniform(-50, 50), rng.uniform(-3.5, 3.5), rng.uniform(0, 15)])
    so = np.array([s[0] + rng.uniform(-30, 30), s[1] + rng.uniform(-30, 30), rng.uniform(-3.5, 3.5), rng.uniform(0, 10)])
    a = rng.uniform(0.5, 12.0)
    beta = 0.0 if i % 3 else rng.uniform(-0.3, 0.3)
    if i == 5:
        so[3] = 0.0
    if i == 6:
        s[3] = 0.0; so[3] = 0.0           # v_rel = 0
    if i == 7:
        so[:2] = s[:2] + 1e-4             # inside ZERO_TOL
    if i == 8:
        a = 100.0                         # inside the disc: cone_boundary = ZERO_TOL
    c = CollisionCone2D(a, s, so, buffer=1.5, beta=beta)
    c.update(s=s)                         # the path DBM_CBF_2DS.update_state takes (obstacles.py:468)
    cone_in[i] = [*s, *so, a, beta]
    cone_out[i] = [float(c.f()), c.dx(), c.dy(), c.dtheta(), c.dv(), c.dt()]
out["cone_in"] = cone_in      # x y th v | cx cy th_o v_o | a(before buffer 1.5) | beta
out["cone_out"] = cone_out    # h h_x h_y h_theta h_v h_t

# ---------------------------------------------------------------- 2. ellipse partials (class)
ell_in = np.zeros((N, 10))
ell_out = np.zeros((N, 4))
for i in range(N):
    x, y = rng.uniform(-50, 50, 2)
    cx, cy = x + rng.uniform(-20, 20), y + rng.uniform(-20, 20)
    a, b = rng.uniform(1, 20), rng.uniform(1, 10)
    th = rng.uniform(-np.pi, np.pi) if i % 4 else 0.0
    buf = rng.uniform(0, 2) if i % 2 else 0.0
    vx, vy = rng.uniform(-5, 5, 2)
    e = Ellipse2D(a=a, b=b, center=Vector2(cx, cy), theta=th, buffer=buf)
    e.s = matrix([x, y, 0.1, 5.0])        # intended semantics: ego state held separately
    e.update_velocity(Vector2(vx, vy))
    ell_in[i] = [x, y, cx, cy, a, b, th, buf, vx, vy]
    ell_out[i] = [e.evaluate(), e.dx(), e.dy(), e.dt()]
out["ellipse_in"] = ell_in
out["ellipse_out"] = ell_out   # h h_x h_y h_t

# ---------------------------------------------------------------- 3. lane (class, real scipy Newton-CG)
lane_in = np.zeros((N, 9))
lane_out = np.zeros((N, 5))
for i in range(N):
    kind = i % 4
    if kind == 0:      # straight lane, as carla multi_obstacle_CBF_local_with_lanes.py:280-297
        co = np.array([rng.uniform(-20, 20), rng.uniform(-0.5, 0.5), 0.0, 0.0])
    elif kind == 1:    # quadratic
        co = np.array([rng.uniform(-20, 20), rng.uniform(-0.5, 0.5), rng.uniform(-0.01, 0.01), 0.0])
    else:              # cubic (default n = 3)
        co = np.array([rng.uniform(-20, 20), rng.uniform(-0.5, 0.5), rng.uniform(-0.01, 0.01), rng.uniform(-1e-4, 1e-4)])
    x = rng.uniform(-40, 40)
    g = co[0] + co[1] * x + co[2] * x * x + co[3] * x ** 3
    y = g + rng.uniform(-6, 6)
    buf = 1.5 if i % 2 else rng.uniform(0, 3)
    s = np.array([x, y, rng.uniform(-3, 3), rng.uniform(1, 12)])
    ln = PolyLane(co, s=s, buffer=buf)
    lane_in[i] = [x, y, *co, buf, s[2], s[3]]
    lane_out[i] = [float(np.ravel(ln.cx)[0]), float(np.ravel(ln.f())[0]), float(np.ravel(ln.dx())[0]),
                   float(np.ravel(ln.dy())[0]), float(np.ravel(ln.eta)[0])]
out["lane_in"] = lane_in       # x y c0 c1 c2 c3 buffer th v
out["lane_out"] = lane_out     # cx h h_x h_y eta

# ---------------------------------------------------------------- 4. DBM_CBF_2DS.solve_cbf (class): rows + output
NC = 64
MAXM = 5
dbm_in = np.zeros((NC, 4 + 2 + 4 + 1 + 1))          # s, u_ref, R, alpha, m
dbm_slot = np.zeros((NC, MAXM, 1 + 8))              # type, fields (oracle layout)
dbm_rows = np.full((NC, MAXM, 3), np.nan)           # A0 A1 b assembled by cbf.py
dbm_out = np.zeros((NC, 5))                         # u0, delta_out, beta_ref, mask, status
for i in range(NC):
    s = np.array([rng.uniform(-20, 20), rng.uniform(-5, 25), rng.uniform(-1.0, 1.0), rng.uniform(2, 12)])
    m = 1 + i % MAXM
    alpha = [1.0, 0.5, 2.0][i % 3]
    Rm = np.diag([0.5, 0.5]) if i % 2 else np.array([[1.0, 0.2], [0.2, 3.0]])
    ctl = DBM_CBF_2DS(alpha=alpha)
    ctl.set_model_params(lr=1.45, lf=1.45)
    for j in range(m):
        if j < 2 and i % 2:
            co = np.array([10.4 + 9.0 * j + rng.uniform(-1, 1), rng.uniform(-0.05, 0.05), 0.0, 0.0]) if i % 4 == 1 else \
                np.array([10.4 + 9.0 * j, rng.uniform(-0.05, 0.05), rng.uniform(-0.002, 0.002), rng.uniform(-2e-5, 2e-5)])
            ctl.obstacle_list2d["lane%d" % j] = PolyLane(co, s=s)
            dbm_slot[i, j] = [2, 1.5, co[0], co[1], co[2], co[3], 0, 0, 0]
        else:
            so = np.array([s[0] + rng.uniform(5, 40), s[1] + rng.uniform(-8, 8), rng.uniform(-3, 3), rng.uniform(0, 8)])
            a = rng.uniform(1, 5)
            ctl.obstacle_list2d[j] = CollisionCone2D(a, s, so)
            dbm_slot[i, j] = [1, so[0], so[1], so[2], so[3], a + 1.5, 0.0, 0, 0]
    ctl.update_state(s=s)
    ctl.set_qp_cost_weight(Rm)
    uref = np.array([rng.uniform(-2, 2), rng.uniform(-0.4, 0.4)])
    solvers.LOG.clear()
    sol, u = ctl.solve_cbf(uref.copy(), return_solver=True)
    log = solvers.LOG[-1]
    dbm_in[i] = [*s, *uref, *Rm.ravel(), alpha, m]
    dbm_rows[i, :m, 0:2] = log["A"]
    dbm_rows[i, :m, 2] = log["b"]
    dbm_out[i] = [u[0], u[1], log["x0"][1], log["mask"], log["status"]]   # x0 = the converted u_ref (cbf.py:185)
out["dbm_in"] = dbm_in
out["dbm_slot"] = dbm_slot
out["dbm_rows"] = dbm_rows
out["dbm_out"] = dbm_out

# ---------------------------------------------------------------- 5. config #1 closed loops (sce.py functions)
SCE = os.path.join(REF, "test_scripts", "stanley_controller_ellipse.py")
ns = extract(SCE,
             names={"State", "pid_control", "stanley_control", "normalize_angle", "calc_target_index",
                    "CBF", "D_CBF", "CBF_A", "CBF_cone", "saturation", "vec_norm"},
             consts={"k", "Kp", "dt", "L", "lr", "lf", "max_steer", "ZERO_TOL"})
ax = [0.0, 100.0, 100.0, 50.0, 60.0]
ay = [0.0, 0.0, -30.0, -20.0, 0.0]
cx, cy, cyaw, ck, sarr = cubic_spline_planner.calc_spline_course(ax, ay, ds=0.1)
out["course"] = np.array([cx, cy, cyaw])


def run_config1(cbf_type):
    """Mirror of main() (stanley_controller_ellipse.py:581-836) for one CBF_TYPE, no plotting."""
    State = ns["State"]
    target_speed = 30.0 / 3.6
    max_simulation_time = 30
    state = State(x=-0.0, y=5.0, yaw=np.radians(20.0), v=10.0)
    last_idx = len(cx) - 1
    time = 0.0
    target_idx, _ = ns["calc_target_index"](state, cx, cy)
    a, b = 20, 10
    obs_idx = int(last_idx * 0.75)
    o_cx, o_cy = cx[obs_idx], cy[obs_idx]
    gamma = 1
    a_cone = np.hypot(a, b) / 2
    Ds = max(a, b) / 2 + 1
    rows = []
    while max_simulation_time >= time and last_idx > target_idx:
        v_ = target_speed
        di, target_idx = ns["stanley_control"](state, cx, cy, cyaw, target_idx)
        pre = [state.x, state.y, state.yaw, state.v]
        solvers.LOG.clear()
        if cbf_type == 0:       # :646-656
            u_des = np.array([v_, v_ * np.tan(di) / ns["L"]])
            s = np.array([state.x, state.y, state.yaw])
            u = ns["CBF"](s, u_des, o_cx, o_cy, a, b, gamma)
            v_cbf, w_cbf = u[0], u[1]
            di_cbf = np.arctan(w_cbf * ns["L"] / v_cbf)
            state.update_by_vel(v_cbf, di_cbf)
            u_out = [v_cbf, di_cbf]
            beta = 0.0
        elif cbf_type == 1:     # :658-668
            u_des = np.array([v_, v_ * np.tan(di) / ns["L"]])
            s = np.array([state.x, state.y, state.yaw])
            u = ns["D_CBF"](s, u_des, o_cx, o_cy, Ds, gamma)
            v_cbf, w_cbf = u[0], u[1]
            di_cbf = np.arctan(w_cbf * ns["L"] / v_cbf)
            state.update_by_vel(v_cbf, di_cbf)
            u_out = [v_cbf, di_cbf]
            beta = 0.0
        elif cbf_type == 2:     # "Without Class" branch :673-681
            a_ = ns["pid_control"](target_speed, state.v)
            beta_ = np.arctan2(ns["lr"] * np.tan(di), ns["lf"] + ns["lr"])
            u_des = np.array([a_, beta_])
            s = np.array([state.x, state.y, state.yaw, state.v])
            u = ns["CBF_A"](s, u_des, o_cx, o_cy, a, b, gamma)
            a_cbf, beta_cbf = u[0], u[1]
            di_cbf = np.arctan2((ns["lf"] + ns["lr"]) * np.tan(beta_cbf), ns["lr"])
            state.update_com(a_cbf, di_cbf)
            u_out = [a_cbf, di_cbf]
            beta = state.beta
        elif cbf_type == 4:     # "With Class" branch :717-748 (the committed default)
            a_ = ns["pid_control"](target_speed, state.v)
            s = np.array([state.x, state.y, state.yaw, state.v])
            s_obs = np.array([o_cx, o_cy, 0, 0])
            ctl = DBM_CBF_2DS(alpha=gamma)
            ctl.set_model_params(lr=ns["lr"], lf=ns["lf"])
            ctl.obstacle_list2d.update({0: CollisionCone2D(a_cone, s, s_obs)})
            ctl.update_state(s=np.array([state.x, state.y, state.yaw, state.v]))
            ctl.set_qp_cost_weight(np.diag([0.5, 0.5]))
            u = ctl.solve_cbf(np.array([a_, di]))
            a_cbf, di_cbf = u[0], u[1]
            state.update_com(a_cbf, di_cbf)
            u_out = [a_cbf, di_cbf]
            beta = state.beta
        log = solvers.LOG[-1]
        time += ns["dt"]
        rows.append(pre + [di, target_idx] + u_out + [beta, log["mask"], log["status"], time,
                                                     log["A"][0, 0], log["A"][0, 1], log["b"][0]])
    return np.array(rows, dtype=np.float64), [state.x, state.y, state.yaw, state.v]


for t in (0, 1, 2, 4):
    rows, fin = run_config1(t)
    out["cfg1_type%d" % t] = rows     # x y yaw v | delta_stanley idx | u0 delta_cbf | beta mask status time | A0 A1 b
    out["cfg1_type%d_final" % t] = np.array(fin)
    print("config1 CBF_TYPE", t, "steps", rows.shape[0], "active", int((rows[:, 9] != 0).sum()), "final idx", int(rows[-1, 5]))

# ---------------------------------------------------------------- 6. single_obstacle_CBF1 (rdo.py)
RDO = os.path.join(REF, "test_scripts", "radial_dynamic_obstacles.py")
nr = extract(RDO, names={"single_obstacle_CBF1"}, consts={"_L", "_lr", "_lf", "_max_steer", "_ZERO_TOL"})
NRD = 64
rad_in = np.zeros((NRD, 13))
rad_out = np.zeros((NRD, 7))
for i in range(NRD):
    s = np.array([rng.uniform(-5, 5), rng.uniform(-5, 5), rng.uniform(-3, 3), rng.uniform(0, 6)])
    ang = rng.uniform(0, 2 * np.pi)
    d = rng.uniform(1.0, 20.0)
    c_obs = np.array([s[0] + d * np.cos(ang), s[1] + d * np.sin(ang)])
    r = rng.uniform(1.5, 2.0)
    spd = rng.uniform(3, 6)
    yaw = np.arctan2(s[1] - c_obs[1], s[0] - c_obs[0])
    c_dot = np.array([spd * np.cos(yaw), spd * np.sin(yaw)])
    uref = np.array([0.0, 0.0]) if i % 2 else np.array([rng.uniform(-1, 1), rng.uniform(-0.3, 0.3)])
    gamma, kv = 1.0, (1.0 if i % 3 else 0.5)
    solvers.LOG.clear()
    u = nr["single_obstacle_CBF1"](s=s, u_ref=uref.copy(), c_obs=c_obs, c_obs_dot=c_dot, a=r, b=r, gamma=gamma, kv=kv)
    log = solvers.LOG[-1]
    rad_in[i] = [*s, *uref, *c_obs, *c_dot, r, gamma, kv]
    rad_out[i] = [u[0], u[1], log["A"][0, 0], log["A"][0, 1], log["b"][0], log["mask"], log["status"]]
out["radial_in"] = rad_in      # s[4] uref[2] c[2] cdot[2] r gamma kv
out["radial_out"] = rad_out    # a delta | A0 A1 b | mask status

dst = os.path.join(HERE, "reference_vectors.npz")
np.savez_compressed(dst, **out)
print("wrote", dst, {k: v.shape for k, v in out.items()})
