"""Generate tests/golden/reference_vectors_lane_sadbm.npz by EXECUTING THE REFERENCE'S OWN CODE
(same route and shims as gen_reference_vectors.py; run once in the build container):

* ``test_scripts/stanley_controller_ellipse.py``: ``CBF_lane``, ``CBF_lane_sqrt``, ``CBF_lane_cf``,
  ``CBF_lane_cf_sqrt`` (lines 416-579) pulled out by AST, driven with the reference's own
  ``PolynomialLaneCurve`` (``test_scripts/lane_cbf_test.py:10-157``, real scipy Newton-CG);
  rows are captured inside the shimmed ``solvers.cp``;
* ``cbf/cbf.py``: ``SADBM_CBF_2DS`` (lines 300-437) with a FIXED ``dt`` over sequences of ticks
  (the class is stateful: beta and the last beta_ref carry over), collision-cone obstacles;
* ``cbf/obstacles.py``: ``PolyLane.fit_polynomial_curve`` (lines 715-773, real scipy ``curve_fit``) with and
  without fixed points / ``fixed_pts_idx``, degrees 1-5.
"""
import ast
import contextlib
import io
import math
import os
import sys
import types

import numpy as np
import scipy
import scipy.optimize  # noqa: F401

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(0, REF)

if not hasattr(np, "math"):
    np.math = math          # lane_cbf_test.py:29 uses np.math.factorial (removed in numpy 2)

from cvxopt import matrix, solvers, sqrt  # noqa: E402  (the shim)
import euclid as euc  # noqa: E402  (the shim)
from euclid import Point2, Vector2  # noqa: E402

from cbf.cbf import SADBM_CBF_2DS  # noqa: E402
from cbf.obstacles import CollisionCone2D, PolyLane  # noqa: E402


def extract(path, names, consts, extra):
    src = open(path).read()
    tree = ast.parse(src)
    keep = []
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            keep.append(node)
        elif isinstance(node, ast.Assign) and all(isinstance(t, ast.Name) and t.id in consts for t in node.targets):
            keep.append(node)
    mod = ast.Module(body=keep, type_ignores=[])
    ns = dict(np=np, matrix=matrix, solvers=solvers, sqrt=sqrt, Point2=Point2, Vector2=Vector2)
    ns.update(extra)
    exec(compile(mod, path, "exec"), ns)
    return ns


out = {}
rng = np.random.default_rng(20261018)

# ---------------------------------------------------------------- 1. lane barrier variants (driver functions)
lns = extract(os.path.join(REF, "test_scripts", "lane_cbf_test.py"), {"PolynomialLaneCurve"}, set(),
              dict(euc=euc, sci=scipy, cp=None))
lutils = types.SimpleNamespace(PolynomialLaneCurve=lns["PolynomialLaneCurve"])
SCE = os.path.join(REF, "test_scripts", "stanley_controller_ellipse.py")
ns = extract(SCE, {"CBF_lane", "CBF_lane_sqrt", "CBF_lane_cf", "CBF_lane_cf_sqrt"},
             {"lr", "lf", "L", "ZERO_TOL"}, dict(lutils=lutils))

NL = 48
lane_in = np.zeros((NL, 4 + 4 + 2 + 2))      # s[4], c0..c3, u_des[2], buffer, alpha
lane_rows = np.zeros((NL, 2, 3))             # [squared, sqrt] x (A0, A1, b)
lane_u = np.zeros((NL, 4, 2))                # CBF_lane, CBF_lane_sqrt, CBF_lane_cf, CBF_lane_cf_sqrt
lane_xc = np.zeros((NL, 4))
lane_psi = np.zeros((NL, 2))
for i in range(NL):
    s = np.array([rng.uniform(-10, 60), rng.uniform(-4, 4), rng.uniform(-0.6, 0.6), rng.uniform(3, 12)])
    co = np.array([rng.uniform(2.0, 5.0) * (1 if i % 2 else -1), rng.uniform(-0.05, 0.05),
                   rng.uniform(-0.002, 0.002), rng.uniform(-2e-5, 2e-5)])
    if i % 5 == 0:
        co[2:] = 0.0                                              # straight lane
    u_des = np.array([rng.uniform(-1, 1), rng.uniform(-0.3, 0.3)])
    buffer = [1.5, 1.0, 2.5][i % 3]
    alpha = [1.0, 0.5, 2.0][i % 3]
    lane = lutils.PolynomialLaneCurve(co)
    lane_in[i] = [*s, *co, *u_des, buffer, alpha]
    for j, fn in enumerate(("CBF_lane", "CBF_lane_sqrt")):
        solvers.LOG.clear()
        u, xc = ns[fn](s.copy(), lane, u_des.copy(), buffer, alpha)
        log = solvers.LOG[-1]
        lane_rows[i, j] = [log["A"][0, 0], log["A"][0, 1], log["b"][0]]
        lane_u[i, j] = [u[0], u[1]]
        lane_xc[i, j] = float(np.ravel(xc)[0])
    for j, fn in enumerate(("CBF_lane_cf", "CBF_lane_cf_sqrt")):
        u, xc, psi = ns[fn](s.copy(), lane, u_des.copy(), buffer, alpha)
        lane_u[i, 2 + j] = [u[0], u[1]]
        lane_xc[i, 2 + j] = float(np.ravel(xc)[0])
        lane_psi[i, j] = float(np.ravel(psi)[0])
out["lanev_in"] = lane_in
out["lanev_rows"] = lane_rows
out["lanev_u"] = lane_u
out["lanev_xc"] = lane_xc
out["lanev_psi"] = lane_psi

# ---------------------------------------------------------------- 2. SADBM_CBF_2DS (class), fixed dt, sequences of ticks
NS, NT, MAXM = 16, 6, 3
sad_cfg = np.zeros((NS, 4 + 4))                    # alpha, dt, lr, lf, R (row-major)
sad_m = np.zeros(NS, dtype=np.int64)
sad_cone = np.zeros((NS, NT, MAXM, 5))             # cx, cy, theta_o, v_o, a (incl. buffer) per tick
sad_in = np.zeros((NS, NT, 4 + 2 + 2))             # s[4], u_ref (a, delta), beta before, beta_ref_last before
sad_rows = np.full((NS, NT, MAXM, 3), np.nan)      # A0 A1 b assembled by cbf.py
sad_out = np.zeros((NS, NT, 6))                    # u0 (a), u1 (delta), beta after, beta_ref_dot (x0[1]), mask, status
for i in range(NS):
    alpha = [1.0, 0.5, 2.0][i % 3]
    dt = [0.001, 0.01, 0.05][i % 3]
    Rm = np.eye(2) if i % 2 else np.array([[1.0, 0.2], [0.2, 3.0]])
    m = 1 + i % MAXM
    ctl = SADBM_CBF_2DS(alpha=alpha, dt=dt)
    ctl.set_model_params(lr=1.45, lf=1.45)
    ctl.set_qp_cost_weight(Rm)
    s = np.array([rng.uniform(-20, 20), rng.uniform(-5, 25), rng.uniform(-1.0, 1.0), rng.uniform(3, 12)])
    obs = []
    for j in range(m):
        so = np.array([s[0] + rng.uniform(6, 35) * math.cos(s[2]), s[1] + rng.uniform(6, 35) * math.sin(s[2]) + rng.uniform(-6, 6),
                       rng.uniform(-3, 3), rng.uniform(0, 6)])
        a = rng.uniform(1, 4)
        ob = CollisionCone2D(a, s, so)
        ctl.obstacle_list2d[j] = ob
        obs.append((ob, so, a + 1.5))
    sad_cfg[i] = [alpha, dt, 1.45, 1.45, *Rm.ravel()]
    sad_m[i] = m
    for t in range(NT):
        uref = np.array([rng.uniform(-2, 2), rng.uniform(-0.35, 0.35)])
        for j, (ob, so, a) in enumerate(obs):
            so[0] += so[3] * math.cos(so[2]) * 0.1
            so[1] += so[3] * math.sin(so[2]) * 0.1
            ob.update(s_obs=so)
            sad_cone[i, t, j] = [so[0], so[1], so[2], so[3], a]
        ctl.update_state(s=s)
        sad_in[i, t] = [*s, *uref, ctl._beta, ctl.beta_ref_last]
        solvers.LOG.clear()
        with contextlib.redirect_stdout(io.StringIO()):
            sol, u = ctl.solve_cbf(uref.copy(), return_solver=True)
        log = solvers.LOG[-1]
        sad_rows[i, t, :m, 0:2] = log["A"]
        sad_rows[i, t, :m, 2] = log["b"]
        sad_out[i, t] = [u[0], u[1], ctl._beta, log["x0"][1], log["mask"], log["status"]]
        # the plant is not part of the class: move the state the way update_com would (any motion will do)
        beta = ctl._beta
        s = np.array([s[0] + s[3] * math.cos(s[2] + beta) * 0.1, s[1] + s[3] * math.sin(s[2] + beta) * 0.1,
                      s[2] + s[3] * math.sin(beta) / 1.45 * 0.1, max(0.5, s[3] + float(u[0]) * 0.1)])
out["sadbm_cfg"] = sad_cfg
out["sadbm_m"] = sad_m
out["sadbm_cone"] = sad_cone
out["sadbm_in"] = sad_in
out["sadbm_rows"] = sad_rows
out["sadbm_out"] = sad_out

# ---------------------------------------------------------------- 3. PolyLane.fit_polynomial_curve (class)
NF_, KMAX = 30, 40
fit_x = np.full((NF_, KMAX), np.nan)
fit_y = np.full((NF_, KMAX), np.nan)
fit_sigma = np.full((NF_, KMAX), np.nan)         # the sigma the reference ends up using per point
fit_k = np.zeros(NF_, dtype=np.int64)
fit_n = np.zeros(NF_, dtype=np.int64)
fit_c = np.zeros((NF_, 6))
for i in range(NF_):
    n = 1 + i % 5
    K = int(rng.integers(n + 3, 28))
    x0 = rng.uniform(-20, 40)
    x = np.sort(x0 + rng.uniform(0, 60, K))
    true = np.array([rng.uniform(-4, 4), rng.uniform(-0.3, 0.3), rng.uniform(-4e-3, 4e-3), rng.uniform(-4e-5, 4e-5),
                     rng.uniform(-2e-7, 2e-7), rng.uniform(-1e-9, 1e-9)])[: n + 1]
    y = sum(true[j] * x ** j for j in range(n + 1)) + rng.normal(0, 0.05, K)
    kw = dict(n=n)
    sig = np.full(K, 10.0)
    xs, ys = x, y
    if i % 3 == 1:                                   # appended fixed points (obstacles.py:749-756)
        xf = np.array([x[0] - 2.0, x[-1] + 2.0]); yf = sum(true[j] * xf ** j for j in range(n + 1))
        kw.update(x_fixed_pts=xf, y_fixed_pts=yf, alpha=0.01)
        xs, ys, sig = np.append(x, xf), np.append(y, yf), np.append(sig, [0.01, 0.01])
    if i % 3 == 2:                                   # pinned existing points (obstacles.py:758-759)
        idx = np.array([0, K // 2])
        kw.update(fixed_pts_idx=idx, alpha=0.05)
        sig[idx] = 0.05
    if i % 4 == 0:
        sg = rng.uniform(0.5, 5.0, K)
        kw.update(sigma=sg.copy())
        sig[:K] = sg
        if i % 3 == 2:
            sig[idx] = 0.05
    lane = PolyLane.fit_polynomial_curve(x, y, **kw)
    co = np.asarray(lane.coeffs, dtype=np.float64).ravel()
    fit_k[i] = xs.size; fit_n[i] = n
    fit_x[i, : xs.size] = xs; fit_y[i, : xs.size] = ys; fit_sigma[i, : xs.size] = sig
    fit_c[i, : n + 1] = co
out["fit_x"] = fit_x
out["fit_y"] = fit_y
out["fit_sigma"] = fit_sigma
out["fit_k"] = fit_k
out["fit_n"] = fit_n
out["fit_c"] = fit_c

dst = os.path.join(HERE, "reference_vectors_lane_sadbm.npz")
np.savez_compressed(dst, **out)
print("wrote", dst, {k: v.shape for k, v in out.items()})
