"""CPU oracle for the CBF-QP hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this module.  The product path (``sccav_cbf_b200``) never does:
it fails loudly when the CUDA library is missing.

What this is
------------
A scalar, plain-Python/numpy *restatement* of the arithmetic of the reference
(Safety-Critical-Control-WIRIN/sccav_cbf) on the path
``barrier evaluation -> 2-variable CBF-QP -> closed-loop Stanley / bicycle rollout``.
Every function cites the reference ``file:line`` it follows (paths relative to the reference
root).  Formula order is kept literal (no algebraic re-association) because the integer
bookkeeping (waypoint index, step count) is compared bit-for-bit.

Where the reference's arithmetic lives in an un-vendored, un-pinned third-party package:

* ``cvxopt.solvers.cp`` (interior point; no version pinned anywhere in the reference; call
  sites cbf/cbf.py:107,213 and test_scripts/stanley_controller_ellipse.py:238,275,318,414) is
  restated as the *exact* optimum of the strictly convex problem it approximates,
  ``min (u-r)^T R (u-r)  s.t.  A u >= b`` with two variables: enumerate the empty working set,
  every single row and every pair of rows, and keep the candidate that satisfies the KKT
  conditions (unique by strict convexity).  cvxopt stops at abstol 1e-7 / reltol 1e-6, so its
  answers sit ~1e-5 (relative) from this optimum on active steps.
* ``scipy.optimize.minimize(method='Newton-CG')`` (cbf/obstacles.py:675-677) is restated as a
  safeguarded Newton iteration with a fixed stopping rule (see ``lane_closest_x``).
* ``euclid`` vectors are plain floats here.

Parity pinning status
---------------------
* Collision-cone + DBM rows + delta<->beta + ``update_com`` + Stanley index logic + loop
  bookkeeping: PINNED by the reference's own ``test_scripts/beta_vs_time.mat`` (277 samples,
  tests/golden/beta_vs_time.json) to 1e-3 deg (the IPM tolerance of the run that wrote it).
* Ellipse / cone / lane partials and the DBM/KBM row assembly: PINNED bit-for-bit by vectors
  generated from the reference's own ``cbf/obstacles.py`` / ``cbf/cbf.py`` executed in the build
  container with shimmed ``euclid``/``cvxopt`` (tests/golden/gen_reference_rows.py).
* Lane closest point: pinned against scipy's Newton-CG (installed here) to 1e-7.
* Radial-dynamic closed loop (rdo.py uses unseeded RNG): parity unpinned beyond the formulas.
"""
from __future__ import annotations

import bisect
import math

import numpy as np

# --------------------------------------------------------------------------------------
# constants (cbf/utils.py:27, test_scripts/stanley_controller_ellipse.py:52-62)
# --------------------------------------------------------------------------------------
ZERO_TOL = 1e-3
PI = float(np.pi)

# obstacle slot types / field layout: mirrors include/sccav_cbf.h (the C-ABI)
SLOT_ELLIPSE = 0   # f: cx, cy, a, b, theta, vx, vy, -
SLOT_CONE = 1      # f: cx, cy, theta_o, v_o, a, beta, -, -
SLOT_LANE = 2      # f: buffer, c0, c1, c2, c3, c4, c5, -
SLOT_RADIAL = 3    # f: cx, cy, a, b, kv, vx, vy, -
SLOT_DISTANCE = 4  # f: cx, cy, Ds, -, -, -, -, -
SLOT_ELLIPSE_PREP = 5  # f: cx, cy, m00, m01, m10, m11, wx, wy   (prepare_ellipse below)
SLOT_LANE_SQRT = 6  # f: as LANE; h = sqrt(d^2) - buffer (CBF_lane_sqrt, stanley_controller_ellipse.py:465-512)
SLOT_TYPE_MASK = 0x3F
SLOT_STATIC = 0x40  # flag: velocity fields are not read, h_t = 0
FLAG_PREPARED_ROWS = 1  # sccav_params.flags (GPU rollout only; the oracle's arithmetic is always canonical)
FLAG_QP_ENUMERATE = 2   # GPU only: no QP shortcut (the oracle always enumerates)
FLAG_FUSED_STEER = 4    # GPU rollout only: beta = clamp(beta*) (the oracle always runs beta* -> delta -> clip -> beta)
FLAG_BETA_IO = 8        # GPU filter step only: beta in / beta out (the oracle keeps the reference's delta interface; tests convert)
FLAG_SEEKER_DIRECT = 16  # GPU rollout only: seeker direction as the normalised offset (the oracle keeps sincos(atan2))
NFIELD = 8

MODEL_DBM = 0      # DBM_CBF_2DS  (cbf/cbf.py:112-220)
MODEL_KBM = 1      # KBM_VC_CBF2D (cbf/cbf.py:33-110)
MODEL_NONE = 2     # rollout only: USE_CBF = False -> State.update (stanley_controller_ellipse.py:828)
MODEL_DUM = 3      # DUM_CBF_2DS  (cbf/cbf.py:222-298), filter step only
MODEL_SADBM = 4    # SADBM_CBF_2DS (cbf/cbf.py:300-437) with a fixed dt, filter step only (stateful: beta, last beta_ref)

STATUS_INACTIVE = 0    # u == u_ref (no row active)
STATUS_ACTIVE = 1      # optimum with 1 or 2 active rows
STATUS_INFEASIBLE = 2  # no KKT point: least-violation candidate returned

QP_FEAS_EPS = 1e-12    # relative feasibility tolerance (fp64 spec)
QP_PAR_EPS = 1e-12     # relative parallel-row tolerance (fp64 spec)
QP_TIE_EPS = 1e-9      # infeasible fallback: a later candidate must beat the incumbent by this relative margin
LANE_MAX_IT = 50
LANE_LS_MAX = 30
LANE_XTOL = 1e-12
LANE_DTOL = 8 * 2.0 ** -52   # a step is accepted if it does not increase D beyond its rounding noise (8 ulp)


def normalize_angle(angle):
    """cbf/utils.py:93-106 == stanley_controller_ellipse.py:172-185 (strict inequalities).
    The reference's loops never terminate for |angle| >~ 1e16 or inf; beyond 1e4 (a diverged
    scenario) whole turns are removed first (same rule in oracle.c and the CUDA path)."""
    if not (abs(angle) <= 1e4):
        angle = angle - (2.0 * PI) * float(np.rint(angle / (2.0 * PI)))
    while angle > PI:
        angle -= 2.0 * PI
    while angle < -PI:
        angle += 2.0 * PI
    return angle


def saturation(x, x_min, x_max):
    """cbf/utils.py:108-114."""
    if x > x_max:
        return x_max
    elif x < x_min:
        return x_min
    return x


# --------------------------------------------------------------------------------------
# barriers: return (h, h_x, h_y, h_theta, h_v, h_t)
# --------------------------------------------------------------------------------------
def ellipse_partials(x, y, cx, cy, a, b, theta, vx, vy):
    """Ellipse2D.evaluate/dx/dy/dtheta/dv/dt -- cbf/obstacles.py:183-236,304-317.

    ``a``/``b`` already include the buffer (obstacles.py:159-160).  ``h_t`` ignores the rotation
    exactly as the reference does (obstacles.py:316).  dtheta = dv = 0 (intended
    semantics of obstacles.py:232-236,304-308).
    """
    dx = x - cx
    dy = y - cy
    ct = float(np.cos(theta))
    st = float(np.sin(theta))
    p = dx * ct + dy * st
    q = -dx * st + dy * ct
    h = (p / a) ** 2 + (q / b) ** 2 - 1
    h_x = (2 * ct / (a ** 2)) * p + (-2 * st / (b ** 2)) * q
    h_y = (2 * st / (a ** 2)) * p + (2 * ct / (b ** 2)) * q
    h_t = -2 * ((dx / a ** 2) * vx + (dy / b ** 2) * vy)
    return h, h_x, h_y, 0.0, 0.0, h_t


def radial_partials(x, y, v, cx, cy, a, b, kv, vx, vy):
    """single_obstacle_CBF1 -- test_scripts/radial_dynamic_obstacles.py:391-405."""
    h = ((x - cx) / a) ** 2 + ((y - cy) / b) ** 2 - 1 - (kv * v / (1 + v))
    h_x = 2 * (x - cx) / (a ** 2)
    h_y = 2 * (y - cy) / (b ** 2)
    h_v = -kv / ((1 + v) ** 2)
    h_t = -2 * (((x - cx) / (a ** 2)) * vx + ((y - cy) / (b ** 2)) * vy)
    return h, h_x, h_y, 0.0, h_v, h_t


def distance_partials(x, y, cx, cy, Ds):
    """D_CBF -- stanley_controller_ellipse.py:251-255 (factor 2 in the partials is the reference's)."""
    h = math.sqrt((x - cx) ** 2 + (y - cy) ** 2) - Ds
    h_x = 2 * (x - cx) / (h + Ds)
    h_y = 2 * (y - cy) / (h + Ds)
    return h, h_x, h_y, 0.0, 0.0, 0.0


def cone_partials(x, y, th, v, cx, cy, th_o, v_o, a, beta=0.0):
    """CollisionCone2D.update/evaluate/dx/dy/dv/dtheta/dt -- cbf/obstacles.py:468-502,401-458.

    ``a`` already includes the buffer (obstacles.py:357).  ZERO_TOL enters additively, twice in
    some denominators (obstacles.py:428,435,456) -- kept literally.
    """
    s_vx = v * float(np.cos(th))
    s_vy = v * float(np.sin(th))
    o_vx = v_o * float(np.cos(th_o + beta))
    o_vy = v_o * float(np.sin(th_o + beta))
    prx = x - cx
    pry = y - cy
    vrx = s_vx - o_vx
    vry = s_vy - o_vy
    dist = math.sqrt(prx * prx + pry * pry)       # vec_norm, cbf/utils.py:123
    vrn = math.sqrt(vrx * vrx + vry * vry)
    if abs(dist) > abs(a):
        cb = math.sqrt(dist ** 2 - a ** 2) + ZERO_TOL
    else:
        cb = ZERO_TOL
    if dist > ZERO_TOL:
        cos_phi = cb / dist
    else:
        cos_phi = 0.0
    h = (prx * vrx + pry * vry) + (dist * vrn * cos_phi)
    h_x = (s_vx - o_vx) + vrn * (x - cx) / (cb + ZERO_TOL)
    h_y = (s_vy - o_vy) + vrn * (y - cy) / (cb + ZERO_TOL)
    cb_ = float(np.cos(th + beta))
    sb_ = float(np.sin(th + beta))
    h_v = ((x - cx) * cb_ + (y - cy) * sb_) + \
        ((s_vx - o_vx) * cb_ + (s_vy - o_vy) * sb_) * cb / (vrn + ZERO_TOL)
    h_th = (-(x - cx) * s_vy + (y - cy) * s_vx) + \
        (-(s_vx - o_vx) * s_vy + (s_vy - o_vy) * s_vx) * cb / (vrn + ZERO_TOL)
    h_t = (-(s_vx - o_vx) * o_vx - (s_vy - o_vy) * o_vy) + \
        (-vrn * ((x - cx) * o_vx + (y - cy) * o_vy) / (cb + ZERO_TOL))
    return h, h_x, h_y, h_th, h_v, h_t


def _poly3(c, x):
    """Value, first and second derivative of sum c_i x^i by Horner, the order
    numpy.polynomial.Polynomial / polyder use (cbf/obstacles.py:589-592)."""
    n = len(c)
    g = 0.0
    for i in range(n - 1, -1, -1):
        g = c[i] + g * x
    dg = 0.0
    for i in range(n - 1, 0, -1):
        dg = (i * c[i]) + dg * x
    ddg = 0.0
    for i in range(n - 1, 1, -1):
        ddg = ((i - 1) * (i * c[i])) + ddg * x
    return g, dg, ddg


def lane_closest_x(c, px, py):
    """Closest abscissa on y = g(x) to (px, py), from x0 = px.

    Restates PolyLane.get_shortest_distance_x (cbf/obstacles.py:641-679): the reference calls
    scipy Newton-CG (xtol 1e-8) on D(x) = (x-px)^2 + (g-py)^2 with the exact gradient and a
    Hessian carrying a typo (obstacles.py:673).  The typo only changes the path,
    not the fixed point, so this oracle runs a safeguarded Newton on the half-gradient
    ``(x-px) + (g-py) g'`` with the correct half-Hessian ``1 + g'^2 + (g-py) g''`` and a
    backtracking line search on D (a step may raise D by its rounding noise, 8 ulp: near the minimum
    the full Newton step changes D by less than that, and rejecting it there turns the quadratic
    convergence into a bisection of ~20 iterations); stops when the accepted step is <= 1e-12 (1+|x|).
    """
    x = px
    for _ in range(LANE_MAX_IT):
        g, dg, ddg = _poly3(c, x)
        ex = x - px
        ey = g - py
        grad = ex + ey * dg
        hess = (1 + dg * dg) + ey * ddg
        if hess > 0:
            step = -grad / hess
        else:
            step = -grad
        D0 = ex * ex + ey * ey
        Dacc = D0 + LANE_DTOL * D0
        t = 1.0
        ok = False
        xn = x
        for _ls in range(LANE_LS_MAX):
            xn = x + t * step
            gn, _, _ = _poly3(c, xn)
            Dn = (xn - px) * (xn - px) + (gn - py) * (gn - py)
            if Dn <= Dacc:
                ok = True
                break
            t = t * 0.5
        if not ok:
            break
        dxn = abs(xn - x)
        lim = LANE_XTOL * (1 + abs(x))
        x = xn
        if dxn <= lim:
            break
    return x


def lane_partials(x, y, c, buffer):
    """PolyLane.update/evaluate/dx/dy -- cbf/obstacles.py:620-636,607-612,681-689.
    ``buffer`` is NOT squared in h (obstacles.py:611)."""
    cx = lane_closest_x(c, x, y)
    g, dg, ddg = _poly3(c, cx)
    eta = 1 + dg * ddg + dg ** 2 - y * ddg
    if abs(eta) < ZERO_TOL:
        eta = ZERO_TOL
    h = (cx - x) ** 2 + (g - y) ** 2 - buffer
    h_x = (2 / eta) * ((x - cx) * (eta - 1) - (y - g) * dg)
    h_y = (2 / eta) * (-(x - cx) * dg + (y - g) * (eta - dg ** 2))
    return h, h_x, h_y, 0.0, 0.0, 0.0


def lane_sqrt_partials(x, y, c, buffer):
    """CBF_lane_sqrt / CBF_lane_cf_sqrt -- test_scripts/stanley_controller_ellipse.py:489-492,573-576:
    h = sqrt(d^2) - buffer and the squared-distance gradient divided by 2 (h + buffer)."""
    cx = lane_closest_x(c, x, y)
    g, dg, ddg = _poly3(c, cx)
    eta = 1 + dg * ddg + dg ** 2 - y * ddg
    if abs(eta) < ZERO_TOL:
        eta = ZERO_TOL
    h = math.sqrt((cx - x) ** 2 + (g - y) ** 2) - buffer
    h_x = ((2 / eta) * ((x - cx) * (eta - 1) - (y - g) * dg)) / (2 * (h + buffer))
    h_y = ((2 / eta) * (-(x - cx) * dg + (y - g) * (eta - dg ** 2))) / (2 * (h + buffer))
    return h, h_x, h_y, 0.0, 0.0, 0.0


def fit_polynomial(x_pts, y_pts, n=3, sigma=None):
    """PolyLane.fit_polynomial_curve -- cbf/obstacles.py:715-773 (scipy ``curve_fit`` of a polynomial with
    per-point ``sigma``; fixed points are ordinary points with a small sigma, :749-756).  The model is linear
    in its coefficients: the optimum curve_fit iterates to is the weighted least-squares solution.  It is
    computed in the centred, scaled abscissa t = (x - mid) / half_range by an orthogonal factorisation (numpy
    lstsq; the raw Vandermonde matrix of x ~ 50 m at degree 5 is too ill-conditioned for that), then expanded
    to powers of x.  Returns c0..cn."""
    x = np.asarray(x_pts, dtype=np.float64).ravel()
    y = np.asarray(y_pts, dtype=np.float64).ravel()
    sg = np.full_like(x, 10.0) if sigma is None else np.asarray(sigma, dtype=np.float64).ravel()
    mid = 0.5 * (x.min() + x.max())
    half = 0.5 * (x.max() - x.min())
    t = (x - mid) / half
    V = np.vander(t, n + 1, increasing=True) / sg[:, None]
    a, *_ = np.linalg.lstsq(V, y / sg, rcond=1e-15)
    # p(t) = sum a_j t^j, t = x / half - mid / half: Horner in polynomials of x
    c = np.zeros(n + 1)
    for j in range(n, -1, -1):
        nc = np.zeros(n + 1)
        nc[1:] = c[:-1] / half
        nc -= c * (mid / half)
        nc[0] += a[j]
        c = nc
    return c


def prepare_ellipse(f):
    """Ingest-time half of Ellipse2D (include/sccav_cbf.h, ELLIPSE_PREP): canonical fields
    (cx, cy, a, b, theta, vx, vy, -) -> (cx, cy, cos/a, sin/a, -sin/b, cos/b, vx/a^2, vy/b^2)."""
    cx, cy, a, b, th, vx, vy = (float(v) for v in f[:7])
    ct = float(np.cos(th))
    st = float(np.sin(th))
    return [cx, cy, ct / a, st / a, -st / b, ct / b, vx / (a * a), vy / (b * b)]


def ellipse_prep_partials(x, y, cx, cy, m00, m01, m10, m11, wx, wy):
    """Per-solve half: with d = (x - cx, y - cy) and (pa, qb) = M d,  h = pa^2 + qb^2 - 1,
    grad h = 2 M^T (pa, qb), h_t = -2 (dx wx + dy wy) -- the functions of cbf/obstacles.py:193,218,229,316
    regrouped (equal to ellipse_partials up to a few ulp)."""
    dx = x - cx
    dy = y - cy
    pa = m00 * dx + m01 * dy
    qb = m10 * dx + m11 * dy
    h = (pa * pa + qb * qb) - 1.0
    h_x = 2.0 * (m00 * pa + m10 * qb)
    h_y = 2.0 * (m01 * pa + m11 * qb)
    h_t = -2.0 * (dx * wx + dy * wy)
    return h, h_x, h_y, 0.0, 0.0, h_t


def slot_partials(slot_desc, f, s):
    """Dispatch on the slot type; ``f`` = the 8 slot fields, ``s`` = (x, y, theta, v)."""
    x, y, th, v = s
    slot_type = int(slot_desc) & SLOT_TYPE_MASK
    static = bool(int(slot_desc) & SLOT_STATIC)
    if slot_type == SLOT_ELLIPSE:
        return ellipse_partials(x, y, f[0], f[1], f[2], f[3], f[4], 0.0 if static else f[5], 0.0 if static else f[6])
    if slot_type == SLOT_ELLIPSE_PREP:
        return ellipse_prep_partials(x, y, f[0], f[1], f[2], f[3], f[4], f[5], 0.0 if static else f[6], 0.0 if static else f[7])
    if slot_type == SLOT_CONE:
        return cone_partials(x, y, th, v, f[0], f[1], f[2], f[3], f[4], f[5])
    if slot_type == SLOT_LANE:
        return lane_partials(x, y, [f[1], f[2], f[3], f[4], f[5], f[6]], f[0])
    if slot_type == SLOT_LANE_SQRT:
        return lane_sqrt_partials(x, y, [f[1], f[2], f[3], f[4], f[5], f[6]], f[0])
    if slot_type == SLOT_RADIAL:
        return radial_partials(x, y, v, f[0], f[1], f[2], f[3], f[4], f[5], f[6])
    if slot_type == SLOT_DISTANCE:
        return distance_partials(x, y, f[0], f[1], f[2])
    raise ValueError("unknown slot type %r" % (slot_type,))


# --------------------------------------------------------------------------------------
# obstacle ingest from bounding boxes (the step BEFORE the path in the CARLA deployment)
# --------------------------------------------------------------------------------------
INGEST_UPDATE = 0
INGEST_REBUILD = 1


def box_to_fields(obs_type, create, buffer, box, old=None):
    """Slot fields of one obstacle from its bounding box (extent.x, extent.y, location.x, location.y, yaw,
    velocity).  create: <obstacle>.from_bounding_box (cbf/obstacles.py:319-331 ellipse, :532-543 cone; the
    buffer is added to a / b); else <obstacle>.update_by_bounding_box on the fields ``old``
    (:294-302 ellipse -- a, b, centre, theta replaced, buffer NOT re-applied, velocity kept; :512-530 cone --
    a = hypot(extent), s_obs = [x, y, 0.0, velocity], beta kept)."""
    ex, ey, lx, ly, yaw, sp = (float(v) for v in box)
    f = [0.0] * 8 if create else [float(v) for v in old]
    if obs_type == SLOT_ELLIPSE:
        f[0], f[1], f[4] = lx, ly, yaw
        f[2] = ex + buffer if create else ex
        f[3] = ey + buffer if create else ey
    elif obs_type == SLOT_CONE:
        a = float(np.hypot(ex, ey))
        f[0], f[1], f[2], f[3] = lx, ly, 0.0, sp
        f[4] = a + buffer if create else a
    else:
        raise ValueError("update_by_bounding_box makes Ellipse2D or CollisionCone2D obstacles")
    return f


def ingest_boxes(obs_type, mode, buffer, M, box_ids, boxes, slot_ids, fields, count):
    """ObstacleList2D.update_by_bounding_box (cbf/obstacles.py:833-858) for ONE vehicle whose list lives in
    M slots: ``slot_ids[m]`` / ``fields[m]`` for m < count.  Boxes: ``box_ids[k]`` (< 0 = none), ``boxes[k]``.
    The reference walks the boxes (update the ids it holds, append the new ones -- :839-846), then pops the
    ids that are gone (:851-853): surviving entries keep their order, new ones follow in box order.  With
    INGEST_REBUILD the list is rebuilt from the boxes alone (multi_obstacle_CBF_local_with_lanes.py:918-928).
    Only M entries fit: the last new ids are dropped (counted).  Returns (slot_ids, fields, count, dropped)."""
    mapping = {}
    if mode == INGEST_UPDATE:
        for j in range(min(max(int(count), 0), M)):
            if slot_ids[j] >= 0:
                mapping[int(slot_ids[j])] = [float(v) for v in fields[j]]
    bbox = {}
    for k, key in enumerate(box_ids):
        if key >= 0 and int(key) not in bbox:
            bbox[int(key)] = boxes[k]
    for key, box in bbox.items():                                          # obstacles.py:838-846
        if key in mapping:
            mapping[key] = box_to_fields(obs_type, False, buffer, box, mapping[key])
        else:
            mapping[key] = box_to_fields(obs_type, True, buffer, box)
    for key in list(mapping.keys()):                                       # obstacles.py:851-853
        if key not in bbox:
            mapping.pop(key)
    keys = list(mapping.keys())
    dropped = max(0, len(keys) - M)
    keys = keys[:M]
    out_ids = [-1] * M
    out_f = [[float(v) for v in fields[m]] for m in range(M)]
    for m, key in enumerate(keys):
        out_ids[m] = key
        out_f[m] = mapping[key]
    return out_ids, out_f, len(keys), dropped


# --------------------------------------------------------------------------------------
# actuator shaping (the step AFTER the path in the CARLA deployment)
# --------------------------------------------------------------------------------------
def actuator_shaping(u_a, delta, throttle_previous, brake_previous, max_steer=1.0, rate=0.1, reset_brake=False):
    """multi_obstacle_CBF_local_with_lanes.py:955-976, literally.  ``brake_previous`` doubles as the driver's
    ``brake`` variable of the last tick: the throttle branch does not touch it (:958-962), so it keeps its
    value unless ``reset_brake``.  Returns (throttle, brake, steer)."""
    brake = brake_previous
    if u_a > 0:
        throttle = float(np.tanh(u_a))
        throttle = max(0.0, min(1.0, throttle))
        if throttle - throttle_previous > rate:
            throttle = throttle_previous + rate
        if reset_brake:
            brake = 0.0
    else:
        throttle = 0
        brake = -float(np.tanh(u_a))
        brake = max(0.0, min(1.0, brake))
        if brake - brake_previous > rate:
            brake = brake_previous + rate
    if delta > 0:
        delta = max(0.0, min(delta, max_steer))
    else:
        delta = max(-max_steer, min(delta, 0.0))
    return float(throttle), float(brake), float(delta)


# --------------------------------------------------------------------------------------
# row assembly: constraint  A0*u0 + A1*u1 >= b
# --------------------------------------------------------------------------------------
def dbm_row(part, th, v, alpha, lr):
    """DBM_CBF_2DS.gc/fc + F -- cbf/cbf.py:159-164,200-207.
    g_c columns [0,0,0,1], [-v sin, v cos, v/lr, 0]; f_c = [v cos, v sin, 0, 0];
    constraint Lf + Lg.u + alpha h + h_t >= 0."""
    h, h_x, h_y, h_th, h_v, h_t = part
    c = float(np.cos(th))
    s = float(np.sin(th))
    A0 = h_v
    A1 = (h_x * (-v * s) + h_y * (v * c)) + h_th * (v / lr)
    Lf = h_x * (v * c) + h_y * (v * s)
    b = -((Lf + alpha * h) + h_t)
    return A0, A1, b


def dum_row(part, th, v, alpha):
    """DUM_CBF_2DS.gc/fc + F -- cbf/cbf.py:237-245,277-286.  g_c columns [0,0,0,1], [0,0,1,0] (u = [v_dot,
    theta_dot]); f_c = [v cos, v sin, 0, 0] (the reference declares this 4-element list as 5 x 1, which cvxopt
    rejects -- the 4-vector it lists is meant)."""
    h, h_x, h_y, h_th, h_v, h_t = part
    c = float(np.cos(th))
    s = float(np.sin(th))
    A0 = h_v
    A1 = h_th
    Lf = h_x * (v * c) + h_y * (v * s)
    b = -((Lf + alpha * h) + h_t)
    return A0, A1, b


def kbm_row(part, th, alpha):
    """KBM_VC_CBF2D.solve_cbf F -- cbf/cbf.py:94-101 (== CBF(), stanley_controller_ellipse.py:226-230).
    g_c columns [cos, sin, 0], [0, 0, 1]; constraint Lg.u + alpha h >= 0."""
    h, h_x, h_y, h_th, _h_v, _h_t = part
    c = float(np.cos(th))
    s = float(np.sin(th))
    A0 = h_x * c + h_y * s
    A1 = h_th
    b = -(alpha * h)
    return A0, A1, b


def delta_to_beta(delta, lr, lf):
    """cbf/cbf.py:175."""
    return float(np.arctan2(lr * np.tan(delta), lf + lr))


def beta_to_delta(beta, lr, lf):
    """cbf/cbf.py:216."""
    return float(np.arctan2((lf + lr) * np.tan(beta), lr))


# --------------------------------------------------------------------------------------
# the 2-variable QP (exact restatement of what cvxopt.solvers.cp approximates)
# --------------------------------------------------------------------------------------
def qp2_exact(A0, A1, b, r0, r1, R, feas_eps=QP_FEAS_EPS, par_eps=QP_PAR_EPS, collect=None):
    """min (u-r)^T R (u-r) s.t. A u >= b  (cbf/cbf.py:182-213), u in R^2.

    Deterministic enumeration order: {} , singles by index, pairs lexicographic; first KKT
    point wins.  If no candidate is a KKT point (infeasible rows) the candidate with the least
    worst-row violation is returned (ties within QP_TIE_EPS keep the earlier candidate: all
    candidates on one boundary line often share their worst row exactly).
    Returns (u0, u1, active_mask, status).  If ``collect`` is a list, *every*
    KKT-satisfying candidate is appended to it (used by the tests to assert uniqueness).
    """
    m = len(b)
    R00, R01, R10, R11 = R
    det = R00 * R11 - R01 * R10
    Ri00 = R11 / det
    Ri01 = -R01 / det
    Ri10 = -R10 / det
    Ri11 = R00 / det

    def resid(k, u0, u1):
        return (A0[k] * u0 + A1[k] * u1) - b[k]

    def tol(k, u0, u1):
        return feas_eps * (abs(A0[k] * u0) + abs(A1[k] * u1) + abs(b[k]))

    def check(u0, u1, skip_a=-1, skip_b=-1):
        """(feasible?, max violation) over all rows; rows in the working set are exempt from
        the feasibility verdict but still counted in the violation measure."""
        feas = True
        worst = -math.inf
        for k in range(m):
            rk = resid(k, u0, u1)
            if -rk > worst:
                worst = -rk
            if k == skip_a or k == skip_b:
                continue
            if not (rk >= -tol(k, u0, u1)):
                feas = False
        return feas, worst

    result = None
    feas, worst = check(r0, r1)
    if feas:
        result = (r0, r1, 0, STATUS_INACTIVE)
        if collect is None:
            return result
        collect.append(result)
    fb = (worst, r0, r1, 0)

    for k in range(m):
        rk = resid(k, r0, r1)
        if not (rk < 0):
            continue
        g0 = Ri00 * A0[k] + Ri01 * A1[k]
        g1 = Ri10 * A0[k] + Ri11 * A1[k]
        den = A0[k] * g0 + A1[k] * g1
        if not (den > 0):
            continue
        t = (-rk) / den
        u0 = r0 + g0 * t
        u1 = r1 + g1 * t
        feas, worst = check(u0, u1, k)
        if feas:
            cand = (u0, u1, 1 << k, STATUS_ACTIVE)
            if result is None:
                result = cand
                if collect is None:
                    return result
            if collect is not None:
                collect.append(cand)
        if worst < fb[0] - QP_TIE_EPS * (abs(worst) + abs(fb[0])):
            fb = (worst, u0, u1, 1 << k)

    for j in range(m):
        for k in range(j + 1, m):
            t1 = A0[j] * A1[k]
            t2 = A1[j] * A0[k]
            det2 = t1 - t2
            if not (abs(det2) > par_eps * (abs(t1) + abs(t2))):
                continue
            u0 = (b[j] * A1[k] - A1[j] * b[k]) / det2
            u1 = (A0[j] * b[k] - b[j] * A0[k]) / det2
            e0 = u0 - r0
            e1 = u1 - r1
            w0 = 2 * (R00 * e0 + R01 * e1)
            w1 = 2 * (R10 * e0 + R11 * e1)
            lj = (w0 * A1[k] - A0[k] * w1) / det2
            lk = (A0[j] * w1 - w0 * A1[j]) / det2
            feas, worst = check(u0, u1, j, k)
            if feas and lj >= 0 and lk >= 0:
                cand = (u0, u1, (1 << j) | (1 << k), STATUS_ACTIVE)
                if result is None:
                    result = cand
                    if collect is None:
                        return result
                if collect is not None:
                    collect.append(cand)
            if worst < fb[0] - QP_TIE_EPS * (abs(worst) + abs(fb[0])):
                fb = (worst, u0, u1, (1 << j) | (1 << k))

    if result is not None:
        return result
    return fb[1], fb[2], fb[3], STATUS_INFEASIBLE


# --------------------------------------------------------------------------------------
# the filter operator: DBM_CBF_2DS.solve_cbf / KBM_VC_CBF2D.solve_cbf for one vehicle
# --------------------------------------------------------------------------------------
def barrier_rows(model, s, slot_types, fields, alpha, lr):
    """Rows (A0[], A1[], b[], h[]) for one vehicle -- cbf/cbf.py:200-207 / :96-101.
    ``fields[m]`` = 8 floats per slot."""
    A0, A1, b, hs = [], [], [], []
    for m, st in enumerate(slot_types):
        part = slot_partials(st, fields[m], s)
        if model == MODEL_DBM:
            r = dbm_row(part, s[2], s[3], alpha, lr)
        elif model == MODEL_DUM:
            r = dum_row(part, s[2], s[3], alpha)
        else:
            r = kbm_row(part, s[2], alpha)
        A0.append(r[0]); A1.append(r[1]); b.append(r[2]); hs.append(part[0])
    return A0, A1, b, hs


def sadbm_row(p, th, v, beta, alpha, lr):
    """SADBM_CBF_2DS gc/fc + F -- cbf/cbf.py:337-346,386-397: state (x, y, theta, v, beta), controls
    (a, d(beta)/dt).  Lg h = [h_v, h_beta]; h_beta = h_theta for the collision cone
    (cbf/obstacles.py:460-466), 0 for every other obstacle (obstacles.py:124-125) -- only the cone has
    h_theta != 0, so h_theta serves as h_beta."""
    h, h_x, h_y, h_th, h_v, h_t = p
    fc0 = v * float(np.cos(th + beta))
    fc1 = v * float(np.sin(th + beta))
    fc2 = v * float(np.sin(beta)) / lr
    Lf = (h_x * fc0 + h_y * fc1) + h_th * fc2
    return h_v, h_th, -((Lf + alpha * h) + h_t)


def sadbm_filter_step(s, u_ref, beta, beta_ref_last, dt, slot_types, fields, alpha, lr, lf, R):
    """One ``SADBM_CBF_2DS.solve_cbf`` call with a FIXED dt (cbf/cbf.py:348-437; the reference's default
    ``dt = 0.001``; its wall-clock mode is not deterministic).  ``beta`` = the augmented state before the
    call -- also what every CONE slot uses as its beta (cbf.py:424-426 pushes it into the cones after each
    solve, so the rows of a call see the value left by the previous one); ``beta_ref_last`` = the converted
    reference of the previous call.  Returns (a, delta, beta_new, beta_ref, mask, status, beta_ref_dot, rows)."""
    beta_ref = delta_to_beta(u_ref[1], lr, lf)                       # cbf.py:359
    beta_ref_dot = (beta_ref - beta_ref_last) / dt                    # cbf.py:367
    A0, A1, b = [], [], []
    for m, st in enumerate(slot_types):
        f = list(fields[m])
        if (int(st) & SLOT_TYPE_MASK) == SLOT_CONE:
            f[5] = beta
        part = slot_partials(st, f, s)
        r = sadbm_row(part, s[2], s[3], beta, alpha, lr)
        A0.append(r[0]); A1.append(r[1]); b.append(r[2])
    u0, u1, mask, status = qp2_exact(A0, A1, b, u_ref[0], beta_ref_dot, R)
    beta_new = beta + u1 * dt                                         # cbf.py:419
    delta = beta_to_delta(beta_new, lr, lf)                           # cbf.py:429
    return u0, delta, beta_new, beta_ref, mask, status, beta_ref_dot, (A0, A1, b)


def filter_step(model, s, u_ref, slot_types, fields, alpha, lr, lf, L, R, kbm_driver_delta=0):
    """One ``solve_cbf`` call.  DBM: u_ref = [a, delta] -> [a, delta] (cbf/cbf.py:166-220).
    KBM: u_ref = [v, delta] -> [v, delta] (cbf/cbf.py:67-110); with ``kbm_driver_delta`` the
    omega -> delta conversion is the driver's ``arctan(w L / v_cbf)``
    (stanley_controller_ellipse.py:652) instead of the class's ``arctan2(w L, v_ref)`` (cbf.py:109).
    Returns (u0, u1_converted, active_mask, status, u1_raw, h_min)."""
    A0, A1, b, hs = barrier_rows(model, s, slot_types, fields, alpha, lr)
    if model == MODEL_DBM:
        r0 = u_ref[0]
        r1 = delta_to_beta(u_ref[1], lr, lf)
    elif model == MODEL_DUM:
        r0, r1 = u_ref[0], u_ref[1]                           # cbf/cbf.py:253: u_ref = [a, omega] as given
    else:
        r0 = u_ref[0]
        r1 = u_ref[0] * float(np.tan(u_ref[1])) / L           # cbf/cbf.py:75
    u0, u1, mask, status = qp2_exact(A0, A1, b, r0, r1, R)
    if model == MODEL_DBM:
        out1 = beta_to_delta(u1, lr, lf)
    elif model == MODEL_DUM:
        out1 = u1                                            # cbf/cbf.py:293
    elif kbm_driver_delta:
        out1 = float(np.arctan(u1 * L / u0))                 # stanley_controller_ellipse.py:652
    else:
        out1 = float(np.arctan2(u1 * L, r0))                 # cbf/cbf.py:109
    return u0, out1, mask, status, u1, (min(hs) if hs else math.inf)


# --------------------------------------------------------------------------------------
# nominal controllers (function form used by config #1) and the plant
# --------------------------------------------------------------------------------------
def calc_target_index(x, y, yaw, cx, cy, L):
    """stanley_controller_ellipse.py:188-212: global first-minimum argmin of hypot over ALL points."""
    fx = x + L * float(np.cos(yaw))
    fy = y + L * float(np.sin(yaw))
    dx = fx - np.asarray(cx, dtype=np.float64)
    dy = fy - np.asarray(cy, dtype=np.float64)
    d = np.hypot(dx, dy)
    idx = int(np.argmin(d))
    fav0 = -float(np.cos(yaw + PI / 2))
    fav1 = -float(np.sin(yaw + PI / 2))
    e = float(dx[idx]) * fav0 + float(dy[idx]) * fav1
    return idx, e


def stanley_control(x, y, yaw, v, cx, cy, cyaw, last_idx, k, L, ks=0.0):
    """stanley_controller_ellipse.py:146-169.  ``ks`` is LateralStanley's softening
    (cbf/controllers.py:144); the function form has ks = 0."""
    idx, e = calc_target_index(x, y, yaw, cx, cy, L)
    if last_idx >= idx:
        idx = last_idx
    theta_e = normalize_angle(float(cyaw[idx]) - yaw)
    theta_d = float(np.arctan2(k * e, v + ks))
    return theta_e + theta_d, idx


def pid_control(target, current, Kp):
    """stanley_controller_ellipse.py:135-143."""
    return Kp * (target - current)


class PID1:
    """cbf/controllers.py:153-180."""

    def __init__(self, kp=1.0, kd=0.0, ki=0.0, dt=0.1):
        self.kp, self.kd, self.ki, self.dt = kp, kd, ki, dt
        self.eprev = 0.0
        self.ie = 0.0

    def control(self, x, xref):
        e = xref - x
        de = (e - self.eprev) / self.dt
        self.ie += self.dt * e
        u = self.kp * e + self.ki * self.ie + self.kd * de
        self.eprev = e
        return u


def plant_update(s, a, delta, dt, L, max_steer):
    """State.update -- stanley_controller_ellipse.py:86-101."""
    x, y, yaw, v = s
    delta = min(max(delta, -max_steer), max_steer)
    x += v * float(np.cos(yaw)) * dt
    y += v * float(np.sin(yaw)) * dt
    yaw += v / L * float(np.tan(delta)) * dt
    yaw = normalize_angle(yaw)
    v += a * dt
    return [x, y, yaw, v]


def plant_update_by_vel(s, v_cmd, delta, dt, L, max_steer):
    """State.update_by_vel -- stanley_controller_ellipse.py:103-120."""
    x, y, yaw, v = s
    delta = min(max(delta, -max_steer), max_steer)
    x += v * float(np.cos(yaw)) * dt
    y += v * float(np.sin(yaw)) * dt
    yaw += v / L * float(np.tan(delta)) * dt
    yaw = normalize_angle(yaw)
    v = v_cmd
    return [x, y, yaw, v]


def plant_update_com(s, a, delta, dt, lr, lf, max_steer):
    """State.update_com -- stanley_controller_ellipse.py:122-131 (no yaw normalisation)."""
    x, y, yaw, v = s
    delta = min(max(delta, -max_steer), max_steer)
    beta = float(np.arctan2(lr * np.tan(delta), lf + lr))
    c = float(np.cos(yaw))
    sn = float(np.sin(yaw))
    x += (v * c - v * sn * beta) * dt
    y += (v * sn + v * c * beta) * dt
    yaw += (v * beta / lr) * dt
    v += a * dt
    return [x, y, yaw, v], beta


def seeker_update(f, ex, ey, dt, k=0.2, v_min=3.0):
    """RadialObstacleSpawner.update_seekers -- radial_dynamic_obstacles.py:193-239.
    ``f`` = RADIAL slot fields (cx, cy, a, b, kv, vx, vy, -), updated in place."""
    cx, cy = f[0], f[1]
    yaw = float(np.arctan2(ey - cy, ex - cx))
    v_mag = k * float(np.hypot(ex - cx, ey - cy))
    if v_mag < v_min:
        v_mag = v_min
    vx = v_mag * float(np.cos(yaw))
    vy = v_mag * float(np.sin(yaw))
    f[5] = vx
    f[6] = vy
    f[0] = cx + vx * dt
    f[1] = cy + vy * dt
    return f


# --------------------------------------------------------------------------------------
# course: restates test_scripts/PathPlanning/CubicSpline/cubic_spline_planner.py:12-190
# --------------------------------------------------------------------------------------
class _Spline:
    def __init__(self, x, y):
        self.x = list(x)
        self.a = [float(v) for v in y]
        nx = len(x)
        h = np.diff(x)
        A = np.zeros((nx, nx))
        A[0, 0] = 1.0
        for i in range(nx - 1):
            if i != nx - 2:
                A[i + 1, i + 1] = 2.0 * (h[i] + h[i + 1])
            A[i + 1, i] = h[i]
            A[i, i + 1] = h[i]
        A[0, 1] = 0.0
        A[nx - 1, nx - 2] = 0.0
        A[nx - 1, nx - 1] = 1.0
        B = np.zeros(nx)
        for i in range(nx - 2):
            B[i + 1] = 3.0 * (self.a[i + 2] - self.a[i + 1]) / h[i + 1] - 3.0 * (self.a[i + 1] - self.a[i]) / h[i]
        self.c = np.linalg.solve(A, B)
        self.b, self.d = [], []
        for i in range(nx - 1):
            self.d.append((self.c[i + 1] - self.c[i]) / (3.0 * h[i]))
            self.b.append((self.a[i + 1] - self.a[i]) / h[i] - h[i] * (self.c[i + 1] + 2.0 * self.c[i]) / 3.0)

    def _i(self, t):
        return bisect.bisect(self.x, t) - 1

    def calc(self, t):
        i = self._i(t)
        dx = t - self.x[i]
        return self.a[i] + self.b[i] * dx + self.c[i] * dx ** 2.0 + self.d[i] * dx ** 3.0

    def calcd(self, t):
        i = self._i(t)
        dx = t - self.x[i]
        return self.b[i] + 2.0 * self.c[i] * dx + 3.0 * self.d[i] * dx ** 2.0


def calc_spline_course(ax, ay, ds=0.1):
    """cubic_spline_planner.py:178-190 -> (cx, cy, cyaw) as float64 arrays."""
    dxs = np.diff(ax)
    dys = np.diff(ay)
    s = [0]
    s.extend(np.cumsum(np.hypot(dxs, dys)))
    sx = _Spline(s, ax)
    sy = _Spline(s, ay)
    ts = list(np.arange(0, s[-1], ds))
    cx = [sx.calc(t) for t in ts]
    cy = [sy.calc(t) for t in ts]
    cyaw = [math.atan2(sy.calcd(t), sx.calcd(t)) for t in ts]
    return np.array(cx, dtype=np.float64), np.array(cy, dtype=np.float64), np.array(cyaw, dtype=np.float64)


# --------------------------------------------------------------------------------------
# closed-loop rollout for one vehicle (the loop of stanley_controller_ellipse.py:630-830 and
# radial_dynamic_obstacles.py:427-507, minus plotting)
# --------------------------------------------------------------------------------------
NOMINAL_STANLEY = 0   # Stanley steering + P speed (config #1/#2/#4/#5)
NOMINAL_CONST = 1     # constant u_ref (radial_dynamic_obstacles.py:444)

DEFAULT_PARAMS = dict(
    model=MODEL_DBM, nominal=NOMINAL_STANLEY,
    k=0.5, Kp=1.0, dt=0.1, L=2.9, lr=1.45, lf=1.45, max_steer=float(np.radians(30.0)),
    target_speed=30.0 / 3.6, alpha=1.0, R=(1.0, 0.0, 0.0, 1.0),
    t_max=30.0, terminate=0, seeker=0, seeker_k=0.2, seeker_vmin=3.0,
    uref0=0.0, uref1=0.0, kbm_driver_delta=0,
)


def rollout(s0, slot_types, fields, course, T, params=None, record=False, teacher=None):
    """Closed loop for ONE vehicle.

    terminate=1 reproduces ``while max_simulation_time >= time and last_idx > target_idx``
    (stanley_controller_ellipse.py:630) with ``time += dt`` accumulated (:830).
    Returns a dict with the final state, step count, last target index, counters and (if
    ``record``) per-step arrays.  ``teacher`` (optional list of states) forces the state at
    the start of every step (teacher forcing for per-step parity).
    """
    p = dict(DEFAULT_PARAMS)
    if params:
        p.update(params)
    cx, cy, cyaw = course if course is not None else ([], [], [])
    last_idx = len(cx) - 1
    s = [float(v) for v in s0]
    fields = [list(map(float, f)) for f in fields]
    time = 0.0
    if p['nominal'] == NOMINAL_STANLEY:
        target_idx, _ = calc_target_index(s[0], s[1], s[2], cx, cy, p['L'])   # :605
    else:
        target_idx = 0
    steps = 0
    n_active = 0
    n_infeas = 0
    h_min = math.inf
    beta_min, beta_max, beta_int = math.inf, -math.inf, 0.0
    rec = dict(state=[], u=[], beta=[], idx=[], mask=[], status=[], t=[]) if record else None
    while steps < T:
        if p['terminate'] and not (p['t_max'] >= time and last_idx > target_idx):
            break
        if teacher is not None:
            s = [float(v) for v in teacher[steps]]
        if p['nominal'] == NOMINAL_STANLEY:
            a_ref = pid_control(p['target_speed'], s[3], p['Kp'])
            d_ref, target_idx = stanley_control(s[0], s[1], s[2], s[3], cx, cy, cyaw, target_idx, p['k'], p['L'])
        else:
            a_ref, d_ref = p['uref0'], p['uref1']
        if p['model'] == MODEL_KBM and p['nominal'] == NOMINAL_STANLEY:
            uref = [p['target_speed'], d_ref]                                  # :646-648
        else:
            uref = [a_ref, d_ref]
        if len(slot_types) > 0 and p['model'] != MODEL_NONE:
            u0, u1, mask, status, _raw, hm = filter_step(p['model'], s, uref, slot_types, fields,
                                                         p['alpha'], p['lr'], p['lf'], p['L'], p['R'],
                                                         p['kbm_driver_delta'])
        else:
            u0, u1, mask, status, hm = uref[0], uref[1], 0, STATUS_INACTIVE, math.inf
        if record:
            rec['state'].append(list(s)); rec['idx'].append(target_idx)
            rec['mask'].append(mask); rec['status'].append(status)
        if p['model'] == MODEL_DBM:
            s, beta = plant_update_com(s, u0, u1, p['dt'], p['lr'], p['lf'], p['max_steer'])
        elif p['model'] == MODEL_KBM:
            s = plant_update_by_vel(s, u0, u1, p['dt'], p['L'], p['max_steer'])
            beta = 0.0
        else:
            s = plant_update(s, u0, u1, p['dt'], p['L'], p['max_steer'])
            beta = 0.0
        if p['seeker']:
            for m, st in enumerate(slot_types):
                if st == SLOT_RADIAL:
                    seeker_update(fields[m], s[0], s[1], p['dt'], p['seeker_k'], p['seeker_vmin'])
        time += p['dt']
        steps += 1
        n_active += 1 if mask else 0
        n_infeas += 1 if status == STATUS_INFEASIBLE else 0
        h_min = min(h_min, hm)
        beta_min = min(beta_min, beta); beta_max = max(beta_max, beta); beta_int += beta * p['dt']
        if record:
            rec['u'].append([u0, u1]); rec['beta'].append(beta); rec['t'].append(time)
    out = dict(state=s, steps=steps, target_idx=target_idx, n_active=n_active, n_infeasible=n_infeas,
               h_min=h_min, beta_min=beta_min, beta_max=beta_max, beta_int=beta_int, time=time,
               fields=fields)
    if record:
        out['rec'] = rec
    return out


def drive_ticks(s0, ego, box_ids, boxes, dts, fixed_types, fixed_fields, M, traj, params=None, drive=None,
                target_idx=0, carry=(0.0, 0.0, 0.0, 0.0)):
    """The per-tick loop of the CARLA driver for ONE ego vehicle, tick by tick
    (carla_scripts/multi_obstacle_CBF_local_with_lanes.py:861-983):

      delta, idx = lateral_stanley.control()   cbf/controllers.py:104-151 (class form: offset lf, softening ks)
      delta *= rad_to_steer                     :876-877
      acc_pid.set_dt(dt); u_a = acc_pid.control(v, trajectory[idx][3])      cbf/controllers.py:153-180, :897-903
      list = fixed obstacles (the lanes, :913-916) + one CollisionCone2D(hypot(extent), s, [x, y, yaw, |v|]) per box (:918-928)
      u = solve_cbf([u_a, delta]) unless the list is empty (:935-953)
      throttle / brake / steer                  :955-976

    ego: list of T states (the simulator's), or None -> stand-in plant State.update_com from s0 with the tick's dt.
    box_ids[t] / boxes[t]: K ids (< 0 = none) / K boxes (extent.x, extent.y, x, y, yaw, speed); dts[t] the tick's dt.
    traj = (x, y, yaw, v) lists.  Returns per-tick lists act (throttle, brake, steer), u, mask, idx + final carry / state."""
    p = dict(DEFAULT_PARAMS)
    if params:
        p.update(params)
    d = dict(kp=1.0, ki=0.01, kd=0.01, rad_to_steer=1.0, max_steer_cmd=1.0, rate=0.1, cone_buffer=1.5, reset_brake=False)
    if drive:
        d.update(drive)
    tx, ty, tyaw, tv = traj
    s = [float(v) for v in (s0 if s0 is not None else ego[0])]
    pid = PID1(kp=d['kp'], kd=d['kd'], ki=d['ki'])
    pid.eprev, pid.ie = float(carry[0]), float(carry[1])
    thr_prev, brk_prev = float(carry[2]), float(carry[3])
    last_idx = int(target_idx)
    T = len(dts)
    out = dict(act=[], u=[], mask=[], idx=[])
    for t in range(T):
        if ego is not None:
            s = [float(v) for v in ego[t]]
        dt = float(dts[t])
        delta, last_idx = stanley_control(s[0], s[1], s[2], s[3], tx, ty, tyaw, last_idx, p['k'], p['L'], p.get('ks', 0.0))
        delta = delta * d['rad_to_steer']
        pid.dt = dt
        u_a = pid.control(s[3], float(tv[last_idx]))
        types = list(fixed_types)
        fields = [list(map(float, f)) for f in fixed_fields]
        for k, bid in enumerate(box_ids[t]):
            if bid < 0 or len(types) >= M:
                continue
            ex, ey, lx, ly, yaw_o, sp = (float(v) for v in boxes[t][k])
            types.append(SLOT_CONE)
            fields.append([lx, ly, yaw_o, sp, float(np.hypot(ex, ey)) + d['cone_buffer'], 0.0, 0.0, 0.0])
        if len(types) < 1:
            u0, u1, mask = u_a, delta, 0
        else:
            u0, u1, mask, _st, _raw, _hm = filter_step(MODEL_DBM, s, [u_a, delta], types, fields, p['alpha'], p['lr'], p['lf'],
                                                       p['L'], p['R'], 0)
        thr, brk, steer = actuator_shaping(u0, u1, thr_prev, brk_prev, d['max_steer_cmd'], d['rate'], d['reset_brake'])
        thr_prev, brk_prev = thr, brk
        out['act'].append([thr, brk, steer]); out['u'].append([u0, u1]); out['mask'].append(mask); out['idx'].append(last_idx)
        if ego is None:
            s, _beta = plant_update_com(s, u0, u1, dt, p['lr'], p['lf'], p['max_steer'])
    out['carry'] = [pid.eprev, pid.ie, thr_prev, brk_prev]
    out['target_idx'] = last_idx
    out['state'] = s
    return out
