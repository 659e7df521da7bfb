"""CPU tests (no GPU): the C oracle against the pinned Python oracle, and properties of the exact
2-variable QP that replaces cvxopt.solvers.cp (KKT residuals, uniqueness of the
KKT working set, m = 1 closed form vs enumeration)."""
import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import oracle as o
from tests import helpers as H


def test_c_oracle_filter_step_equals_python_oracle():
    rng = np.random.default_rng(42)
    slots = [o.SLOT_ELLIPSE, o.SLOT_CONE, o.SLOT_LANE, o.SLOT_RADIAL, o.SLOT_DISTANCE, o.SLOT_CONE, o.SLOT_LANE_SQRT]
    N = 300
    s = H.random_states(rng, N); ob = H.random_slots(rng, N, slots, s); ur = H.random_uref(rng, N)
    R = (1.0, 0.3, 0.3, 2.5)
    for model in (o.MODEL_DBM, o.MODEL_KBM):
        c = co.filter_step(co.default_params(model=model, R=R, alpha=0.7), slots, s, ob, ur, rows=True)
        for n in range(N):
            fields = [list(ob[m, :, n]) for m in range(len(slots))]
            A0, A1, b, hs = o.barrier_rows(model, list(s[:, n]), slots, fields, 0.7, 1.45)
            # numpy's scalar sin/cos and glibc's may differ by an ulp: rows to 1e-12, decisions equal
            assert np.allclose(c["A"][0, :, n], A0, rtol=1e-12, atol=1e-12)
            assert np.allclose(c["A"][1, :, n], A1, rtol=1e-12, atol=1e-12)
            assert np.allclose(c["b"][:, n], b, rtol=1e-12, atol=1e-12)
            u0, u1, mask, status, _, hm = o.filter_step(model, list(s[:, n]), list(ur[:, n]), slots, fields, 0.7, 1.45, 1.45, 2.9, R)
            assert mask == int(c["mask"][n]) and status == int(c["status"][n]), n
            assert abs(u0 - c["u"][0, n]) <= 1e-10 * (1 + abs(u0)) and abs(u1 - c["u"][1, n]) <= 1e-10 * (1 + abs(u1))
            assert abs(hm - c["h_min"][n]) <= 1e-12 * (1 + abs(hm))


def test_c_oracle_rollout_equals_python_oracle():
    from sccav_cbf_b200 import scenarios as sc
    b = sc.config2(n_total=65536, M=8, T=150, lo=100, hi=108)
    c = co.rollout(co.default_params(), b.slot_desc, b.state, b.obst, b.course, b.T, record_stride=1)
    for n in range(b.N):
        p = o.rollout(b.state[:, n], b.slot_desc, [b.obst[m, :, n] for m in range(b.M)], b.course, b.T, record=True)
        assert p["steps"] == c["steps"][n] and p["target_idx"] == c["target_idx"][n]
        assert p["n_active"] == c["n_active"][n] and p["n_infeasible"] == c["n_infeasible"][n]
        assert np.array_equal(np.array(p["rec"]["idx"]), c["traj_idx"][:, n])
        assert np.array_equal(np.array(p["rec"]["mask"], dtype=np.uint32), c["traj_mask"][:, n])
        assert np.abs(np.array(p["rec"]["state"]) - c["traj"][:, 0:4, n]).max() <= 1e-9
        assert np.abs(np.array(p["state"]) - c["state"][:, n]).max() <= 1e-9


def test_c_oracle_seeker_rollout_equals_python_oracle():
    from sccav_cbf_b200 import scenarios as sc
    b = sc.config3(n_total=262144, M=4, T=90, lo=0, hi=6)
    prm = dict(b.params)
    c = co.rollout(co.default_params(**prm), b.slot_desc, b.state, b.obst, b.course, b.T)
    for n in range(b.N):
        p = o.rollout(b.state[:, n], b.slot_desc, [b.obst[m, :, n] for m in range(b.M)], None, b.T,
                      params=dict(seeker=1, dt=1.0 / 30.0, nominal=o.NOMINAL_CONST))
        assert p["n_active"] == c["n_active"][n] and p["n_infeasible"] == c["n_infeasible"][n]
        assert np.abs(np.array(p["state"]) - c["state"][:, n]).max() <= 1e-8
        assert np.abs(np.array(p["fields"]) - c["obst"][:, :, n]).max() <= 1e-8


def _random_qp(rng, m):
    A0 = rng.normal(size=m) * rng.choice([0.0, 1.0, 30.0], size=m)
    A1 = rng.normal(size=m) * rng.choice([1.0, 50.0], size=m)
    b = rng.normal(size=m) * 10 - 5
    r = rng.normal(size=2) * np.array([2.0, 0.3])
    Rm = np.array([[1.0, 0.0], [0.0, 1.0]]) if rng.uniform() < 0.5 else np.array([[1.5, 0.4], [0.4, 2.0]])
    return A0, A1, b, r, Rm


def test_qp2_unique_kkt_point_and_residuals():
    rng = np.random.default_rng(7)
    n_act = 0
    for it in range(1500):
        m = int(rng.integers(1, 9))
        A0, A1, b, r, Rm = _random_qp(rng, m)
        cands = []
        u0, u1, mask, status = o.qp2_exact(list(A0), list(A1), list(b), r[0], r[1], tuple(Rm.ravel()), collect=cands)
        if status == o.STATUS_INFEASIBLE:
            assert not cands
            continue
        # strict convexity: every KKT working set gives the same point
        for c in cands:
            assert abs(c[0] - u0) <= 1e-8 * (1 + abs(u0)) and abs(c[1] - u1) <= 1e-8 * (1 + abs(u1))
        A = np.stack([A0, A1])[:, :, None]
        prim, stat, lmin = H.kkt_residuals(A, b[:, None], np.array([[u0], [u1]]), r[:, None], Rm.ravel(), np.array([mask]))
        assert prim[0] <= 1e-9 and stat[0] <= 1e-9 and lmin[0] >= -1e-9
        n_act += status == o.STATUS_ACTIVE
        # cost is minimal among a cloud of feasible points
        cost = (np.array([u0, u1]) - r) @ Rm @ (np.array([u0, u1]) - r)
        pts = np.array([u0, u1]) + rng.normal(size=(64, 2)) * 0.5
        feas = ((pts[:, 0:1] * A0 + pts[:, 1:2] * A1) >= b - 1e-12).all(axis=1)
        for q in pts[feas]:
            assert (q - r) @ Rm @ (q - r) >= cost - 1e-9 * (1 + cost)
    assert n_act > 200


def test_qp2_single_row_closed_form():
    """m = 1: u = r + R^-1 A^T (b - A r)/(A R^-1 A^T) when violated (the closed form the survey
    probe used against beta_vs_time.mat; also members_scripts' psi-projection)."""
    rng = np.random.default_rng(9)
    for _ in range(500):
        A0, A1, b, r, Rm = _random_qp(rng, 1)
        u0, u1, mask, status = o.qp2_exact(list(A0), list(A1), list(b), r[0], r[1], tuple(Rm.ravel()))
        a = np.array([A0[0], A1[0]])
        if a @ r >= b[0] or not np.any(a):
            assert (u0, u1, mask) == (r[0], r[1], 0) or not np.any(a)
        else:
            Ri = np.linalg.inv(Rm)
            ex = r + Ri @ a * (b[0] - a @ r) / (a @ Ri @ a)
            assert np.allclose([u0, u1], ex, rtol=1e-12, atol=1e-12) and mask == 1


def test_qp2_degenerate_cases():
    R = (1.0, 0.0, 0.0, 1.0)
    # v = 0: every row is 0*u >= b: vacuous if b <= 0, infeasible otherwise
    assert o.qp2_exact([0.0, 0.0], [0.0, 0.0], [-1.0, -2.0], 0.3, 0.1, R) == (0.3, 0.1, 0, o.STATUS_INACTIVE)
    assert o.qp2_exact([0.0], [0.0], [1.0], 0.3, 0.1, R) == (0.3, 0.1, 0, o.STATUS_INFEASIBLE)
    # ellipse rows constrain beta only: interval clipping, tightest bound wins, a untouched
    u0, u1, mask, st = o.qp2_exact([0.0, 0.0, 0.0], [1.0, 2.0, -1.0], [0.2, 0.9, -1.0], 0.5, 0.0, R)
    assert (u0, u1, mask, st) == (0.5, 0.45, 0b010, o.STATUS_ACTIVE)
    # duplicate rows: lowest index reported
    u0, u1, mask, st = o.qp2_exact([0.0, 0.0], [1.0, 1.0], [0.2, 0.2], 0.5, 0.0, R)
    assert (u1, mask, st) == (0.2, 0b01, o.STATUS_ACTIVE)
    # conflicting parallel rows: infeasible, least-violation candidate, flagged
    u0, u1, mask, st = o.qp2_exact([0.0, 0.0], [1.0, -1.0], [0.3, -0.1], 0.5, 0.0, R)
    assert st == o.STATUS_INFEASIBLE and mask in (0b01, 0b10)
    # vertex solution
    u0, u1, mask, st = o.qp2_exact([1.0, 0.0], [0.0, 1.0], [1.0, 1.0], 0.0, 0.0, R)
    assert (u0, u1, mask, st) == (1.0, 1.0, 0b11, o.STATUS_ACTIVE)


def test_c_oracle_qp_degenerate_cases_match_python():
    """Same degenerate inputs through the C oracle's filter (distance slots give hand-set rows is
    not possible; instead compare on many random problems incl. infeasible ones)."""
    rng = np.random.default_rng(13)
    slots = [o.SLOT_LANE, o.SLOT_LANE, o.SLOT_ELLIPSE, o.SLOT_ELLIPSE]
    N = 400
    s = H.random_states(rng, N); ob = H.random_slots(rng, N, slots, s); ur = H.random_uref(rng, N)
    c = co.filter_step(co.default_params(), slots, s, ob, ur)
    n_inf = 0
    for n in range(N):
        fields = [list(ob[m, :, n]) for m in range(len(slots))]
        u0, u1, mask, status, _, _ = o.filter_step(o.MODEL_DBM, list(s[:, n]), list(ur[:, n]), slots, fields, 1.0, 1.45, 1.45, 2.9, (1.0, 0.0, 0.0, 1.0))
        assert mask == int(c["mask"][n]) and status == int(c["status"][n]), n
        n_inf += status == o.STATUS_INFEASIBLE
    assert n_inf > 5          # all rows here constrain beta only: conflicts are common


def test_prepared_ellipse_equals_canonical_ellipse_in_both_oracles():
    """ELLIPSE_PREP (ingest once, solve many): the regrouped functions equal Ellipse2D's
    (cbf/obstacles.py:193,218,229,316) to a few ulp, in the Python and in the C oracle, with and
    without the STATIC flag."""
    rng = np.random.default_rng(9)
    N = 200
    s = H.random_states(rng, N)
    slots = [o.SLOT_ELLIPSE, o.SLOT_ELLIPSE | o.SLOT_STATIC, o.SLOT_CONE]
    ob = H.random_slots(rng, N, slots, s)
    ur = H.random_uref(rng, N)
    prep = ob.copy()
    sd_p = list(slots)
    for m in (0, 1):
        for n in range(N):
            prep[m, :, n] = o.prepare_ellipse(ob[m, :, n])
        sd_p[m] = (slots[m] & ~o.SLOT_TYPE_MASK) | o.SLOT_ELLIPSE_PREP
    for n in range(0, N, 7):
        st = tuple(s[:, n])
        for m in (0, 1):
            pc = o.slot_partials(slots[m], ob[m, :, n], st)
            pp = o.slot_partials(sd_p[m], prep[m, :, n], st)
            assert np.allclose(pc, pp, rtol=1e-12, atol=1e-13), (m, n)
        assert o.slot_partials(sd_p[1], prep[1, :, n], st)[5] == 0.0          # STATIC: h_t = 0 although wx, wy != 0
    c_can = co.filter_step(co.default_params(alpha=0.9), slots, s, ob, ur, rows=True)
    c_pre = co.filter_step(co.default_params(alpha=0.9), sd_p, s, prep, ur, rows=True)
    assert np.allclose(c_can["A"], c_pre["A"], rtol=1e-11, atol=1e-12) and np.allclose(c_can["b"], c_pre["b"], rtol=1e-11, atol=1e-11)
    assert np.array_equal(c_can["mask"], c_pre["mask"]) and np.array_equal(c_can["status"], c_pre["status"])
    assert np.allclose(c_can["u"], c_pre["u"], rtol=1e-10, atol=1e-12)


def test_qp_shortcut_theory_most_violated_row_in_the_metric_of_R():
    """The statement the kernels' one-scan QP shortcut rests on (csrc/path.cuh, QpScan): if the optimum has exactly ONE
    active row, that row is the one with the largest rk^2 / (A_k R^-1 A_k^T) among the rows the reference point violates
    (its projection has to clear every other violated half-plane).  Checked against the enumerating oracle, with a numpy
    mirror of the scan -- including the direction the kernel relies on: whenever the scan's candidate passes the
    feasibility test (and is not a near-tie), it IS the oracle's answer."""
    rng = np.random.default_rng(21)
    n_single = n_accept = 0
    for it in range(4000):
        m = int(rng.integers(1, 9))
        A0, A1, b, r, Rm = _random_qp(rng, m)
        if it % 3 == 0:
            A0[:] = 0.0                                               # the all-ellipse DBM shape: rows constrain u1 only
        u0, u1, mask, status = o.qp2_exact(list(A0), list(A1), list(b), r[0], r[1], tuple(Rm.ravel()))
        Ri = np.linalg.inv(Rm)
        rk = A0 * r[0] + A1 * r[1] - b
        g = Ri @ np.stack([A0, A1])
        den = A0 * g[0] + A1 * g[1]
        viol = (rk < 0) & (den > 0)
        if not viol.any():
            continue
        ratio = np.where(viol, rk * rk / np.where(den > 0, den, 1.0), -1.0)
        kb = int(np.argmax(ratio))
        runner = np.partition(ratio, -2)[-2] if m > 1 else -1.0
        if status == o.STATUS_ACTIVE and bin(mask).count("1") == 1:
            n_single += 1
            k = mask.bit_length() - 1
            assert ratio[k] >= ratio[kb] * (1 - 1e-9), (it, k, kb)    # the active row is the most violated one
        # the kernel's direction: candidate of kb feasible (and a clear winner) => it is the oracle's answer
        t = -rk[kb] / den[kb]
        c = r + g[:, kb] * t
        res = A0 * c[0] + A1 * c[1] - b
        tol = 1e-12 * (np.abs(A0 * c[0]) + np.abs(A1 * c[1]) + np.abs(b))
        ok = all(res[j] >= -tol[j] for j in range(m) if j != kb)
        if ok and runner < ratio[kb] * (1 - 1e-6):
            n_accept += 1
            assert status == o.STATUS_ACTIVE and mask == 1 << kb, (it, mask, kb)
            assert abs(c[0] - u0) <= 1e-12 * (1 + abs(u0)) and abs(c[1] - u1) <= 1e-12 * (1 + abs(u1))
    assert n_single > 500 and n_accept > 500


def _scipy_qp(A0, A1, b, r, R):
    """An INDEPENDENT solver for min (u - r)^T R (u - r) s.t. A u >= b: scipy's SLSQP (nothing of this repo inside)."""
    from scipy.optimize import minimize
    R = np.asarray(R, dtype=np.float64).reshape(2, 2)
    A = np.stack([A0, A1], axis=1)
    res = minimize(lambda u: (u - r) @ R @ (u - r), x0=np.array(r, dtype=np.float64), jac=lambda u: 2.0 * R @ (u - r), method="SLSQP",
                   constraints=[{"type": "ineq", "fun": lambda u: A @ u - b, "jac": lambda u: A}], options={"ftol": 1e-15, "maxiter": 200})
    return res.x, res.success


def random_qps(rng, n, m):
    """Feasible random problems around a reference point that violates one to three rows (single and pair optima)."""
    out = []
    while len(out) < n:
        A0 = rng.normal(0, 1, m); A1 = rng.normal(0, 1, m)
        inner = rng.normal(0, 1, 2)                                    # a point every row admits: the problem is feasible
        b = A0 * inner[0] + A1 * inner[1] - rng.uniform(0.05, 2.0, m)
        r = inner + rng.normal(0, 2.5, 2)
        R = (1.0, 0.0, 0.0, 1.0) if len(out) % 2 else (2.0, 0.3, 0.3, 0.7)
        out.append((A0, A1, b, r, R))
    return out


def test_exact_qp_against_an_independent_solver():
    """the golden vectors take their QP solutions from this repo's own exact solver, so the solver needs a check
    that shares no code with it.  300 random feasible problems (m = 2..8 rows, optima on no row, one row and a PAIR of rows,
    diagonal and full weights) against scipy's SLSQP: same optimum to 1e-6, and the active set qp2_exact reports is tight."""
    rng = np.random.default_rng(20261017)
    n_pair = n_single = n_ok = 0
    for i, (A0, A1, b, r, R) in enumerate(random_qps(rng, 300, 2 + rng.integers(0, 7))):
        u0, u1, mask, status = o.qp2_exact(list(A0), list(A1), list(b), float(r[0]), float(r[1]), R)[:4]
        xs, ok = _scipy_qp(A0, A1, b, r, R)
        assert status != o.STATUS_INFEASIBLE
        if not ok:                       # (SLSQP sometimes stops on its line search at this ftol; it is the checker, not the subject)
            continue
        n_ok += 1
        assert abs(u0 - xs[0]) <= 1e-6 * (1 + abs(xs[0])) and abs(u1 - xs[1]) <= 1e-6 * (1 + abs(xs[1])), (i, u0, u1, xs)
        act = [k for k in range(len(b)) if (mask >> k) & 1]
        for k in act:
            assert abs(A0[k] * u0 + A1[k] * u1 - b[k]) <= 1e-9 * (1 + abs(b[k]))
        n_pair += len(act) == 2
        n_single += len(act) == 1
    assert n_ok > 150 and n_pair > 30 and n_single > 60, (n_ok, n_pair, n_single)
