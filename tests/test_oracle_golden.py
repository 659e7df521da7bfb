"""The oracle (oracle/oracle.py) against every golden vector the reference offers for the path.

1. test_scripts/beta_vs_time.mat -- the reference's only numeric artefact (tests/golden/beta_vs_time.json).
2. tests/golden/reference_vectors.npz -- produced by executing the reference's own cbf/obstacles.py,
   cbf/cbf.py and the solver/plant/Stanley functions of test_scripts/*.py (gen_reference_vectors.py).
"""
import json
import math
import os

import numpy as np
import pytest

from oracle import oracle as o

AX = [0.0, 100.0, 100.0, 50.0, 60.0]
AY = [0.0, 0.0, -30.0, -20.0, 0.0]


@pytest.fixture(scope="module")
def course():
    return o.calc_spline_course(AX, AY, 0.1)


def test_course_bit_exact_vs_reference_planner(course, refvec):
    cx, cy, cyaw = course
    assert len(cx) == 2034
    assert np.array_equal(cx, refvec["course"][0])
    assert np.array_equal(cy, refvec["course"][1])
    assert np.array_equal(cyaw, refvec["course"][2])
    assert cx[1524] == 75.63083239590846 and cy[1524] == -35.149858832676415


def test_beta_vs_time_mat(course, golden_dir):
    """277 samples, t_arr bit-exact, |dbeta| <= 1e-3 deg (IPM tolerance)."""
    g = json.load(open(os.path.join(golden_dir, "beta_vs_time.json")))
    cx, cy, cyaw = course
    oi = int((len(cx) - 1) * 0.75)
    a_cone = float(np.hypot(20, 10) / 2)
    fields = [[cx[oi], cy[oi], 0.0, 0.0, a_cone + 1.5, 0.0, 0.0, 0.0]]
    out = o.rollout([-0.0, 5.0, float(np.radians(20.0)), 10.0], [o.SLOT_CONE], fields, course, T=10 ** 6,
                    params=dict(terminate=1, R=(0.5, 0.0, 0.0, 0.5)), record=True)
    beta = np.degrees(np.array([0.0] + out["rec"]["beta"]))
    t = np.array([0.0] + out["rec"]["t"])
    assert len(beta) == 277 == len(g["beta_deg"])
    assert np.array_equal(t, np.array(g["t_arr"]))          # fp64 time accumulation, bit-exact
    assert t[-1] == 27.600000000000122
    assert np.abs(beta - np.array(g["beta_deg"])).max() <= 1e-3
    assert out["n_active"] == 59 and out["target_idx"] == 2033


def test_cone_partials_bit_exact(refvec):
    for row, ref in zip(refvec["cone_in"], refvec["cone_out"]):
        x, y, th, v, cx, cy, tho, vo, a, beta = row
        got = o.cone_partials(x, y, th, v, cx, cy, tho, vo, a + 1.5, beta)
        assert tuple(float(g) for g in got) == tuple(ref), (row, got, ref)


def test_ellipse_partials_bit_exact(refvec):
    for row, ref in zip(refvec["ellipse_in"], refvec["ellipse_out"]):
        x, y, cx, cy, a, b, th, buf, vx, vy = row
        h, hx, hy, hth, hv, ht = o.ellipse_partials(x, y, cx, cy, a + buf, b + buf, th, vx, vy)
        assert (float(h), float(hx), float(hy), float(ht)) == tuple(ref)
        assert hth == 0.0 and hv == 0.0


def test_lane_vs_reference_newton_cg(refvec):
    """cx within the reference's own xtol (1e-8 .. scipy stops early); h, h_x, h_y follow."""
    for row, ref in zip(refvec["lane_in"], refvec["lane_out"]):
        x, y, c0, c1, c2, c3, buf, th, v = row
        c = [c0, c1, c2, c3, 0.0, 0.0]
        cx = o.lane_closest_x(c, x, y)
        assert abs(cx - ref[0]) <= 2e-7 * (1 + abs(cx)), (row, cx, ref[0])
        h, hx, hy, _, _, _ = o.lane_partials(x, y, c, buf)
        assert abs(h - ref[1]) <= 1e-6 * (1 + abs(ref[1]))
        assert abs(hx - ref[2]) <= 1e-6 * (1 + abs(ref[2]))
        assert abs(hy - ref[3]) <= 1e-6 * (1 + abs(ref[3]))
        # and the oracle's point is at least as good a minimiser as the reference's
        def D(t):
            g, _, _ = o._poly3(c, t)
            return (t - x) ** 2 + (g - y) ** 2
        assert D(cx) <= D(ref[0]) * (1 + 1e-12) + 1e-300


def test_lane_linear_closed_form():
    rng = np.random.default_rng(3)
    for _ in range(50):
        c0, c1 = rng.uniform(-20, 20), rng.uniform(-2, 2)
        px, py = rng.uniform(-50, 50, 2)
        cx = o.lane_closest_x([c0, c1, 0, 0, 0, 0], px, py)
        exact = (px + c1 * (py - c0)) / (1 + c1 * c1)
        assert abs(cx - exact) <= 1e-12 * (1 + abs(exact))


def test_dbm_rows_and_output_vs_reference_class(refvec):
    """Rows assembled by cbf/cbf.py:200-207 (captured inside solvers.cp) and the final [a, delta]."""
    for i in range(refvec["dbm_in"].shape[0]):
        s = list(refvec["dbm_in"][i, 0:4]); uref = list(refvec["dbm_in"][i, 4:6])
        R = tuple(refvec["dbm_in"][i, 6:10]); alpha = refvec["dbm_in"][i, 10]; m = int(refvec["dbm_in"][i, 11])
        types = [int(refvec["dbm_slot"][i, j, 0]) for j in range(m)]
        fields = [list(refvec["dbm_slot"][i, j, 1:]) for j in range(m)]
        A0, A1, b, _ = o.barrier_rows(o.MODEL_DBM, s, types, fields, alpha, 1.45)
        ref = refvec["dbm_rows"][i, :m]
        for j in range(m):
            if types[j] == o.SLOT_LANE:      # lane rows inherit the Newton-CG stopping error
                assert np.allclose([A0[j], A1[j], b[j]], ref[j], rtol=1e-6, atol=1e-6)
            else:
                assert (float(A0[j]), float(A1[j]), float(b[j])) == tuple(ref[j]), (i, j)
        u0, d, mask, status, raw, _ = o.filter_step(o.MODEL_DBM, s, uref, types, fields, alpha, 1.45, 1.45, 2.9, R)
        ru0, rd, rbeta, rmask, rstatus = refvec["dbm_out"][i]
        assert o.delta_to_beta(uref[1], 1.45, 1.45) == rbeta
        assert mask == int(rmask) and status == int(rstatus)
        tol = 1e-6 if o.SLOT_LANE in types else 1e-13
        assert abs(u0 - ru0) <= tol * (1 + abs(ru0)) and abs(d - rd) <= tol * (1 + abs(rd))


@pytest.mark.parametrize("cbf_type", [0, 2, 4])
def test_config1_closed_loop_vs_reference_functions(course, refvec, cbf_type):
    """Closed loop of stanley_controller_ellipse.py main() for CBF_TYPE 0 (KBM, CBF()),
    2 (DBM ellipse, CBF_A) and 4 (DBM cone, class API): step count, waypoint indices and active
    sets identical; states/controls to 1e-9 (the twins CBF/CBF_A use algebraically equal but
    differently rounded axis-aligned formulas, so bit equality is not expected for 0 and 2)."""
    ref = refvec["cfg1_type%d" % cbf_type]
    cx, cy, cyaw = course
    oi = int((len(cx) - 1) * 0.75)
    if cbf_type == 4:
        types = [o.SLOT_CONE]
        fields = [[cx[oi], cy[oi], 0.0, 0.0, float(np.hypot(20, 10) / 2) + 1.5, 0.0, 0, 0]]
        prm = dict(terminate=1, R=(0.5, 0.0, 0.0, 0.5))
    else:
        types = [o.SLOT_ELLIPSE]
        fields = [[cx[oi], cy[oi], 20.0, 10.0, 0.0, 0.0, 0.0, 0]]
        prm = dict(terminate=1, model=o.MODEL_KBM if cbf_type == 0 else o.MODEL_DBM, kbm_driver_delta=1)
    out = o.rollout([-0.0, 5.0, float(np.radians(20.0)), 10.0], types, fields, course, T=10 ** 6, params=prm, record=True)
    assert out["steps"] == ref.shape[0]
    rec = out["rec"]
    assert np.array_equal(np.array(rec["idx"]), ref[:, 5].astype(int))
    assert np.array_equal(np.array(rec["mask"]), ref[:, 9].astype(int))
    assert np.array_equal(np.array(rec["t"]), ref[:, 11])
    st = np.array(rec["state"]); u = np.array(rec["u"])
    if cbf_type == 4:
        # class path: rows are bit-exact (test_dbm_rows...); the trajectory is not quite, because
        # the reference's np.dot([dx, dy], front_axle_vec) (sce.py:210) goes through BLAS ddot,
        # whose FMA use is build-dependent (1 ulp in the cross-track error).
        assert np.abs(st - ref[:, 0:4]).max() <= 1e-11
        assert np.abs(u - ref[:, 6:8]).max() <= 1e-11
        assert np.abs(np.array(rec["beta"]) - ref[:, 8]).max() <= 1e-11
        # teacher-forced: rows from the reference's states are bit-exact
        for i in range(0, ref.shape[0], 7):
            A0, A1, b, _ = o.barrier_rows(o.MODEL_DBM, list(ref[i, 0:4]), types, fields, 1, 1.45)
            assert (float(A0[0]), float(A1[0]), float(b[0])) == tuple(ref[i, 12:15])
    else:
        assert np.abs(st - ref[:, 0:4]).max() <= 1e-9
        # type 0 runs with the driver's omega->delta conversion (sce.py:652), see filter_step
        assert np.abs(u - ref[:, 6:8]).max() <= 1e-9


def test_radial_rows_vs_reference_function(refvec):
    for row, ref in zip(refvec["radial_in"], refvec["radial_out"]):
        s = list(row[0:4]); uref = list(row[4:6]); c = row[6:8]; cd = row[8:10]; r, gamma, kv = row[10:13]
        f = [[c[0], c[1], r, r, kv, cd[0], cd[1], 0.0]]
        A0, A1, b, _ = o.barrier_rows(o.MODEL_DBM, s, [o.SLOT_RADIAL], f, gamma, 1.45)
        assert np.allclose([A0[0], A1[0], b[0]], ref[2:5], rtol=1e-13, atol=1e-13)
        u0, d, mask, status, _, _ = o.filter_step(o.MODEL_DBM, s, uref, [o.SLOT_RADIAL], f, gamma, 1.45, 1.45, 2.9, (1.0, 0.0, 0.0, 1.0))
        assert mask == int(ref[5]) and status == int(ref[6])
        assert abs(u0 - ref[0]) <= 1e-12 * (1 + abs(ref[0])) and abs(d - ref[1]) <= 1e-12 * (1 + abs(ref[1]))


@pytest.fixture(scope="module")
def refvec2(golden_dir):
    """tests/golden/reference_vectors_lane_sadbm.npz -- produced by gen_reference_vectors_lane_sadbm.py from
    the reference's CBF_lane / CBF_lane_sqrt driver functions and its SADBM_CBF_2DS class."""
    return np.load(os.path.join(golden_dir, "reference_vectors_lane_sadbm.npz"))


def test_lane_squared_and_distance_rows_vs_reference_functions(refvec2):
    """CBF_lane / CBF_lane_sqrt (stanley_controller_ellipse.py:416-512): row and solution of the LANE and
    LANE_SQRT slots.  The closest abscissa comes from scipy's Newton-CG in the reference (xtol 1e-8), hence
    1e-7; the closed-form variants CBF_lane_cf* (different, mis-parenthesised Lg) are not a target."""
    I, rows, U = refvec2["lanev_in"], refvec2["lanev_rows"], refvec2["lanev_u"]
    nact = 0
    for i in range(I.shape[0]):
        s = list(I[i, :4]); co = list(I[i, 4:8]) + [0.0, 0.0]; ud = I[i, 8:10]; buf, al = I[i, 10], I[i, 11]
        for j, t in enumerate((o.SLOT_LANE, o.SLOT_LANE_SQRT)):
            f = [[buf] + co + [0.0]]
            A0, A1, b, _ = o.barrier_rows(o.MODEL_DBM, s, [t], f, al, 1.45)
            assert A0[0] == rows[i, j, 0] == 0.0
            assert abs(A1[0] - rows[i, j, 1]) <= 1e-7 * (1 + abs(rows[i, j, 1]))
            assert abs(b[0] - rows[i, j, 2]) <= 1e-7 * (1 + abs(rows[i, j, 2]))
            u0, u1, mask, status = o.qp2_exact(A0, A1, b, ud[0], ud[1], (1.0, 0.0, 0.0, 1.0))
            assert abs(u0 - U[i, j, 0]) <= 1e-8 and abs(u1 - U[i, j, 1]) <= 1e-7 * (1 + abs(U[i, j, 1]))
            nact += mask != 0
    assert nact > 20


def test_sadbm_class_sequences_vs_reference(refvec2):
    """SADBM_CBF_2DS.solve_cbf (cbf/cbf.py:348-437) with a fixed dt over 16 sequences of 6 ticks: rows as the
    class assembles them, converted reference, solution, carried beta / beta_ref_last."""
    cfg, M, cone = refvec2["sadbm_cfg"], refvec2["sadbm_m"], refvec2["sadbm_cone"]
    I, rows, out = refvec2["sadbm_in"], refvec2["sadbm_rows"], refvec2["sadbm_out"]
    nact = 0
    for i in range(cfg.shape[0]):
        alpha, dt, lr, lf = cfg[i, :4]; R = tuple(cfg[i, 4:8]); m = int(M[i])
        beta, brl = I[i, 0, 6], I[i, 0, 7]
        for t in range(I.shape[1]):
            s = list(I[i, t, :4]); ur = list(I[i, t, 4:6])
            assert abs(beta - I[i, t, 6]) <= 1e-13 and abs(brl - I[i, t, 7]) <= 1e-13      # carried state
            fields = [[*cone[i, t, j], 0.0, 0.0, 0.0] for j in range(m)]
            u0, d, bn, br, mask, st, brd, (A0, A1, b) = o.sadbm_filter_step(s, ur, beta, brl, dt, [o.SLOT_CONE] * m, fields, alpha, lr, lf, R)
            for j in range(m):
                for got, ref in ((A0[j], rows[i, t, j, 0]), (A1[j], rows[i, t, j, 1]), (b[j], rows[i, t, j, 2])):
                    assert abs(got - ref) <= 1e-13 * (1 + abs(ref))
            assert mask == int(out[i, t, 4]) and st == int(out[i, t, 5])
            assert abs(u0 - out[i, t, 0]) <= 1e-12 and abs(d - out[i, t, 1]) <= 1e-12 and abs(bn - out[i, t, 2]) <= 1e-12
            assert abs(brd - out[i, t, 3]) <= 1e-12 * (1 + abs(out[i, t, 3]))
            nact += mask != 0
            beta, brl = bn, br
    assert nact > 30


def _wsse(c, x, y, sg):
    return float(np.sum(((np.polyval(np.asarray(c)[::-1], x) - y) / sg) ** 2))


def test_polynomial_lane_fit_vs_reference_curve_fit(refvec2):
    """PolyLane.fit_polynomial_curve (cbf/obstacles.py:715-773) runs scipy's Levenberg-Marquardt from zero
    coefficients; the restatement solves the same weighted least-squares problem directly.  It must never have a
    LARGER weighted residual than the reference's answer, and the two curves agree to the reference's own
    convergence (1e-5 m on the data range up to degree 4; LM stalls earlier on some degree-5 fits)."""
    g = refvec2
    for i in range(g["fit_k"].shape[0]):
        K, n = int(g["fit_k"][i]), int(g["fit_n"][i])
        x, y, sg = g["fit_x"][i, :K], g["fit_y"][i, :K], g["fit_sigma"][i, :K]
        c = o.fit_polynomial(x, y, n, sg)
        ref = g["fit_c"][i, : n + 1]
        assert _wsse(c, x, y, sg) <= _wsse(ref, x, y, sg) * (1 + 1e-9) + 1e-18
        if n <= 4:
            xx = np.linspace(x.min(), x.max(), 64)
            assert np.abs(np.polyval(c[::-1], xx) - np.polyval(ref[::-1], xx)).max() <= 1e-5


def test_polynomial_lane_fit_equals_numpy_polyfit():
    """Independent check of the lane-fit restatement: numpy.polyfit with weights 1 / sigma minimises the same weighted
    residual (on well-conditioned inputs -- polyfit works on the raw Vandermonde matrix)."""
    rng = np.random.default_rng(4)
    for n in (1, 2, 3):
        for _ in range(20):
            K = int(rng.integers(n + 2, 30))
            x = np.sort(rng.uniform(-5, 5, K))
            y = rng.normal(size=n + 1) @ np.vander(x, n + 1, increasing=True).T + rng.normal(0, 0.05, K)
            sg = rng.uniform(0.5, 3.0, K)
            c = o.fit_polynomial(x, y, n, sg)
            ref = np.polyfit(x, y, n, w=1.0 / sg)[::-1]
            assert np.allclose(c, ref, rtol=1e-9, atol=1e-10)


# ------------------------------------------------------------------------------------------ config 3 closed loop
@pytest.fixture(scope="module")
def radial_loop(golden_dir):
    """Produced by EXECUTING the reference's radial_dynamic_obstacles.py -- the whole module: RadialObstacleSpawner,
    single_obstacle_CBF1 and animate() for 600 frames (tests/golden/gen_reference_vectors_radial_loop.py)."""
    return np.load(os.path.join(golden_dir, "reference_vectors_radial_loop.npz"))


RADIAL_PRM = dict(model=o.MODEL_DBM, nominal=o.NOMINAL_CONST, uref0=0.0, uref1=0.0, seeker=1, dt=1.0 / 30.0, alpha=1.0)


def radial_fields(obs_row):
    """RADIAL slot fields (include/sccav_cbf.h) from a recorded (cx, cy, vx, vy, r): a = b = r, kv = 1 (rdo.py:462-463)."""
    cx, cy, vx, vy, r = obs_row
    return [cx, cy, r, r, 1.0, vx, vy, 0.0]


@pytest.mark.parametrize("tag", ["m1_s0", "m1_s1", "m1_s2"])
def test_radial_closed_loop_reproduces_reference_animate(radial_loop, tag):
    """BASELINE config 3 pinned to the reference ITSELF: the order filter -> update_com(dt = 1/30) -> update_seekers
    (rdo.py:456-487), the spawn state (seeker velocity = ego speed = 0, :186), the seeker law (:193-239).  The scalar
    oracle reproduces all 599 frames after the spawn BIT FOR BIT -- states, controls, active sets -- including the
    frames after the seeker has reached the ego and circles it."""
    ego, u, row, obs = (radial_loop[tag + "_" + k] for k in ("ego", "u", "row", "obs"))
    assert np.all(ego[0] == 0.0) and np.all(ego[1] == 0.0) and np.all(u[0] == 0.0)       # frame 0: nothing spawned yet
    r = o.rollout(ego[1], [o.SLOT_RADIAL], [radial_fields(obs[1, 0])], None, 599, params=RADIAL_PRM, record=True)
    st = np.array(r["rec"]["state"]); uu = np.array(r["rec"]["u"]); mk = np.array(r["rec"]["mask"])
    assert np.array_equal(st, ego[1:])
    assert np.array_equal(uu, u[1:])
    assert np.array_equal(mk != 0, row[1:, 3] != 0) and (mk != 0).sum() > 40
    assert np.array_equal(np.array(r["state"]), radial_loop[tag + "_final_ego"])
    fo = radial_loop[tag + "_final_obs"][0]
    assert r["fields"][0][0] == fo[0] and r["fields"][0][1] == fo[1] and r["fields"][0][5] == fo[2] and r["fields"][0][6] == fo[3]


def test_radial_closed_loop_c_oracle(radial_loop):
    """The C oracle (the checker of the GPU tests and the CPU baseline) on the same three runs, as ONE batch."""
    from oracle import c_oracle as co
    tags = ["m1_s0", "m1_s1", "m1_s2"]
    state = np.stack([radial_loop[t + "_ego"][1] for t in tags], axis=1)
    obst = np.stack([np.array(radial_fields(radial_loop[t + "_obs"][1, 0])) for t in tags], axis=1)[None]
    res = co.rollout(co.default_params(**RADIAL_PRM), [o.SLOT_RADIAL], state, np.ascontiguousarray(obst), None, 599, record_stride=1)
    for j, t in enumerate(tags):
        ego, u = radial_loop[t + "_ego"], radial_loop[t + "_u"]
        assert np.abs(res["traj"][:, 0:4, j] - ego[1:]).max() <= 1e-12
        assert np.abs(res["traj"][:, 4:6, j] - u[1:]).max() <= 1e-12
        assert np.array_equal(res["traj_mask"][:, j] != 0, radial_loop[t + "_row"][1:, 3] != 0)


def test_seeker_law_for_a_crowd(radial_loop):
    """16 seekers spawned by the reference's spawner and moved by its update_seekers for 598 frames: the oracle's
    seeker_update, fed the ego states the reference visited, reproduces every centre and velocity bit for bit."""
    ego, obs = radial_loop["m16_s0_ego"], radial_loop["m16_s0_obs"]
    assert obs.shape == (600, 16, 5) and not np.isnan(obs[1:]).any()
    d0 = np.hypot(obs[1, :, 0], obs[1, :, 1])
    assert (d0 >= 10.0).all() and (d0 <= 20.0).all() and (obs[1, :, 4] >= 1.5).all() and (obs[1, :, 4] <= 2.0).all()   # rdo.py:55-56
    assert np.all(obs[1, :, 2:4] == 0.0)                                   # spawned with the ego's speed (0), rdo.py:186
    f = [radial_fields(obs[1, j]) for j in range(16)]
    for i in range(1, 599):
        for j in range(16):
            o.seeker_update(f[j], ego[i + 1, 0], ego[i + 1, 1], 1.0 / 30.0)
            assert f[j][0] == obs[i + 1, j, 0] and f[j][1] == obs[i + 1, j, 1]
            assert f[j][5] == obs[i + 1, j, 2] and f[j][6] == obs[i + 1, j, 3]
