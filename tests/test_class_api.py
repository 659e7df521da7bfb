"""The reference-facing class API (sccav_cbf_b200.cbf / .obstacles / .controllers / .geometry / .utils):
same names, argument meaning and error behaviour as cbf/*.py of the reference.

* not-gpu tests: host logic only (mapping semantics, type / value errors, geometry, lane fit).
* gpu tests: the classes, used the way the reference's drivers use them, reproduce the golden
  vectors produced by the reference's OWN classes (tests/golden/reference_vectors.npz) and the
  oracle; scalar calls and batched calls agree bit for bit.
"""
import math
import os

import numpy as np
import pytest
import torch

from oracle import oracle as o
from sccav_cbf_b200 import (DBM_CBF_2DS, KBM_VC_CBF2D, BoundingBox, CollisionCone2D, Ellipse2D, LateralStanley,
                            Obstacle2DBase, Obstacle2DTypes, ObstacleList2D, PID1, Point2, PolyLane, Rotation, Transform,
                            Vector2, Vector3, ZERO_TOL, normalize_angle, saturation, vec_norm)
from sccav_cbf_b200.utils import Timer, TimerError, convert_LH_to_RH, get_closest_idx

gpu = pytest.mark.gpu


def rel(a, b):
    return abs(a - b) / (1.0 + abs(b))


# =========================================================================================== host logic
def test_obstacle_list_is_an_ordered_typed_mapping():
    lst = ObstacleList2D()
    with pytest.raises(TypeError):                     # cbf/obstacles.py:817-818
        lst["x"] = 3.0
    e = Ellipse2D(2.0, 1.0, Vector2(1.0, 2.0), buffer=0.5, id=7)
    c = CollisionCone2D(1.0, [0, 0, 0, 1], [5, 0, 0, 0])
    lst.update({"b": e})
    lst[3] = c
    assert len(lst) == 2 and list(lst) == ["b", 3] and lst["b"] is e
    assert e.a == 2.5 and e.b == 1.5 and e.id == 7      # buffer added to both semi-axes (obstacles.py:159-160)
    assert c.a == 2.5                                   # default cone buffer 1.5 (obstacles.py:341,357)
    assert lst.pop("b") is e and list(lst) == [3]
    lst.set_timestamp(1.5)
    assert lst.timestamp == 1.5
    assert e.type == Obstacle2DTypes.ELLIPSE2D and c.type == Obstacle2DTypes.COLLISION_CONE2D


def test_type_and_value_errors_match_the_reference():
    with pytest.raises(TypeError):                     # obstacles.py:154-155
        Ellipse2D(1.0, 1.0, center=(0.0, 0.0))
    with pytest.raises(TypeError):                     # obstacles.py:295-296
        Ellipse2D(1.0, 1.0, Vector2()).update_by_bounding_box("box")
    with pytest.raises(TypeError):
        CollisionCone2D.from_bounding_box(bbox=1)
    f = DBM_CBF_2DS()
    f.set_model_params(1.45, 1.45)
    with pytest.raises(ValueError):                    # cbf.py:177-180 (checked before anything touches the GPU)
        f.solve_cbf([0.0, 0.0])
    with pytest.raises(ValueError):                    # cbf.py:156-157
        f.set_qp_cost_weight(np.eye(3))
    with pytest.raises(ValueError):
        KBM_VC_CBF2D().solve_cbf([1.0, 0.0])
    with pytest.raises(TypeError):                     # controllers.py:106-108
        LateralStanley().control(front_coords=(1.0, 2.0))
    with pytest.raises(ValueError):
        PolyLane(np.ones(8))
    with pytest.raises(ValueError):
        PolyLane.fit_polynomial_curve([0, 1, 2], [0, 1])


def test_buffer_bookkeeping():
    e = Ellipse2D(2.0, 1.0, Vector2(), buffer=0.5)
    with pytest.warns(UserWarning):
        e.apply_buffer()
    e.remove_buffer()
    assert (e.a, e.b, e.BUFFER_FLAG) == (2.0, 1.0, False)
    e.apply_buffer()
    e.update(buffer=1.0)
    assert (e.a, e.b, e.buffer) == (3.0, 2.0, 1.0)
    e.update(b=4.0)                                     # D3: the reference writes a here
    assert e.b == 4.0 and e.a == 3.0
    c = CollisionCone2D(1.0, [0, 0, 0, 1], [5, 0, 0, 0], buffer=0.5)
    c.update(buffer=2.0)
    assert c.a == 3.0


def test_bounding_box_ingest_add_update_remove():
    lst = ObstacleList2D()
    b1 = BoundingBox(Vector3(2.0, 1.0, 0.8), Vector3(10.0, -3.0, 0.0), Rotation(yaw=0.3), velocity=4.0)
    b2 = BoundingBox(Vector3(3.0, 1.5, 0.8), Vector3(20.0, 5.0, 0.0), Rotation(yaw=-0.2))
    lst.update_by_bounding_box({11: b1, 12: b2}, obs_type=Obstacle2DTypes.ELLIPSE2D, buffer=0.5)
    assert list(lst) == [11, 12]
    e = lst[11]
    assert (e.a, e.b, e.theta, e.center.x, e.center.y, e.id) == (2.5, 1.5, 0.3, 10.0, -3.0, 11)
    b1.location = Vector3(11.0, -3.5, 0.0)
    lst.update_by_bounding_box({11: b1})                # 12 left the scene, 11 moved (obstacles.py:842-856)
    assert list(lst) == [11] and lst[11].center.x == 11.0
    cones = ObstacleList2D()
    cones.update_by_bounding_box({5: b1}, obs_type=Obstacle2DTypes.COLLISION_CONE2D, buffer=0.5)
    c = cones[5]
    assert c.a == math.hypot(2.0, 1.0) + 0.5 and list(c.s_obs) == [11.0, -3.5, 0.0, 4.0]
    assert len(b1.get_local_vertices()) == 8 and len(b1.get_world_vertices(Transform())) == 8


def test_geometry_rotation_and_transform():
    r = Rotation(yaw=0.0)
    f, u, rt = r.get_forward_vector(), r.get_up_vector(), r.get_right_vector()
    assert (f.x, f.y, f.z) == (1.0, 0.0, 0.0) and (u.x, u.y, u.z) == (0.0, 0.0, 1.0) and (rt.x, rt.y, rt.z) == (0.0, -1.0, 0.0)
    assert Rotation(yaw=0.1) != Rotation(yaw=0.2) and Rotation(0.1, 0.2, 0.3) == Rotation(0.1, 0.2, 0.3)
    r2 = Rotation.from_quaternion(*(lambda q: (q.w, q.x, q.y, q.z))(Rotation(0.2, -0.1, 0.7).get_quaternion()))
    assert abs(r2.yaw - 0.7) < 1e-12 and abs(r2.pitch + 0.1) < 1e-12 and abs(r2.roll - 0.2) < 1e-12
    # the quaternion and the matrix describe the same rotation (euclid conventions)
    rot = Rotation(0.3, -0.2, 0.9)
    t = Transform(Vector3(0.0, 0.0, 0.0), rot)
    for v in (Vector3(1.0, 0.0, 0.0), Vector3(0.0, 1.0, 0.0), Vector3(0.3, -2.0, 0.5)):
        a, b = rot.get_quaternion() * v, t.transform(v)
        assert abs(a.x - b.x) < 1e-12 and abs(a.y - b.y) < 1e-12 and abs(a.z - b.z) < 1e-12
    t = Transform(Vector3(1.0, 2.0, 3.0), Rotation(yaw=0.4))
    p = Vector3(0.5, -1.0, 2.0)
    q = t.transform_inverse(t.transform(p))
    assert abs(q.x - p.x) < 1e-12 and abs(q.y - p.y) < 1e-12 and abs(q.z - p.z) < 1e-12
    c = convert_LH_to_RH("y", Vector3(1.0, 2.0, 3.0))
    assert (c.x, c.y, c.z) == (1.0, -2.0, 3.0)
    with pytest.raises(ValueError):
        convert_LH_to_RH("w", Vector3())


def test_utils_scalar_behaviour():
    assert ZERO_TOL == 1e-3
    assert normalize_angle(3 * math.pi) == o.normalize_angle(3 * math.pi)
    assert normalize_angle(math.pi) == math.pi and normalize_angle(-math.pi) == -math.pi      # inclusive bounds
    a = torch.tensor([0.5, 3 * math.pi, -7.0, math.pi, -math.pi, 1e9, 100.0], dtype=torch.float64)
    got = normalize_angle(a)
    want = torch.tensor([o.normalize_angle(float(v)) for v in a], dtype=torch.float64)
    assert torch.allclose(got, want, rtol=0, atol=1e-9)
    assert saturation(5, -1, 2) == 2 and saturation(-5, -1, 2) == -1 and saturation(0.5, -1, 2) == 0.5
    assert vec_norm([3.0, 4.0]) == 5.0
    assert get_closest_idx(2.2, [0, 1, 2, 3]) == 2
    t = Timer(1.0)
    t.timestamp = 2.0
    with pytest.raises(TimerError):
        t.timestamp = 1.5


def test_pid1_matches_reference_recurrence():
    p = PID1(kp=1.2, kd=0.05, ki=0.3)
    p.set_dt(0.1)
    ref = o.PID1(kp=1.2, kd=0.05, ki=0.3, dt=0.1)
    x = 0.0
    for k in range(20):
        u, ur = p.control(x, 8.0), ref.control(x, 8.0)
        assert u == ur
        x += 0.1 * u
    # tensors: one controller instance steers a whole batch
    pt = PID1(kp=1.0)
    u = pt.control(torch.tensor([1.0, 2.0]), torch.tensor([3.0, 3.0]))
    assert torch.equal(u, torch.tensor([2.0, 1.0]))


def test_lane_fit_reproduces_weighted_least_squares():
    # the 4-point cubic of lane_cbf_test.py:199-200 passes exactly through its points
    x = np.array([-1.371, -0.75, 0.0, 0.333]); y = np.array([0.0, 6.938, 3.0, 1.852])
    lane = PolyLane.fit_polynomial_curve(x, y, n=3)
    assert np.allclose(np.polynomial.polynomial.polyval(x, lane.coeffs), y, atol=1e-9)
    assert lane.order == 3
    # over-determined fit with a pinned point (sigma = alpha on the fixed point, obstacles.py:752-757)
    rng = np.random.default_rng(0)
    xs = np.linspace(0, 30, 40); ys = 2.0 + 0.1 * xs - 0.004 * xs ** 2 + rng.normal(0, 0.05, xs.size)
    fit = PolyLane.fit_polynomial_curve(xs, ys, n=2, x_fixed_pts=[0.0], y_fixed_pts=[2.0], alpha=1e-4)
    assert abs(fit.coeffs[0] - 2.0) < 1e-6 and abs(fit.coeffs[1] - 0.1) < 0.01
    from scipy.optimize import curve_fit                # what the reference calls
    sig = np.append(np.full(xs.size, 10.0), 1e-4)
    ref, _ = curve_fit(lambda t, *p: np.polynomial.polynomial.polyval(t, p), np.append(xs, 0.0), np.append(ys, 2.0), np.zeros(3), sigma=sig)
    assert np.allclose(fit.coeffs, ref, rtol=1e-6, atol=1e-8)


def test_base_obstacle_is_inert():
    b = Obstacle2DBase()
    assert b.evaluate() == 0 and b.dx() == 0 and b.dy() == 0 and b.dtheta() == 0 and b.dv() == 0 and b.dt() == 0 and b.dbeta() == 0


# =========================================================================================== GPU parity
@gpu
def test_cone_class_reproduces_reference_class_vectors(refvec):
    for row, ref in zip(refvec["cone_in"], refvec["cone_out"]):
        s, so, a, beta = row[0:4], row[4:8], row[8], row[9]
        c = CollisionCone2D(a, s, so, buffer=1.5, beta=beta)
        c.update(s=s)                                     # the path DBM_CBF_2DS.update_state takes
        got = [c.f(), c.dx(), c.dy(), c.dtheta(), c.dv(), c.dt()]
        assert all(isinstance(g, float) for g in got)
        assert max(rel(g, r) for g, r in zip(got, ref)) < 1e-12, (row, got, ref)
        assert c.dbeta() == c.dtheta()                    # obstacles.py:465
    # the same 96 cases as ONE batch of tensors
    inp = torch.from_numpy(refvec["cone_in"].T.copy())
    c = CollisionCone2D(inp[8], inp[0:4], inp[4:8], buffer=1.5, beta=inp[9])
    out = torch.stack([c.f(), c.dx(), c.dy(), c.dtheta(), c.dv(), c.dt()]).cpu().numpy().T
    assert out.shape == refvec["cone_out"].shape
    assert (np.abs(out - refvec["cone_out"]) / (1 + np.abs(refvec["cone_out"]))).max() < 1e-12


@gpu
def test_ellipse_class_reproduces_reference_class_vectors(refvec):
    for row, ref in zip(refvec["ellipse_in"], refvec["ellipse_out"]):
        x, y, cx, cy, a, b, th, buf, vx, vy = row
        e = Ellipse2D(a=a, b=b, center=Vector2(cx, cy), theta=th, buffer=buf)
        e.update(s=[x, y, 0.1, 5.0])
        e.update_velocity(Vector2(vx, vy))
        got = [e.evaluate(), e.dx(), e.dy(), e.dt()]
        assert max(rel(g, r) for g, r in zip(got, ref)) < 1e-12, (row, got, ref)
        assert e.dtheta() == 0.0 and e.dv() == 0.0        # D1: the reference raises TypeError here
        assert e.f() == e.evaluate() and e.gradient()[:2] == [got[1], got[2]]
    # D2: pushing the ego state does not disturb the obstacle's own orientation / velocity
    e = Ellipse2D(3.0, 2.0, Vector2(5.0, 1.0), theta=0.4)
    e.update_velocity_by_magnitude(2.0)
    e.update_state([0.0, 0.0, 1.2, 7.0], None)
    assert e.theta == 0.4 and abs(e.vel.x - 2.0 * math.cos(0.4)) < 1e-15
    e.update_orientation(0.9)
    assert abs(e.vel.magnitude() - 2.0) < 1e-12 and abs(e.vel.y - 2.0 * math.sin(0.9)) < 1e-12


@gpu
def test_lane_class_vs_reference_newton_cg(refvec):
    for row, ref in zip(refvec["lane_in"], refvec["lane_out"]):
        x, y, c0, c1, c2, c3, buf, th, v = row
        ln = PolyLane(np.array([c0, c1, c2, c3]), s=[x, y, th, v], buffer=buf)
        got = [ln.f(), ln.dx(), ln.dy()]
        assert max(rel(g, r) for g, r in zip(got, ref[1:4])) < 1e-6      # scipy's xtol 1e-8 on the closest point
        assert ln.dtheta() == 0.0 and ln.dv() == 0.0 and ln.dt() == 0.0


@gpu
def test_dbm_class_solve_reproduces_reference_class(refvec):
    """The usage pattern of stanley_controller_ellipse.py:733-742 / the CARLA driver on the golden
    cases generated with the reference's DBM_CBF_2DS."""
    for i in range(refvec["dbm_in"].shape[0]):
        s = refvec["dbm_in"][i, 0:4]; uref = refvec["dbm_in"][i, 4:6]
        R = refvec["dbm_in"][i, 6:10].reshape(2, 2); alpha = refvec["dbm_in"][i, 10]; m = int(refvec["dbm_in"][i, 11])
        ctl = DBM_CBF_2DS(alpha=alpha)
        ctl.set_model_params(lr=1.45, lf=1.45)
        lane = False
        for j in range(m):
            t, f = int(refvec["dbm_slot"][i, j, 0]), refvec["dbm_slot"][i, j, 1:]
            if t == o.SLOT_LANE:
                ctl.obstacle_list2d["lane%d" % j] = PolyLane(f[1:5], s=s)
                lane = True
            else:
                ctl.obstacle_list2d[j] = CollisionCone2D(f[4] - 1.5, s, f[0:4])
        ctl.update_state(s=s)
        ctl.set_qp_cost_weight(R)
        info, u = ctl.solve_cbf(list(uref), return_solver=True)
        ru0, rd, _, rmask, rstatus = refvec["dbm_out"][i]
        tol = 1e-6 if lane else 1e-11
        assert u.shape == (2,) and not u.is_cuda
        assert rel(float(u[0]), ru0) <= tol and rel(float(u[1]), rd) <= tol, (i, u, ru0, rd)
        assert info["active_mask"] == int(rmask) and info["status"] == int(rstatus)
        u2 = ctl.solve_cbf(np.array(uref))                 # default: only u comes back (cbf.py:217-220)
        assert torch.equal(u2, u)


@gpu
def test_batched_solve_equals_scalar_solves_and_oracle():
    rng = np.random.default_rng(42)
    N, M = 257, 6
    s = np.stack([rng.uniform(-5, 5, N), rng.uniform(-5, 5, N), rng.uniform(-1, 1, N), rng.uniform(3, 12, N)])
    ur = np.stack([rng.uniform(-2, 2, N), rng.uniform(-0.4, 0.4, N)])
    ctl = DBM_CBF_2DS(alpha=1.0)
    ctl.set_model_params(1.45, 1.45)
    fields = []
    for j in range(M):
        cx = s[0] + rng.uniform(4, 25, N) * np.cos(s[2]); cy = s[1] + rng.uniform(4, 25, N) * np.sin(s[2]) + rng.uniform(-3, 3, N)
        a, b, th = rng.uniform(2, 5, N), rng.uniform(1, 3, N), rng.uniform(-3, 3, N)
        e = Ellipse2D(torch.from_numpy(a), torch.from_numpy(b), Vector2(torch.from_numpy(cx), torch.from_numpy(cy)),
                      theta=torch.from_numpy(th), buffer=0.5)
        ctl.obstacle_list2d[j] = e
        fields.append(np.stack([cx, cy, a + 0.5, b + 0.5, th, 0 * a, 0 * a, 0 * a]))
    ctl.update_state(torch.from_numpy(s))
    # default: unchanged ellipses are evaluated in their ingested form (a few ulp from the reference's operation order)
    assert ctl.obstacle_list2d.use_prepared
    info_p, u_p = ctl.solve_cbf(torch.from_numpy(ur), return_solver=True)
    _, u_p2 = ctl.solve_cbf(torch.from_numpy(ur), return_solver=True)          # second call: the cached image, nothing re-packed
    assert torch.equal(u_p, u_p2)
    # the reference's operation order on every call: bit-identical to the one-scenario calls below
    ctl.obstacle_list2d.use_prepared = False
    info, u = ctl.solve_cbf(torch.from_numpy(ur), return_solver=True)
    assert u.is_cuda and u.shape == (2, N)
    assert ((u_p - u).abs() <= 1e-12 * (1 + u.abs())).all()
    assert torch.equal(info_p["active_mask"], info["active_mask"]) and torch.equal(info_p["status"], info["status"])
    u = u.cpu().numpy(); mask = info["active_mask"].cpu().numpy().view(np.uint32); status = info["status"].cpu().numpy()
    assert (status != o.STATUS_INACTIVE).sum() > 10
    for n in range(0, N, 7):
        f = [list(fields[j][:, n]) for j in range(M)]
        u0, d, mk, st, _, _ = o.filter_step(o.MODEL_DBM, list(s[:, n]), list(ur[:, n]), [o.SLOT_ELLIPSE] * M, f, 1.0, 1.45, 1.45, 2.9,
                                            (1.0, 0.0, 0.0, 1.0))
        assert rel(u[0, n], u0) < 1e-9 and rel(u[1, n], d) < 1e-9 and mask[n] == mk and status[n] == st
        # the same vehicle through the scalar API
        one = DBM_CBF_2DS(alpha=1.0)
        one.set_model_params(1.45, 1.45)
        for j in range(M):
            one.obstacle_list2d[j] = Ellipse2D(f[j][2] - 0.5, f[j][3] - 0.5, Vector2(f[j][0], f[j][1]), theta=f[j][4], buffer=0.5)
        one.update_state(list(s[:, n]))
        us = one.solve_cbf(list(ur[:, n]))
        assert float(us[0]) == u[0, n] and float(us[1]) == u[1, n]
    # stacked getters of the list: [M, N]
    h = ctl.obstacle_list2d.f()
    assert h.shape == (M, N) and torch.equal(h[2], ctl.obstacle_list2d[2].f())
    assert ctl.obstacle_list2d.gradient().shape == (M, 4, N)


@gpu
def test_kbm_class_vs_oracle():
    """KBM_VC_CBF2D used as stanley_controller_ellipse.py:702-711 intends (== CBF(), :214-238)."""
    rng = np.random.default_rng(3)
    for _ in range(40):
        x, y, th = rng.uniform(-5, 5), rng.uniform(-5, 5), rng.uniform(-1, 1)
        cx, cy = x + rng.uniform(3, 20) * math.cos(th), y + rng.uniform(3, 20) * math.sin(th) + rng.uniform(-2, 2)
        ctl = KBM_VC_CBF2D(alpha=1.0)
        ctl.set_model_params(L=2.9)
        ctl.obstacle_list2d.update({0: Ellipse2D(6.0, 3.0, Vector2(cx, cy))})
        ctl.update_state(Point2(x, y), th)
        uref = [rng.uniform(4, 10), rng.uniform(-0.3, 0.3)]
        info, u = ctl.solve_cbf(uref)
        u0, d, mk, st, _, _ = o.filter_step(o.MODEL_KBM, [x, y, th, 0.0], uref, [o.SLOT_ELLIPSE], [[cx, cy, 6.0, 3.0, 0, 0, 0, 0]],
                                            1.0, 1.45, 1.45, 2.9, (1.0, 0.0, 0.0, 1.0))
        assert rel(float(u[0]), u0) < 1e-10 and rel(float(u[1]), d) < 1e-10 and info["active_mask"] == mk and info["status"] == st


@gpu
def test_lateral_stanley_class_vs_oracle():
    from sccav_cbf_b200.course import config1_course
    cx, cy, cyaw = config1_course()
    traj = [(cx[i], cy[i], cyaw[i], 8.0) for i in range(len(cx))]
    rng = np.random.default_rng(9)
    ctl = LateralStanley(lr=1.45, lf=1.45, k=0.5, ks=0.01)
    ctl.set_trajectory(traj)
    last = 0
    for step in range(60):                               # a vehicle progressing along the course, scalar API
        i = min(len(cx) - 1, 30 * step + int(rng.integers(0, 10)))
        x, y = cx[i] + rng.normal(0, 1.0), cy[i] + rng.normal(0, 1.0)
        yaw, v = cyaw[i] + rng.normal(0, 0.2), rng.uniform(2, 10)
        ctl.update_state(x, y, yaw, v)
        delta, idx = ctl.control()
        d_ref, i_ref = o.stanley_control(x, y, yaw, v, cx, cy, cyaw, last, 0.5, 1.45, ks=0.01)
        assert idx == i_ref and rel(delta, d_ref) < 1e-10
        last = i_ref
    # batch: N vehicles, two consecutive calls (the second exercises the monotone clamp), external front axle
    N = 4096
    base = rng.integers(0, len(cx), N)
    st = np.stack([cx[base] + rng.normal(0, 2, N), cy[base] + rng.normal(0, 2, N), cyaw[base] + rng.normal(0, 0.3, N), rng.uniform(1, 12, N)])
    b = LateralStanley(lr=1.45, lf=1.45, k=0.5, ks=0.0)
    b.set_trajectory(traj)
    b.update_state(*[torch.from_numpy(r) for r in st])
    d1, i1 = b.control()
    st2 = st.copy(); st2[0] -= 3.0 * np.cos(st[2]); st2[1] -= 3.0 * np.sin(st[2])      # moved backwards: index must not decrease
    b.update_state(*[torch.from_numpy(r) for r in st2])
    d2, i2 = b.control()
    i1, i2, d1, d2 = i1.cpu().numpy(), i2.cpu().numpy(), d1.cpu().numpy(), d2.cpu().numpy()
    assert (i2 >= i1).all()
    for n in range(0, N, 41):
        dr, ir = o.stanley_control(*st[:, n], cx, cy, cyaw, 0, 0.5, 1.45)
        assert i1[n] == ir and rel(d1[n], dr) < 1e-10
        dr2, ir2 = o.stanley_control(*st2[:, n], cx, cy, cyaw, ir, 0.5, 1.45)
        assert i2[n] == ir2 and rel(d2[n], dr2) < 1e-10
    f = LateralStanley(lr=1.45, lf=1.45, k=0.5, ks=0.0)
    f.set_trajectory(traj)
    f.update_state(10.0, 1.0, 0.1, 5.0)
    d, i = f.control(front_coords=Vector2(10.0 + 1.45 * math.cos(0.1), 1.0 + 1.45 * math.sin(0.1)))
    dr, ir = o.stanley_control(10.0, 1.0, 0.1, 5.0, cx, cy, cyaw, 0, 0.5, 1.45)
    assert i == ir and rel(d, dr) < 1e-12


@gpu
def test_reference_driver_loop_with_the_class_api_reproduces_beta_vs_time(golden_dir):
    """stanley_controller_ellipse.py main() with CBF_TYPE = 4 (:717-750), written against THIS package's
    classes exactly as the reference writes it against cbf/*: a fresh DBM_CBF_2DS + CollisionCone2D every
    tick, update_state, set_qp_cost_weight(diag(.5,.5)), solve_cbf.  Must reproduce beta_vs_time.mat."""
    import json
    import os
    from sccav_cbf_b200.course import config1_course
    gold = json.load(open(os.path.join(golden_dir, "beta_vs_time.json")))
    cx, cy, cyaw = config1_course()
    k, Kp, dt, L, lr, lf, max_steer = 0.5, 1.0, 0.1, 2.9, 1.45, 1.45, np.radians(30.0)
    x, y, yaw, v = -0.0, 5.0, np.radians(20.0), 10.0
    target_speed = 30.0 / 3.6
    last_idx = len(cx) - 1
    oi = int(last_idx * 0.75)
    a_cone = np.hypot(20, 10) / 2
    target_idx, _ = o.calc_target_index(x, y, yaw, cx, cy, L)
    time, betas = 0.0, [0.0]
    while 30 >= time and last_idx > target_idx:
        a_ = Kp * (target_speed - v)
        di, target_idx = o.stanley_control(x, y, yaw, v, cx, cy, cyaw, target_idx, k, L)
        s = [x, y, yaw, v]
        cbf = DBM_CBF_2DS(alpha=1)
        cbf.set_model_params(lr=lr, lf=lf)
        cbf.obstacle_list2d.update({0: CollisionCone2D(a=a_cone, s=s, s_obs=[cx[oi], cy[oi], 0, 0])})
        cbf.update_state(s)
        cbf.set_qp_cost_weight(np.diag([0.5, 0.5]))
        u = cbf.solve_cbf(np.array([a_, di]))
        (x, y, yaw, v), beta = o.plant_update_com([x, y, yaw, v], float(u[0]), float(u[1]), dt, lr, lf, max_steer)
        betas.append(beta)
        time += dt
    assert len(betas) == 277
    assert np.abs(np.degrees(betas) - np.array(gold["beta_deg"])).max() <= 1e-3


@pytest.fixture(scope="module")
def refvec2(golden_dir):
    return np.load(os.path.join(golden_dir, "reference_vectors_lane_sadbm.npz"))


@gpu
def test_sadbm_class_sequences_vs_reference_class(refvec2):
    """SADBM_CBF_2DS used the way the reference's object is (cbf/cbf.py:300-437): 16 sequences of 6 ticks with
    1-3 collision cones, fixed dt; controls, the carried beta and the active sets against vectors produced by
    the reference's own class (tests/golden/gen_reference_vectors_lane_sadbm.py)."""
    from sccav_cbf_b200 import SADBM_CBF_2DS
    cfg, M, cone = refvec2["sadbm_cfg"], refvec2["sadbm_m"], refvec2["sadbm_cone"]
    I, out = refvec2["sadbm_in"], refvec2["sadbm_out"]
    for i in range(cfg.shape[0]):
        alpha, dt, lr, lf = cfg[i, :4]
        m = int(M[i])
        ctl = SADBM_CBF_2DS(alpha=alpha, dt=dt)
        ctl.set_model_params(lr=lr, lf=lf)
        ctl.set_qp_cost_weight(cfg[i, 4:8].reshape(2, 2))
        s0 = I[i, 0, :4]
        obs = []
        for j in range(m):
            c = cone[i, 0, j]
            ob = CollisionCone2D(c[4] - 1.5, s0, c[:4])                       # default buffer 1.5 (obstacles.py:341)
            ctl.obstacle_list2d[j] = ob
            obs.append(ob)
        for t in range(I.shape[1]):
            s, ur = I[i, t, :4], I[i, t, 4:6]
            for j, ob in enumerate(obs):
                ob.update(s_obs=cone[i, t, j, :4])
            ctl.update_state(s=s)
            info, u = ctl.solve_cbf(list(ur), return_solver=True)
            assert info["active_mask"] == int(out[i, t, 4]) and info["status"] == int(out[i, t, 5]), (i, t)
            assert abs(float(u[0]) - out[i, t, 0]) <= 1e-9 * (1 + abs(out[i, t, 0]))
            assert abs(float(u[1]) - out[i, t, 1]) <= 1e-9
            assert abs(float(ctl._beta[0]) - out[i, t, 2]) <= 1e-10
    with pytest.raises(ValueError):
        SADBM_CBF_2DS().solve_cbf([0.0, 0.0])


@gpu
def test_sadbm_batched_ticks_vs_oracle():
    """model SADBM through the C-ABI on a batch: three consecutive ticks with the carried state (beta, last
    beta_ref) updated in place, cones + an ellipse + a lane, per-vehicle counts; against oracle.sadbm_filter_step."""
    from oracle import oracle as o
    from sccav_cbf_b200 import ops
    from tests import helpers as H
    N = 1536
    slots = [o.SLOT_CONE, o.SLOT_CONE, o.SLOT_ELLIPSE, o.SLOT_CONE, o.SLOT_LANE]
    rng = np.random.default_rng(99)
    R = [1.0, 0.2, 0.2, 2.0]
    dt = 0.02
    prm = ops.make_params(model=4, R=R, alpha=0.9, sadbm_dt=dt)
    dev = torch.device("cuda", 0)
    aug = torch.zeros((2, N), dtype=torch.float64, device=dev)
    aug_ref = np.zeros((2, N))
    nact = 0
    for tick in range(3):
        s = H.random_states(rng, N)
        ob = H.random_slots(rng, N, slots, s)
        ur = H.random_uref(rng, N)
        u, mask, status, hmin = ops.filter_step(prm, slots, torch.from_numpy(s).to(dev), torch.from_numpy(ob).to(dev),
                                                torch.from_numpy(ur).to(dev), aug=aug)
        u_, m_, st_, a_ = u.cpu().numpy(), mask.cpu().numpy().view(np.uint32), status.cpu().numpy(), aug.cpu().numpy()
        for n in range(N):
            fields = [list(ob[k, :, n]) for k in range(len(slots))]
            u0, d, bn, br, mk, st, _, _ = o.sadbm_filter_step(list(s[:, n]), list(ur[:, n]), aug_ref[0, n], aug_ref[1, n], dt, slots,
                                                              fields, 0.9, 1.45, 1.45, R)
            assert mk == int(m_[n]) and st == int(st_[n]), (tick, n)
            assert abs(u0 - u_[0, n]) <= 1e-9 * (1 + abs(u0)) and abs(d - u_[1, n]) <= 1e-9 * (1 + abs(d))
            assert abs(bn - a_[0, n]) <= 1e-9 * (1 + abs(bn)) and abs(br - a_[1, n]) <= 1e-12
            aug_ref[0, n], aug_ref[1, n] = bn, br
            nact += mk != 0
    assert nact > 100
    # stateful model: the carried state is mandatory, and there is no closed loop for it
    with pytest.raises(Exception):
        ops.filter_step(prm, slots, torch.from_numpy(s).to(dev), torch.from_numpy(ob).to(dev), torch.from_numpy(ur).to(dev))


@gpu
def test_sadbm_wall_clock_mode_and_lane_distance_form():
    """SADBM_CBF_2DS(dt=None) measures the step on the host as cbf.py:363-365 does (floored at ZERO_TOL) and still
    integrates beta consistently; PolyLane(distance_form=True) is the CBF_lane_sqrt barrier
    (stanley_controller_ellipse.py:465-512) -- class getters and solve_cbf against the oracle."""
    import time
    from sccav_cbf_b200 import SADBM_CBF_2DS
    from sccav_cbf_b200.utils import ZERO_TOL
    ctl = SADBM_CBF_2DS(alpha=1.0, dt=None)
    ctl.set_model_params(lr=1.45, lf=1.45)
    s = [0.0, 0.0, 0.1, 8.0]
    ctl.obstacle_list2d[0] = CollisionCone2D(2.0, s, [18.0, 1.0, 3.0, 2.0])
    ctl.update_state(s=s)
    u1 = ctl.solve_cbf([0.5, 0.05])
    b1 = float(ctl._beta[0])
    assert ctl._dt >= ZERO_TOL and math.isfinite(float(u1[1]))
    time.sleep(0.02)
    ctl.update_state(s=s)
    u2 = ctl.solve_cbf([0.5, 0.05])
    assert ctl._dt >= 0.02                                          # the measured step went into the kernel
    # same reference twice: beta_ref_dot = 0 on the second call, so beta moves by u[1]-rate * dt only
    assert abs(float(ctl.beta_ref_last[0]) - math.atan2(1.45 * math.tan(0.05), 2.9)) < 1e-15
    assert math.isfinite(b1) and math.isfinite(float(ctl._beta[0]))
    # distance-form lane barrier
    co = np.array([3.0, 0.02, 0.001])
    st = [10.0, 1.2, 0.3, 7.0]
    ln = PolyLane(co, s=st, buffer=1.5, distance_form=True)
    h, hx, hy, _, _, _ = o.lane_sqrt_partials(st[0], st[1], list(co) + [0.0, 0.0, 0.0], 1.5)
    assert rel(ln.f(), h) < 1e-9 and rel(ln.dx(), hx) < 1e-9 and rel(ln.dy(), hy) < 1e-9
    sq = PolyLane(co, s=st, buffer=1.5)
    assert abs((ln.f() + 1.5) ** 2 - (sq.f() + 1.5)) < 1e-9           # sqrt(d^2) - buffer  vs  d^2 - buffer
    d = DBM_CBF_2DS(alpha=1.0)
    d.set_model_params(lr=1.45, lf=1.45)
    d.obstacle_list2d["l"] = ln
    d.update_state(st)
    info, u = d.solve_cbf([0.0, 0.3], return_solver=True)
    ref = o.filter_step(o.MODEL_DBM, st, [0.0, 0.3], [o.SLOT_LANE_SQRT], [[1.5] + list(co) + [0.0, 0.0, 0.0, 0.0]], 1.0, 1.45, 1.45, 2.9, (1, 0, 0, 1))
    assert info["status"] == ref[3] and abs(float(u[1]) - ref[1]) < 1e-7 and abs(float(u[0]) - ref[0]) < 1e-9


@gpu
def test_pack_cache_follows_every_change_of_the_list():
    """ObstacleList2D.pack caches the packed / ingested obstacle image; any change of any field -- a method call, a
    plain attribute assignment, an in-place tensor update, an obstacle added or removed -- must be seen."""
    rng = np.random.default_rng(3)
    N = 64
    s = torch.from_numpy(np.stack([rng.uniform(-5, 5, N), rng.uniform(-5, 5, N), rng.uniform(-1, 1, N), rng.uniform(3, 12, N)]))
    ur = torch.from_numpy(np.stack([rng.uniform(-2, 2, N), rng.uniform(-0.4, 0.4, N)]))

    def build(cx, cy, a, th, vel=None):
        c = DBM_CBF_2DS(alpha=1.0)
        c.set_model_params(1.45, 1.45)
        for j in range(3):
            e = Ellipse2D(a[j], 1.5, Vector2(cx[j].clone(), cy[j].clone()), theta=th[j], buffer=0.5)
            if vel is not None:
                e.update_velocity(vel[j])
            c.obstacle_list2d[j] = e
        c.update_state(s)
        return c

    cx = [torch.from_numpy(rng.uniform(5, 20, N)) for _ in range(3)]
    cy = [torch.from_numpy(rng.uniform(-4, 4, N)) for _ in range(3)]
    a, th = [3.0, 4.0, 2.5], [0.1, -0.7, 1.3]
    ctl = build(cx, cy, a, th)
    u0 = ctl.solve_cbf(ur)
    assert torch.equal(u0, ctl.solve_cbf(ur))
    # 1. plain attribute assignment of a scalar field
    ctl.obstacle_list2d[1].theta = 0.4
    th2 = [0.1, 0.4, 1.3]
    assert torch.equal(ctl.solve_cbf(ur), build(cx, cy, a, th2).solve_cbf(ur))
    # 2. in-place update of a tensor field
    ctl.obstacle_list2d[0].center.x.add_(1.25)
    cx2 = [cx[0] + 1.25, cx[1], cx[2]]
    assert torch.equal(ctl.solve_cbf(ur), build(cx2, cy, a, th2).solve_cbf(ur))
    # 3. a method of the reference API, and a moving obstacle (no longer STATIC)
    ctl.obstacle_list2d[2].update_velocity(Vector2(0.5, -0.25))
    ref = build(cx2, cy, a, th2, vel=[Vector2(0, 0), Vector2(0, 0), Vector2(0.5, -0.25)])
    assert torch.equal(ctl.solve_cbf(ur), ref.solve_cbf(ur))
    # 4. an obstacle leaves
    ctl.obstacle_list2d.pop(1)
    ref.obstacle_list2d.pop(1)
    assert torch.equal(ctl.solve_cbf(ur), ref.solve_cbf(ur))
    assert not torch.equal(u0, ctl.solve_cbf(ur))
