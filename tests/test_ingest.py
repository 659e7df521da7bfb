"""Obstacle ingest from bounding boxes: the batched ObstacleList2D.update_by_bounding_box.

CPU: the oracle's slot-array restatement against the reference's literal dict semantics
(cbf/obstacles.py:833-858, written out here on plain dicts).  GPU: the ingest kernel (KB) against the
oracle -- ids / order / counts bit-exact -- and the filter on lists of different length (per-vehicle count)."""
import numpy as np
import pytest
import torch

from oracle import c_oracle as co
from oracle import oracle as o
from tests import helpers as H


def random_scene(rng, N, M, K, n_ticks, p_leave=0.3, id_pool=40):
    """Per vehicle a population of actor ids drifting in and out of range: tick t -> (box_id [K,N], box [K,6,N])."""
    ticks = []
    cur = [list(rng.choice(id_pool, size=rng.integers(0, K + 1), replace=False)) for _ in range(N)]
    for _ in range(n_ticks):
        bid = -np.ones((K, N), np.int32)
        box = np.zeros((K, 6, N))
        for n in range(N):
            keep = [i for i in cur[n] if rng.uniform() > p_leave]
            new = [i for i in rng.permutation(id_pool) if i not in keep][: rng.integers(0, K - len(keep) + 1)]
            ids = list(rng.permutation(keep + [int(i) for i in new]))
            cur[n] = ids
            pos = rng.permutation(K)[: len(ids)]                  # boxes sit at arbitrary positions, padding in between
            for k, i in zip(sorted(pos), ids):
                bid[k, n] = i
                box[k, :, n] = [rng.uniform(1.5, 3), rng.uniform(0.8, 1.5), rng.uniform(-30, 120), rng.uniform(-40, 10),
                                rng.uniform(-3, 3), rng.uniform(0, 12)]
        ticks.append((bid, box))
    return ticks


def reference_dict_update(mapping, bbox_dict, obs_type, buffer):
    """cbf/obstacles.py:833-858 on a plain dict id -> fields (Python dicts keep insertion order)."""
    for key, bbox in bbox_dict.items():
        if key in mapping:
            mapping[key] = o.box_to_fields(obs_type, False, buffer, bbox, mapping[key])        # .update_by_bounding_box
        else:
            mapping[key] = o.box_to_fields(obs_type, True, buffer, bbox)                        # .from_bounding_box
    for key in list(mapping.keys()):
        if key not in list(bbox_dict.keys()):
            mapping.pop(key)


@pytest.mark.parametrize("obs_type", [o.SLOT_ELLIPSE, o.SLOT_CONE])
def test_oracle_ingest_equals_reference_dict_semantics(obs_type):
    rng = np.random.default_rng(3 + obs_type)
    N, M, K = 40, 12, 12                                          # M >= K: nothing is ever dropped
    ticks = random_scene(rng, N, M, K, 8)
    for n in range(N):
        mapping = {}
        ids, fields, cnt = [-1] * M, [[0.0] * 8 for _ in range(M)], 0
        for bid, box in ticks:
            bbox_dict = {int(bid[k, n]): box[k, :, n] for k in range(K) if bid[k, n] >= 0}
            reference_dict_update(mapping, bbox_dict, obs_type, 0.5)
            ids, fields, cnt, dropped = o.ingest_boxes(obs_type, o.INGEST_UPDATE, 0.5, M, bid[:, n], box[:, :, n].T.reshape(K, 6) if False else [box[k, :, n] for k in range(K)], ids, fields, cnt)
            assert dropped == 0 and cnt == len(mapping)
            assert ids[:cnt] == list(mapping.keys())              # dict order == constraint order
            assert all(i == -1 for i in ids[cnt:])
            for m, key in enumerate(mapping):
                assert fields[m] == mapping[key]


def test_oracle_ingest_capacity_and_rebuild():
    box = [np.array([2.0, 1.0, 10.0 + k, -3.0, 0.1 * k, 4.0]) for k in range(5)]
    ids, f, cnt, dropped = o.ingest_boxes(o.SLOT_CONE, o.INGEST_UPDATE, 1.5, 3, [7, -1, 9, 4, 2], box, [-1] * 3, [[0.0] * 8] * 3, 0)
    assert ids == [7, 9, 4] and cnt == 3 and dropped == 1
    assert f[0][4] == float(np.hypot(2.0, 1.0)) + 1.5 and f[0][3] == 4.0 and f[0][2] == 0.0
    # update keeps the order and drops the buffer (obstacles.py:528: self.a = hypot(...)); rebuild re-applies it
    ids2, f2, cnt2, _ = o.ingest_boxes(o.SLOT_CONE, o.INGEST_UPDATE, 1.5, 3, [4, 7, -1, -1, -1], box, ids, f, cnt)
    assert ids2 == [7, 4, -1] and cnt2 == 2 and f2[0][4] == float(np.hypot(2.0, 1.0))
    ids3, f3, cnt3, _ = o.ingest_boxes(o.SLOT_CONE, o.INGEST_REBUILD, 1.5, 3, [4, 7, -1, -1, -1], box, ids2, f2, cnt2)
    assert ids3 == [4, 7, -1] and cnt3 == 2 and f3[0][4] == float(np.hypot(2.0, 1.0)) + 1.5


def test_c_oracle_per_vehicle_count():
    """count[n] obstacles per vehicle: the C oracle on M slots with a count equals itself on the truncated list."""
    rng = np.random.default_rng(8)
    N, M = 64, 6
    slots = [o.SLOT_ELLIPSE] * M
    s = H.random_states(rng, N); ob = H.random_slots(rng, N, slots, s); ur = H.random_uref(rng, N)
    count = rng.integers(0, M + 1, N).astype(np.int32)
    full = co.filter_step(co.default_params(), slots, s, ob, ur, rows=True, count=count)
    for n in range(N):
        c = int(count[n])
        if c == 0:
            assert np.array_equal(full["u"][:, n], ur[:, n]) and full["mask"][n] == 0 and full["status"][n] == 0
            continue
        one = co.filter_step(co.default_params(), slots[:c], s[:, n:n + 1], ob[:c, :, n:n + 1], ur[:, n:n + 1], rows=True)
        assert np.array_equal(full["u"][:, n], one["u"][:, 0]) and full["mask"][n] == one["mask"][0]
        assert np.array_equal(full["b"][:c, n], one["b"][:, 0]) and np.all(full["b"][c:, n] == -np.inf)


# ------------------------------------------------------------------------------------------------ GPU
def T(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(torch.device("cuda", 0))


@pytest.mark.gpu
@pytest.mark.parametrize("obs_type,M,K", [(o.SLOT_ELLIPSE, 8, 8), (o.SLOT_CONE, 5, 12), (o.SLOT_CONE, 32, 32)])
def test_ingest_kernel_vs_oracle_over_ticks(obs_type, M, K):
    from sccav_cbf_b200 import ops
    rng = np.random.default_rng(100 + M)
    N = 700
    ticks = random_scene(rng, N, M, K, 6, id_pool=3 * K)
    ids_g = torch.full((M, N), -1, dtype=torch.int32, device="cuda")
    ob_g = torch.zeros((M, 8, N), dtype=torch.float64, device="cuda")
    cnt_g = torch.zeros(N, dtype=torch.int32, device="cuda")
    drop_g = torch.zeros(N, dtype=torch.int32, device="cuda")
    ids_o = [[-1] * M for _ in range(N)]
    f_o = [[[0.0] * 8 for _ in range(M)] for _ in range(N)]
    cnt_o = [0] * N
    any_drop = 0
    for t, (bid, box) in enumerate(ticks):
        mode = o.INGEST_REBUILD if t == 3 else o.INGEST_UPDATE
        ops.ingest_boxes(obs_type, T(bid), T(box), ids_g, ob_g, cnt_g, buffer=0.5, mode=mode, dropped=drop_g)
        ig, og, cg, dg = ids_g.cpu().numpy(), ob_g.cpu().numpy(), cnt_g.cpu().numpy(), drop_g.cpu().numpy()
        for n in range(N):
            ids_o[n], f_o[n], cnt_o[n], d = o.ingest_boxes(obs_type, mode, 0.5, M, bid[:, n], [box[k, :, n] for k in range(K)],
                                                          ids_o[n], f_o[n], cnt_o[n])
            assert cg[n] == cnt_o[n] and dg[n] == d, (t, n)
            assert list(ig[:, n]) == ids_o[n], (t, n)                                  # ids and their order: exact
            c = cnt_o[n]
            ref = np.array(f_o[n][:c]).reshape(c, 8)
            assert np.allclose(og[:c, :, n], ref, rtol=1e-15, atol=0.0), (t, n)         # hypot: <= 1 ulp
            any_drop += d
    if M < K:
        assert any_drop > 0


@pytest.mark.gpu
def test_filter_with_per_vehicle_count_and_batched_list_class():
    from sccav_cbf_b200 import DBM_CBF_2DS, BatchedObstacleList2D, Obstacle2DTypes, ops
    rng = np.random.default_rng(21)
    N, M, K = 3000, 8, 8
    s = H.random_states(rng, N)
    ur = H.random_uref(rng, N)
    # boxes placed ahead of each vehicle so that a good share of the rows is active
    near = H.random_slots(rng, N, [o.SLOT_CONE] * K, s)
    bid = np.where(rng.uniform(size=(K, N)) < 0.6, rng.permuted(np.tile(np.arange(K, dtype=np.int32)[:, None], (1, N)), axis=0), -1).astype(np.int32)
    box = np.zeros((K, 6, N))
    box[:, 0], box[:, 1] = rng.uniform(1.5, 3, (K, N)), rng.uniform(0.8, 1.5, (K, N))
    box[:, 2], box[:, 3] = near[:, 0], near[:, 1]
    box[:, 5] = rng.uniform(0, 6, (K, N))
    lst = BatchedObstacleList2D(N, capacity=M, obs_type=Obstacle2DTypes.COLLISION_CONE2D)
    lst.update_by_bounding_box(T(bid), T(box), buffer=1.5)
    cnt = lst.count.cpu().numpy()
    assert cnt.min() == 0 and cnt.max() >= 6                      # lists of different length, some empty
    cbf = DBM_CBF_2DS(alpha=1.0)
    cbf.obstacle_list2d = lst
    cbf.set_model_params(lr=1.45, lf=1.45)
    cbf.update_state(T(s))
    info, u = cbf.solve_cbf(T(ur), return_solver=True)
    ref = co.filter_step(co.default_params(), lst.slot_desc, s, lst.obst.cpu().numpy(), ur, count=cnt, rows=True)
    assert np.array_equal(info["active_mask"].cpu().numpy().view(np.uint32), ref["mask"])
    assert np.array_equal(info["status"].cpu().numpy(), ref["status"])
    assert np.abs(u.cpu().numpy() - ref["u"]).max() <= 1e-9 * (1 + np.abs(ref["u"]).max())
    assert np.array_equal(u.cpu().numpy()[:, cnt == 0], ur[:, cnt == 0])          # no obstacle: u = u_ref exactly
    assert (ref["mask"] != 0).mean() > 0.02
    # rows / QP entry points honour the count as well
    prm = ops.make_params()
    A, b, h = ops.barrier_rows(prm, lst.slot_desc, T(s), lst.obst, count=lst.count)
    assert np.allclose(A.cpu().numpy(), ref["A"], rtol=1e-9, atol=1e-12)
    bb = b.cpu().numpy()
    for m in range(M):
        assert np.all(bb[m, cnt <= m] == -np.inf)
    r = T(np.stack([ur[0], np.arctan2(1.45 * np.tan(ur[1]), 2.9)]))
    for warp in (False, True):
        u2, mask2, status2 = ops.qp2_solve(prm, A, b, r, warp_per_problem=warp, count=lst.count)
        assert np.array_equal(mask2.cpu().numpy().view(np.uint32), ref["mask"]) and np.array_equal(status2.cpu().numpy(), ref["status"])
    # closed loop with lists of different length
    from sccav_cbf_b200 import scenarios as sc
    b2 = sc.config2(n_total=65536, M=8, T=300, lo=0, hi=512)
    c2 = rng.integers(0, 9, b2.N).astype(np.int32)
    g = ops.rollout(ops.make_params(), b2.slot_desc, T(b2.state), T(b2.obst), tuple(T(c) for c in b2.course), b2.T, count=T(c2))
    r2 = co.rollout(co.default_params(), b2.slot_desc, b2.state, b2.obst, b2.course, b2.T, count=c2)
    same = (g["n_active"].cpu().numpy() == r2["n_active"]) & (g["target_idx"].cpu().numpy() == r2["target_idx"])
    assert same.mean() >= 0.995
    assert (g["n_active"].cpu().numpy()[c2 == 0] == 0).all()


def test_oracle_actuator_shaping_follows_the_driver():
    # one braking episode followed by throttle: the driver's brake variable is never reset (:958-968)
    thr_p = brk_p = 0.0
    seq = []
    for ua in (2.0, 2.0, -3.0, -3.0, 0.5):
        thr, brk, st = o.actuator_shaping(ua, 0.3, thr_p, brk_p, max_steer=1.0)
        seq.append((thr, brk))
        thr_p, brk_p = thr, brk
    assert seq[0] == (0.1, 0.0) and seq[1] == (0.2, 0.0)                # rate limit 0.1 per tick
    assert seq[2] == (0.0, 0.1) and seq[3] == (0.0, 0.2)
    assert seq[4][0] == 0.1 and seq[4][1] == 0.2                          # stale brake kept while throttling
    assert o.actuator_shaping(0.5, 0.3, 0.0, 0.2, reset_brake=True)[1] == 0.0
    assert o.actuator_shaping(0.0, 1.7, 0.0, 0.0)[2] == 1.0 and o.actuator_shaping(0.0, -1.7, 0.0, 0.0)[2] == -1.0


@pytest.mark.gpu
@pytest.mark.parametrize("reset", [False, True])
def test_actuator_kernel_vs_oracle_over_ticks(reset):
    from sccav_cbf_b200 import ops
    rng = np.random.default_rng(4)
    N = 5000
    tp = torch.zeros(N, dtype=torch.float64, device="cuda")
    bp = torch.zeros(N, dtype=torch.float64, device="cuda")
    tp_o = np.zeros(N); bp_o = np.zeros(N)
    for t in range(6):
        u = np.stack([rng.normal(0, 1.5, N), rng.normal(0, 0.8, N)])
        u[0, :10] = 0.0
        thr, brk, st = ops.actuator_shaping(T(u), tp, bp, max_steer=0.6, rate=0.1, reset_brake=reset)
        ref = np.array([o.actuator_shaping(u[0, n], u[1, n], tp_o[n], bp_o[n], 0.6, 0.1, reset) for n in range(N)]).T
        tp_o, bp_o = ref[0], ref[1]
        assert np.allclose(thr.cpu().numpy(), ref[0], rtol=1e-14, atol=1e-16)     # tanh: CUDA vs numpy, <= 2 ulp
        assert np.allclose(brk.cpu().numpy(), ref[1], rtol=1e-14, atol=1e-16)
        assert np.array_equal(st.cpu().numpy(), ref[2])                            # clamp: exact
        assert np.allclose(tp.cpu().numpy(), tp_o, rtol=1e-14, atol=1e-16) and np.allclose(bp.cpu().numpy(), bp_o, rtol=1e-14, atol=1e-16)
        tp_o, bp_o = tp.cpu().numpy(), bp.cpu().numpy()                            # teacher-force the previous values


@pytest.mark.gpu
def test_device_course_generation_vs_reference_planner_restatement():
    """KC against sccav_cbf_b200.course.spline_course (bit-exact with the reference's planner on its own
    way-points, tests/test_course*.py): same number of samples, points / yaw to 1e-12, and closed loops on
    device-generated roads (one rollout launch per road) against the oracle on the same roads."""
    from sccav_cbf_b200 import ops, scenarios as sc
    from sccav_cbf_b200.course import CONFIG1_WAYPOINTS, spline_course
    rng = np.random.default_rng(12)
    K = 5
    roads = [CONFIG1_WAYPOINTS]
    for _ in range(7):
        wx = np.cumsum(rng.uniform(15, 45, K)); wy = rng.uniform(-25, 25, K)
        roads.append((list(wx), list(wy)))
    wx = np.array([r[0] for r in roads]); wy = np.array([r[1] for r in roads])
    cx, cy, cyaw, npts, ck = ops.spline_courses(T(wx), T(wy), ds=0.1, curvature=True)
    npts = npts.cpu().numpy()
    for c, (rx, ry) in enumerate(roads):
        hx, hy, hyaw = spline_course(rx, ry, 0.1)
        assert npts[c] == len(hx), (c, npts[c], len(hx))                                  # sample count: exact
        n = len(hx)
        assert np.allclose(cx[c, :n].cpu().numpy(), hx, rtol=1e-12, atol=1e-11)
        assert np.allclose(cy[c, :n].cpu().numpy(), hy, rtol=1e-12, atol=1e-11)
        dyaw = np.angle(np.exp(1j * (cyaw[c, :n].cpu().numpy() - hyaw)))
        assert np.abs(dyaw).max() <= 1e-10
        assert torch.isnan(cx[c, n:]).all()
        # curvature: calc_curvature (cubic_spline_planner.py:156-165) from the host spline coefficients
        from sccav_cbf_b200.course import _natural_spline
        seg = np.hypot(np.diff(rx), np.diff(ry))
        sk = np.concatenate([[0.0], np.cumsum(seg)])
        _, bx_, cx_, dx_ = _natural_spline(list(sk), rx)
        _, by_, cy_, dy_ = _natural_spline(list(sk), ry)
        ts = np.arange(0, sk[-1], 0.1)
        ii = np.clip(np.searchsorted(sk, ts, side="right") - 1, 0, len(sk) - 2)
        u = ts - sk[ii]
        gx = np.array(bx_)[ii] + 2.0 * np.array(cx_)[ii] * u + 3.0 * np.array(dx_)[ii] * u ** 2
        gy = np.array(by_)[ii] + 2.0 * np.array(cy_)[ii] * u + 3.0 * np.array(dy_)[ii] * u ** 2
        hx2 = 2.0 * np.array(cx_)[ii] + 6.0 * np.array(dx_)[ii] * u
        hy2 = 2.0 * np.array(cy_)[ii] + 6.0 * np.array(dy_)[ii] * u
        kref = (hy2 * gx - hx2 * gy) / ((gx ** 2 + gy ** 2) ** (3 / 2))
        assert np.allclose(ck[c, :n].cpu().numpy(), kref, rtol=1e-9, atol=1e-12)
    assert npts[0] == 2034
    # closed loop: a group of vehicles per road, one launch each, course tensors straight from KC
    b = sc.config2(n_total=65536, M=8, T=250, lo=0, hi=3 * 128)
    for g in range(3):
        c = g + 1
        n = int(npts[c])
        course_d = (cx[c, :n].contiguous(), cy[c, :n].contiguous(), cyaw[c, :n].contiguous())
        course_h = tuple(t.cpu().numpy() for t in course_d)
        sl = slice(g * 128, (g + 1) * 128)
        st = np.ascontiguousarray(b.state[:, sl]); ob = np.ascontiguousarray(b.obst[:, :, sl])
        st[0] += course_h[0][0]; st[1] += course_h[1][0] - 5.0                               # start near this road's first point
        out = ops.rollout(ops.make_params(), b.slot_desc, T(st), T(ob), course_d, b.T)
        ref = co.rollout(co.default_params(), b.slot_desc, st, ob, course_h, b.T)
        same = (out["target_idx"].cpu().numpy() == ref["target_idx"]) & (out["n_active"].cpu().numpy() == ref["n_active"])
        assert same.mean() >= 0.99


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 1e-4)])
def test_device_lane_fit_vs_oracle_and_reference_vectors(dtype, tol, golden_dir):
    """KL (sccav_fit_lanes_*): the batched weighted polynomial lane fit against the oracle's least-squares
    solution -- the reference's own inputs (PolyLane.fit_polynomial_curve vectors, ragged point counts, mixed
    degrees launched per degree) and 4,096 random lanes; degenerate lanes report status 1.  (fp32 storage: the
    kernel computes in double -- powers of x about x = 0 cancel in fp32 -- so only the rounding of the inputs and
    of the six coefficients is left.)"""
    import os
    from sccav_cbf_b200 import ops
    g = np.load(os.path.join(golden_dir, "reference_vectors_lane_sadbm.npz"))
    dev = torch.device("cuda", 0)
    for n in range(1, 6):
        sel = np.nonzero(g["fit_n"] == n)[0]
        K = int(g["fit_k"][sel].max())
        x = np.nan_to_num(g["fit_x"][sel, :K].T.copy()); y = np.nan_to_num(g["fit_y"][sel, :K].T.copy())
        sg = np.nan_to_num(g["fit_sigma"][sel, :K].T.copy(), nan=1.0)
        cnt = torch.from_numpy(g["fit_k"][sel].astype(np.int32)).to(dev)
        c, st = ops.fit_lanes(torch.from_numpy(x).to(dev, dtype), torch.from_numpy(y).to(dev, dtype), n=n,
                              sigma=torch.from_numpy(sg).to(dev, dtype), count=cnt)
        assert int(st.sum()) == 0
        c = c.cpu().double().numpy()
        for j, i in enumerate(sel):
            k = int(g["fit_k"][i])
            ref = o.fit_polynomial(g["fit_x"][i, :k], g["fit_y"][i, :k], n, g["fit_sigma"][i, :k])
            xx = np.linspace(g["fit_x"][i, :k].min(), g["fit_x"][i, :k].max(), 64)
            got = np.polyval(c[: n + 1, j][::-1], xx); want = np.polyval(ref[::-1], xx)
            assert np.abs(got - want).max() <= tol * (1 + np.abs(want).max()), (n, i)
            assert (c[n + 1:, j] == 0).all()
    # random cubic lanes, default sigma
    rng = np.random.default_rng(12)
    C, K = 4096, 24
    x = np.sort(rng.uniform(-10, 70, (K, C)), axis=0)
    true = np.stack([rng.uniform(-4, 4, C), rng.uniform(-0.2, 0.2, C), rng.uniform(-3e-3, 3e-3, C), rng.uniform(-3e-5, 3e-5, C)])
    y = sum(true[j] * x ** j for j in range(4)) + rng.normal(0, 0.03, (K, C))
    x[:, 7] = 3.0                                                     # a degenerate lane: one abscissa only
    c, st = ops.fit_lanes(torch.from_numpy(x).to(dev, dtype), torch.from_numpy(y).to(dev, dtype), n=3)
    st = st.cpu().numpy(); c = c.cpu().double().numpy()
    assert st[7] == 1 and np.isnan(c[:, 7]).all() and st.sum() == 1
    for j in list(range(0, C, 97)):
        if j == 7:
            continue
        ref = o.fit_polynomial(x[:, j], y[:, j], 3)
        xx = np.linspace(x[:, j].min(), x[:, j].max(), 32)
        assert np.abs(np.polyval(c[:4, j][::-1], xx) - np.polyval(ref[::-1], xx)).max() <= tol * 10


@pytest.mark.gpu
def test_batched_lane_fit_feeds_the_lane_barrier():
    """PolyLane.fit_polynomial_curves -> a batched PolyLane in an obstacle list -> solve_cbf: the fitted coefficients
    go straight from kernel KL into the LANE slot fields, one lane per vehicle."""
    from sccav_cbf_b200 import DBM_CBF_2DS, PolyLane
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(3)
    C, K = 64, 16
    x = np.sort(rng.uniform(0, 60, (K, C)), axis=0)
    c0 = rng.uniform(2.5, 4.0, C)
    y = c0 + 0.01 * x + rng.normal(0, 0.01, (K, C))
    lane = PolyLane.fit_polynomial_curves(torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev), n=1, buffer=1.5)
    assert int(lane.fit_status.sum()) == 0 and lane.coeffs.shape == (2, C)
    assert np.abs(lane.coeffs[0].cpu().numpy() - c0).max() < 0.05
    ctl = DBM_CBF_2DS(alpha=1.0)
    ctl.set_model_params(lr=1.45, lf=1.45)
    ctl.obstacle_list2d["left"] = lane
    s = torch.stack([torch.full((C,), 20.0), torch.full((C,), 1.9), torch.full((C,), 0.25), torch.full((C,), 8.0)]).double().to(dev)
    ctl.update_state(s)
    info, u = ctl.solve_cbf(torch.stack([torch.zeros(C), torch.full((C,), 0.1)]).double().to(dev), return_solver=True)
    ref = []
    for n_ in range(C):
        f = [[1.5, float(lane.coeffs[0, n_]), float(lane.coeffs[1, n_]), 0, 0, 0, 0, 0]]
        ref.append(o.filter_step(o.MODEL_DBM, [20.0, 1.9, 0.25, 8.0], [0.0, 0.1], [o.SLOT_LANE], f, 1.0, 1.45, 1.45, 2.9, (1, 0, 0, 1)))
    assert np.abs(u[1].cpu().numpy() - np.array([r[1] for r in ref])).max() < 1e-7      # (lane parity proper: test_gpu_parity)
    assert (info["status"].cpu().numpy() == np.array([r[3] for r in ref])).all() and int((info["status"] == 1).sum()) > C // 2


@pytest.mark.gpu
def test_lane_fit_degenerate_inputs():
    """KL: too few points for the degree, and argument errors through the C-ABI."""
    from sccav_cbf_b200 import ops
    dev = torch.device("cuda", 0)
    x = torch.tensor([[0.0, 0.0], [1.0, 1.0], [2.0, 2.0]], dtype=torch.float64, device=dev)       # K = 3, C = 2
    y = torch.tensor([[1.0, 1.0], [2.0, 3.0], [3.0, 5.0]], dtype=torch.float64, device=dev)
    c, st = ops.fit_lanes(x, y, n=1)
    assert st.tolist() == [0, 0]
    assert torch.allclose(c[:2].cpu(), torch.tensor([[1.0, 1.0], [1.0, 2.0]], dtype=torch.float64), atol=1e-12) and (c[2:] == 0).all()
    c, st = ops.fit_lanes(x, y, n=3)                                   # 3 points cannot fix 4 coefficients
    assert st.tolist() == [1, 1] and torch.isnan(c).all()
    cnt = torch.tensor([3, 1], dtype=torch.int32, device=dev)
    c, st = ops.fit_lanes(x, y, n=1, count=cnt)
    assert st.tolist() == [0, 1] and torch.isnan(c[:, 1]).all()
    with pytest.raises(ValueError):                                     # SCCAV_EINVAL -> ValueError, like the other entry points
        ops.fit_lanes(x, y, n=6)
