"""KD, the persistent driver-tick kernel (sccav_drive_ticks_*), against the scalar restatement of the CARLA driver's loop
(oracle.drive_ticks; carla_scripts/multi_obstacle_CBF_local_with_lanes.py:861-983): LateralStanley class form -> PID1 with
ki / kd and a different dt every tick -> two lane barriers + one fresh collision cone per actor box -> DBM_CBF_2DS ->
throttle / brake / steer, 600 ticks in one launch."""
import numpy as np
import pytest
import torch

from oracle import oracle as o

gpu = pytest.mark.gpu


def scenario(N, T, K, seed):
    rng = np.random.default_rng(seed)
    # the lane of the driver's map: boundaries y = 19.4 and y = 10.4 (:275-297), reference line between them
    tx = np.arange(-95.0, 40.0, 0.1)
    ty = 14.9 + 0.6 * np.sin((tx + 95.0) / 18.0)
    tyaw = np.arctan2(np.gradient(ty), np.gradient(tx))
    tv = 7.0 + 1.5 * np.sin((tx + 95.0) / 25.0)
    s0 = np.stack([rng.uniform(-94.0, -90.0, N), 14.9 + rng.uniform(-1.0, 1.0, N), rng.uniform(-0.15, 0.15, N), rng.uniform(3.0, 9.0, N)])
    dts = 1.0 / 30.0 + rng.uniform(-0.004, 0.004, T)
    dts[0] = 0.05
    tt = np.concatenate([[0.0], np.cumsum(dts)[:-1]])
    box_id = np.full((T, K, N), -1, np.int32)
    box = np.zeros((T, K, 6, N))
    for n in range(N):
        for k in range(K):
            if rng.uniform() < 0.25:
                continue                                       # this actor never shows up for this ego
            t0, t1 = sorted(rng.integers(0, T, 2))
            if t1 - t0 < 30:
                t1 = min(T, t0 + 200)
            vk = rng.uniform(0.0, 6.0)
            x0 = s0[0, n] + rng.uniform(12.0, 60.0)
            y0 = 14.9 + rng.uniform(-2.5, 2.5)
            yaw = rng.uniform(-0.2, 0.2)
            sl = slice(t0, t1)
            box_id[sl, k, n] = 100 + k
            box[sl, k, 0, n] = rng.uniform(1.8, 2.6); box[sl, k, 1, n] = rng.uniform(0.8, 1.1)
            box[sl, k, 2, n] = x0 + vk * np.cos(yaw) * tt[sl]; box[sl, k, 3, n] = y0 + vk * np.sin(yaw) * tt[sl]
            box[sl, k, 4, n] = yaw; box[sl, k, 5, n] = vk
    lanes = [[1.5, 19.4, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0], [1.5, 10.4, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0]]
    return (tx, ty, tyaw, tv), s0, dts, box_id, box, lanes


PRM = dict(k=1.2, ks=10.0, L=1.45, lr=1.45, lf=1.45, alpha=1.0)         # LateralStanley(lr, lf, k=1.2, ks=10) :657; DBM_CBF_2DS() :662
DRV = dict(kp=1.0, kd=0.01, ki=0.01, rad_to_steer=1.0 / 1.2217, max_steer_cmd=1.0, rate=0.1, cone_buffer=1.5)


def run_gpu(traj, s0, dts, box_id, box, lanes, ego=None, flags=0, real=torch.float64):
    from sccav_cbf_b200 import ops
    dev = torch.device("cuda", 0)
    T, K, N = box_id.shape
    M = 2 + K
    t = lambda a, dt=real: torch.from_numpy(np.ascontiguousarray(a)).to(dt).to(dev)
    obst = torch.zeros((M, 8, N), dtype=real, device=dev)
    for m in range(2):
        obst[m, :, :] = t(np.array(lanes[m]))[:, None]
    sd = [o.SLOT_LANE | 0x80, o.SLOT_LANE | 0x80] + [o.SLOT_CONE] * K
    prm = ops.make_params(k_stanley=PRM["k"], ks_stanley=PRM["ks"], L=PRM["L"], lr=PRM["lr"], lf=PRM["lf"], alpha=PRM["alpha"])
    tidx = torch.zeros((N,), dtype=torch.int32, device=dev)
    carry = torch.zeros((4, N), dtype=real, device=dev)
    out = ops.drive_ticks(prm, sd, 2, obst, tuple(t(c) for c in traj), tidx, carry, T, state0=None if ego is not None else t(s0),
                          ego=None if ego is None else t(ego), box_id=t(box_id, torch.int32), box=t(box), dt=t(dts), act_flags=flags, **DRV)
    torch.cuda.synchronize()
    res = {k: (v.cpu().numpy() if v is not None else None) for k, v in out.items()}
    res["carry"] = carry.cpu().numpy(); res["last_idx"] = tidx.cpu().numpy()
    return res


def run_oracle(traj, s0, dts, box_id, box, lanes, n, ego=None, reset_brake=False):
    T = len(dts)
    return o.drive_ticks(None if ego is not None else s0[:, n], None if ego is None else [ego[t, :, n] for t in range(T)],
                         [box_id[t, :, n] for t in range(T)], [box[t, :, :, n] for t in range(T)], dts,
                         [o.SLOT_LANE, o.SLOT_LANE], lanes, 2 + box_id.shape[1], traj, params=PRM, drive=dict(DRV, reset_brake=reset_brake))


@gpu
@pytest.mark.parametrize("flags", [0, 1])
def test_driver_ticks_with_a_simulator_stream(flags):
    """The simulator owns the plant: ego states come as a stream.  No feedback through our arithmetic, so every tick must
    match the scalar restatement strictly: commands 1e-9, active sets and target indices identical, carried PID state."""
    N, T, K = 24, 600, 5
    traj, s0, dts, box_id, box, lanes = scenario(N, T, K, 5)
    rng = np.random.default_rng(9)
    tx, ty, tyaw, tv = traj
    # a plausible recorded drive: along the reference line with lateral wander, heading and speed noise
    k = (np.linspace(20, 1150, T)[:, None] + rng.uniform(0, 40, N)[None, :]).astype(int)
    ego = np.stack([tx[k] + rng.normal(0, 0.05, (T, N)), ty[k] + 0.8 * np.sin(np.arange(T)[:, None] / 40.0 + rng.uniform(0, 6, N)[None, :]),
                    tyaw[k] + rng.normal(0, 0.03, (T, N)), 6.0 + 2.0 * np.sin(np.arange(T)[:, None] / 90.0) + rng.normal(0, 0.1, (T, N))], axis=1)
    g = run_gpu(traj, s0, dts, box_id, box, lanes, ego=ego, flags=flags)
    nact = 0
    for n in range(N):
        r = run_oracle(traj, s0, dts, box_id, box, lanes, n, ego=ego, reset_brake=bool(flags))
        assert np.abs(g["act"][:, :, n] - np.array(r["act"])).max() <= 1e-9
        ru = np.array(r["u"])
        assert (np.abs(g["u"][:, :, n] - ru) <= 1e-9 * (1 + np.abs(ru))).all()
        assert np.array_equal(g["active_mask"][:, n].view(np.uint32), np.array(r["mask"], dtype=np.uint32))
        assert np.array_equal(g["target_idx"][:, n], np.array(r["idx"]))
        assert np.abs(g["carry"][:, n] - np.array(r["carry"])).max() <= 1e-9 and g["last_idx"][n] == r["target_idx"]
        nact += int((np.array(r["mask"]) != 0).sum())
    assert nact > 200, nact


@gpu
def test_driver_ticks_closed_loop_with_the_stand_in_plant():
    """No simulator: State.update_com closes the loop for 600 ticks of varying dt.  Free-running, so the comparison is the
    usual one: identical active sets / indices on (almost) every tick, commands and final states to 1e-6."""
    N, T, K = 16, 600, 5
    traj, s0, dts, box_id, box, lanes = scenario(N, T, K, 11)
    g = run_gpu(traj, s0, dts, box_id, box, lanes)
    same, tot, nact, whole = 0, 0, 0, 0
    for n in range(N):
        r = run_oracle(traj, s0, dts, box_id, box, lanes, n)
        m = np.array(r["mask"], dtype=np.uint32)
        eq = (g["active_mask"][:, n].view(np.uint32) == m) & (g["target_idx"][:, n] == np.array(r["idx"]))
        same += int(eq.sum()); tot += T; nact += int((m != 0).sum())
        if eq.all():
            whole += 1
            assert np.abs(g["act"][:, :, n] - np.array(r["act"])).max() <= 1e-6
            assert np.abs(g["state"][:, n] - np.array(r["state"])).max() <= 1e-6
        else:
            # an ego boxed in by random actors meets contradictory rows; once a least-violation answer differs in its last
            # bits the two runs are different drives.  Their common beginning must still agree.
            first = max(0, int(np.argmin(eq)) - 100)
            assert first == 0 or np.abs(g["act"][:first, :, n] - np.array(r["act"])[:first]).max() <= 1e-4
    assert whole >= N - 3 and same >= 0.95 * tot, (whole, same, tot)
    assert nact > 100, nact
    # the lane barriers were honoured (an ego whose rows contradict each other may leave)
    assert ((g["state"][1] > 10.4) & (g["state"][1] < 19.4)).sum() >= N - 3


def test_oracle_driver_tick_matches_the_reference_pid_and_stanley_classes():
    """The pieces oracle.drive_ticks chains are the reference's classes: PID1 with ki / kd / variable dt
    (cbf/controllers.py:153-180) and LateralStanley.control (:104-151), restated literally here from those lines."""
    rng = np.random.default_rng(2)
    pid = o.PID1(kp=1.0, kd=0.01, ki=0.01)
    e_prev, ie = 0.0, 0.0
    for _ in range(200):
        x, xref, dt = rng.uniform(0, 10), rng.uniform(0, 10), rng.uniform(0.01, 0.06)
        pid.dt = dt
        u = pid.control(x, xref)
        e = xref - x                                   # controllers.py:174-179
        de = (e - e_prev) / dt
        ie += dt * e
        assert u == 1.0 * e + 0.01 * ie + 0.01 * de
        e_prev = e
    traj, s0, dts, box_id, box, lanes = scenario(2, 4, 1, 1)
    tx, ty, tyaw, tv = traj
    last = 0
    for _ in range(100):
        x, y, yaw, v = rng.uniform(-90, 30), 14.9 + rng.uniform(-2, 2), rng.uniform(-0.3, 0.3), rng.uniform(0, 9)
        d, idx = o.stanley_control(x, y, yaw, v, tx, ty, tyaw, last, 1.2, 1.45, 10.0)
        fx, fy = x + 1.45 * np.cos(yaw), y + 1.45 * np.sin(yaw)                                  # :78-81
        dd = np.hypot(fx - tx, fy - ty)
        ti = int(np.argmin(dd))                                                                  # :92-93
        efa = np.dot([fx - tx[ti], fy - ty[ti]], [-np.cos(yaw + np.pi / 2), -np.sin(yaw + np.pi / 2)])   # :96-99
        if last >= ti:
            ti = last                                                                            # :118-119
        ref = o.normalize_angle(tyaw[ti] - yaw) + np.arctan2(1.2 * efa, v + 10.0)                # :140-146
        assert idx == ti and abs(d - ref) <= 1e-15
        last = idx if rng.uniform() < 0.8 else 0


@gpu
def test_driver_ticks_fp32_variant_follows_the_fp64_one():
    """The reported fp32 variant of KD on a simulator stream (no feedback through our arithmetic): commands within 1e-3 of
    the fp64 run, the same target indices on (almost) every tick."""
    N, T, K = 24, 300, 5
    traj, s0, dts, box_id, box, lanes = scenario(N, T, K, 5)
    rng = np.random.default_rng(9)
    tx, ty, tyaw, tv = traj
    k = (np.linspace(20, 1150, T)[:, None] + rng.uniform(0, 40, N)[None, :]).astype(int)
    ego = np.stack([tx[k], ty[k] + 0.8 * np.sin(np.arange(T)[:, None] / 40.0 + rng.uniform(0, 6, N)[None, :]), tyaw[k],
                    6.0 + 2.0 * np.sin(np.arange(T)[:, None] / 90.0) + 0 * tx[k]], axis=1)
    g64 = run_gpu(traj, s0, dts, box_id, box, lanes, ego=ego)
    g32 = run_gpu(traj, s0, dts, box_id, box, lanes, ego=ego, real=torch.float32)
    assert (g64["target_idx"] == g32["target_idx"]).mean() > 0.98
    same = g64["active_mask"] == g32["active_mask"]
    assert same.mean() > 0.97
    d = np.abs(g64["act"] - g32["act"].astype(np.float64))
    assert np.median(d) < 1e-5 and (d[:, :, same.all(axis=0)] < 2e-2).mean() > 0.99
