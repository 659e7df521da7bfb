"""Shared helpers for the parity tests: seeded problem generators in the C-ABI layouts."""
import numpy as np

from oracle import oracle as o

NF = 8


def random_states(rng, N, v_lo=2.0, v_hi=12.0):
    s = np.empty((4, N))
    s[0] = rng.uniform(-20, 120, N)
    s[1] = rng.uniform(-45, 15, N)
    s[2] = rng.uniform(-3.5, 3.5, N)
    s[3] = rng.uniform(v_lo, v_hi, N)
    return s


def random_slots(rng, N, slot_types, s, near=True):
    """obst [M, 8, N] for the given slot types, placed around the vehicles so that a good share
    of the rows is active."""
    M = len(slot_types)
    ob = np.zeros((M, NF, N))
    for m, t in enumerate(slot_types):
        t &= 0x3F
        ahead = rng.uniform(2, 30, N) if near else rng.uniform(20, 60, N)
        lat = rng.uniform(-6, 6, N)
        cx = s[0] + ahead * np.cos(s[2]) - lat * np.sin(s[2])
        cy = s[1] + ahead * np.sin(s[2]) + lat * np.cos(s[2])
        if t == o.SLOT_ELLIPSE:
            ob[m, 0], ob[m, 1] = cx, cy
            ob[m, 2] = rng.uniform(2, 6, N) + 0.5
            ob[m, 3] = rng.uniform(1, 3, N) + 0.5
            ob[m, 4] = rng.uniform(-np.pi, np.pi, N)
            ob[m, 5] = rng.uniform(-2, 2, N)
            ob[m, 6] = rng.uniform(-2, 2, N)
        elif t == o.SLOT_CONE:
            ob[m, 0], ob[m, 1] = cx, cy
            ob[m, 2] = rng.uniform(-3.2, 3.2, N)
            ob[m, 3] = rng.uniform(0, 8, N)
            ob[m, 4] = rng.uniform(1, 4, N) + 1.5
            ob[m, 5] = np.where(rng.uniform(size=N) < 0.7, 0.0, rng.uniform(-0.2, 0.2, N))
        elif t in (o.SLOT_LANE, o.SLOT_LANE_SQRT):
            ob[m, 0] = 1.5
            c1 = rng.uniform(-0.3, 0.3, N)
            ob[m, 2] = c1
            ob[m, 3] = rng.uniform(-0.004, 0.004, N) * (m % 2)
            ob[m, 4] = rng.uniform(-4e-5, 4e-5, N) * (m % 2)
            x = s[0]
            g_wo = c1 * x + ob[m, 3] * x * x + ob[m, 4] * x ** 3
            ob[m, 1] = s[1] - g_wo + rng.uniform(-5, 5, N)       # c0: lane passes within 5 m of the vehicle
        elif t == o.SLOT_RADIAL:
            ob[m, 0], ob[m, 1] = cx, cy
            r = rng.uniform(1.5, 2.0, N)
            ob[m, 2], ob[m, 3] = r, r
            ob[m, 4] = 1.0
            ob[m, 5] = rng.uniform(-4, 4, N)
            ob[m, 6] = rng.uniform(-4, 4, N)
        elif t == o.SLOT_DISTANCE:
            ob[m, 0], ob[m, 1] = cx, cy
            ob[m, 2] = rng.uniform(2, 8, N)
    return ob


def random_uref(rng, N, kbm=False):
    u = np.empty((2, N))
    u[0] = rng.uniform(4, 10, N) if kbm else rng.uniform(-2, 2, N)
    u[1] = rng.uniform(-0.45, 0.45, N)
    return u


def kkt_residuals(A, b, u, r, R, mask):
    """Independent KKT check of a claimed optimum of min (u-r)^T R (u-r) s.t. A u >= b.
    A [2,M,N], b [M,N], u/r [2,N], R (4,) row-major, mask [N] uint.  Returns
    (primal violation, stationarity residual, min multiplier) per problem, via least squares on
    the claimed active rows."""
    M, N = b.shape
    R = np.asarray(R, dtype=np.float64).reshape(2, 2)
    prim = np.zeros(N); stat = np.zeros(N); lmin = np.zeros(N)
    for n in range(N):
        res = A[0, :, n] * u[0, n] + A[1, :, n] * u[1, n] - b[:, n]
        scale = np.abs(A[0, :, n] * u[0, n]) + np.abs(A[1, :, n] * u[1, n]) + np.abs(b[:, n]) + 1e-300
        prim[n] = np.max(-res / scale)
        g = 2.0 * R @ (u[:, n] - r[:, n])
        act = [k for k in range(M) if (int(mask[n]) >> k) & 1]
        if not act:
            stat[n] = np.linalg.norm(g)
            lmin[n] = 0.0
            continue
        AW = np.stack([A[:, k, n] for k in act], axis=1)          # 2 x |W|
        lam, *_ = np.linalg.lstsq(AW, g, rcond=None)
        stat[n] = np.linalg.norm(AW @ lam - g) / (np.linalg.norm(g) + 1e-300)
        lmin[n] = lam.min()
        # active rows must be tight
        prim[n] = max(prim[n], np.max(np.abs(res[act]) / scale[act]))
    return prim, stat, lmin
