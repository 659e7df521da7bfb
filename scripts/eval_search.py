#!/usr/bin/env python
"""CPU evaluation of the nearest-way-point search on real closed-loop queries: the C oracle rolls out a prefix of
config 2 and records every step; the front-axle points of every step are then fed to the search through the
test hook (same __host__ __device__ code as the kernels) with the hint the rollout kernel would use
(previous index + previous advance).  Reports evaluations per query and the SIMT proxy: the mean over
(warp of 32 consecutive vehicles, step) of the maximum over its lanes."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import c_oracle as co  # noqa: E402
from sccav_cbf_b200 import _native as nv, scenarios as sc  # noqa: E402


def main(n=1024, T=1000):
    b = sc.config2(n_total=65536, M=8, T=T, seed=0, lo=0, hi=n)
    prm = co.default_params(**b.params)
    res = co.rollout(prm, b.slot_desc, b.state, b.obst, b.course, T, record_stride=1)
    tr = res["traj"]                                   # [T][7][N]
    x, y, yaw = tr[:, 0], tr[:, 1], tr[:, 2]
    fx = x + 2.9 * np.cos(yaw); fy = y + 2.9 * np.sin(yaw)
    cx, cy, _ = b.course
    L = nv.lib()
    near = np.zeros(n, np.int32); adv = np.zeros(n, np.int32)
    ev_all = np.zeros((T, n), np.int64)
    for t in range(T):
        hint = (near + adv).astype(np.int32)
        idx = np.empty(n, np.int32); full = np.empty(n, np.int32); ev = np.empty(n, np.int64)
        fxt = np.ascontiguousarray(fx[t]); fyt = np.ascontiguousarray(fy[t])
        rc = L.sccav_debug_course_index_host(cx.ctypes.data, cy.ctypes.data, len(cx), fxt.ctypes.data, fyt.ctypes.data,
                                             hint.ctypes.data, n, 64, idx.ctypes.data, full.ctypes.data, ev.ctypes.data)
        assert rc == 0
        assert np.array_equal(idx, full), t
        a = idx - near
        a[np.abs(a) > 32] = 0
        adv = a.astype(np.int32); near = idx
        ev_all[t] = ev
    ev = ev_all[1:]
    w = ev.reshape(ev.shape[0], n // 32, 32)
    print("queries %d  evals/query mean %.1f  median %.0f  p99 %.0f  max %d   warp-max mean %.1f" %
          (ev.size, ev.mean(), np.median(ev), np.percentile(ev, 99), ev.max(), w.max(axis=2).mean()))


if __name__ == "__main__":
    main(*(int(a) for a in sys.argv[1:]))
