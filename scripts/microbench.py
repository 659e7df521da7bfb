#!/usr/bin/env python
"""Developer tool: kernel-only timings of the rollout (config 2) and of the fused operator (2 M x 8),
CUDA events, median of a few launches.  Not a bench line -- bench.py is the contract."""
import argparse
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from sccav_cbf_b200 import ops, scenarios as sc  # noqa: E402
from sccav_cbf_b200.rollout import ClosedLoopRollout  # noqa: E402


def timed(fn, reps):
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return statistics.median(ms)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--vehicles", type=int, default=65536)
    ap.add_argument("--T", type=int, default=1000)
    ap.add_argument("--M", type=int, default=8)
    ap.add_argument("--dtype", default="f64")
    ap.add_argument("--no-rollout", action="store_true")
    ap.add_argument("--no-operator", action="store_true")
    ap.add_argument("--ingest", action="store_true", help="time the bounding-box ingest kernel (KB)")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    dtype = torch.float64 if a.dtype == "f64" else torch.float32
    out = []
    M = a.M
    if not a.no_rollout:
        batch = sc.config2(n_total=a.vehicles, M=M, T=a.T, seed=0, lo=0, hi=a.vehicles)
        cl = ClosedLoopRollout(batch, dtype=dtype, device=dev)
        for _ in range(2):
            res = cl.run()
        ms = timed(cl.run, 5)
        solves = float(res["steps"].sum().item()) * M
        out.append("rollout %.3f ms  %.4g solves/s  (nact %d, ninf %d, evals/step %.1f, %s)" % (
            ms, solves / ms * 1e3, int(res["n_active"].sum().item()), int(res["n_infeasible"].sum().item()),
            float(res["n_evals"].double().sum().item()) / float(res["steps"].sum().item()),
            ops.rollout_launch_info(batch.slot_desc, batch.N, 2034, dtype)))
        b2 = sc.config2(n_total=a.vehicles, M=M, T=a.T, seed=0, lo=0, hi=a.vehicles)
        b2.params = dict(b2.params, flags=1)
        cl2 = ClosedLoopRollout(b2, dtype=dtype, device=dev)
        for _ in range(2):
            res2 = cl2.run()
        ms2 = timed(cl2.run, 5)
        same = (res2["steps"] == res["steps"]) & (res2["target_idx"] == res["target_idx"]) & (res2["n_active"] == res["n_active"])
        out.append("rollout[prepared rows] %.3f ms  %.4g solves/s  (identical bookkeeping vs canonical %.5f)" % (
            ms2, solves / ms2 * 1e3, float(same.double().mean().item())))
        for fl in (4, 5):
            b3 = sc.config2(n_total=a.vehicles, M=M, T=a.T, seed=0, lo=0, hi=a.vehicles)
            b3.params = dict(b3.params, flags=fl)
            cl3 = ClosedLoopRollout(b3, dtype=dtype, device=dev)
            for _ in range(2):
                res3 = cl3.run()
            ms3 = timed(cl3.run, 5)
            same3 = (res3["steps"] == res["steps"]) & (res3["target_idx"] == res["target_idx"]) & (res3["n_active"] == res["n_active"])
            out.append("rollout[flags=%d: fused steer%s] %.3f ms  %.4g solves/s  (identical bookkeeping vs canonical %.5f)" % (
                fl, " + prepared rows" if fl & 1 else "", ms3, solves / ms3 * 1e3, float(same3.double().mean().item())))
    if not a.no_operator:
        n_op = 2 * 1024 * 1024
        gen = torch.Generator(device=dev); gen.manual_seed(1234)
        st = torch.empty((4, n_op), dtype=dtype, device=dev)
        st[0].uniform_(-20, 120, generator=gen); st[1].uniform_(-45, 15, generator=gen)
        st[2].uniform_(-3.2, 3.2, generator=gen); st[3].uniform_(2, 12, generator=gen)
        ob = torch.zeros((M, 8, n_op), dtype=dtype, device=dev)
        ob[:, 0].uniform_(-20, 120, generator=gen); ob[:, 1].uniform_(-45, 15, generator=gen)
        ob[:, 2].uniform_(2.5, 6.5, generator=gen); ob[:, 3].uniform_(1.5, 3.5, generator=gen)
        ob[:, 4].uniform_(-3.1, 3.1, generator=gen)
        ur = torch.zeros((2, n_op), dtype=dtype, device=dev); ur[1].uniform_(-0.3, 0.3, generator=gen)
        prm = ops.make_params()
        sd = [0] * M
        for _ in range(3):
            u, mask, status, hmin = ops.filter_step(prm, sd, st, ob, ur)
        ms = timed(lambda: ops.filter_step(prm, sd, st, ob, ur), 7)
        es = 8 if a.dtype == "f64" else 4
        bps = 7 * es + (4 * es + 2 * es + 2 * es + 4 + 1) / M
        out.append("operator %.4f ms  %.1f GB/s algorithmic  %.4g solves/s  (active %.3f%%, infeasible %.4f%%)" % (
            ms, bps * n_op * M / ms / 1e6, n_op * M / ms * 1e3,
            100.0 * float((status == 1).sum().item()) / n_op, 100.0 * float((status == 2).sum().item()) / n_op))
        for static in (0, 0x40):
            sdp, obp = ops.prepare_obstacles([static] * M, ob)
            for _ in range(3):
                u2, mask2, status2, _ = ops.filter_step(prm, sdp, st, obp, ur)
            ms = timed(lambda: ops.filter_step(prm, sdp, st, obp, ur), 7)
            nf = 6 if static else 8
            bps = nf * es + (4 * es + 2 * es + 2 * es + 4 + 1) / M
            out.append("operator[prepared%s] %.4f ms  %.1f GB/s algorithmic (%.1f B/solve)  %.4g solves/s  masks equal %.6f  max|du| %.2e" % (
                ", static" if static else "", ms, bps * n_op * M / ms / 1e6, bps, n_op * M / ms * 1e3,
                float((mask2 == mask).double().mean().item()), float((u2 - u).abs().max().item())))
        msp = timed(lambda: ops.prepare_obstacles([0] * M, ob, out=obp), 5)
        out.append("prepare %.4f ms  %.1f GB/s (128 B per slot)" % (msp, 128.0 * n_op * M / msp / 1e6))
    if a.ingest:
        n_v, K = 2 * 1024 * 1024, 8
        gen = torch.Generator(device=dev); gen.manual_seed(7)
        ids = torch.full((M, n_v), -1, dtype=torch.int32, device=dev)
        ob = torch.zeros((M, 8, n_v), dtype=dtype, device=dev)
        cnt = torch.zeros(n_v, dtype=torch.int32, device=dev)
        box = torch.rand((K, 6, n_v), dtype=dtype, device=dev, generator=gen) * 10 + 1
        es = 8 if a.dtype == "f64" else 4

        def tick(seed):
            gen.manual_seed(seed)
            bid = torch.randint(0, 24, (K, n_v), dtype=torch.int32, device=dev, generator=gen)
            bid = torch.where(torch.rand((K, n_v), device=dev, generator=gen) < 0.7, bid, torch.full_like(bid, -1))
            return bid
        bids = [tick(s) for s in range(6)]
        for b in bids[:2]:
            ops.ingest_boxes(0, b, box, ids, ob, cnt)
        ms = []
        for b in bids[2:]:
            ms.append(timed(lambda: ops.ingest_boxes(0, b, box, ids, ob, cnt), 1))
        ms = statistics.median(ms)
        byt = n_v * (K * (4 + 6 * es) + M * (4 + 8 * es) * 2 + 8)
        out.append("ingest %.4f ms  %.1f GB/s (upper-bound bytes: K boxes read, M slots read + written)  %.3g vehicle-lists/s  mean count %.2f" % (
            ms, byt / ms / 1e6, n_v / ms * 1e3, float(cnt.double().mean().item())))
    print("\n".join(out), flush=True)


if __name__ == "__main__":
    main()
