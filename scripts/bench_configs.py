#!/usr/bin/env python
"""BASELINE.json configs 2-5 at their FULL per-GPU sizes on one B200: throughput of
the rollout kernel (CUDA events, inputs resident) plus the size-independent checks the full sizes
allow -- bookkeeping invariants, and a sample of vehicles taken from INSIDE the full-size batch
re-run by the CPU oracle (a vehicle's result must not depend on the batch it sits in).

    python scripts/bench_configs.py [--configs 2,3,4,5] [--out profiles/rNN_configs.jsonl]

One JSON line per (config, dtype).  bench.py remains the contract bench (config 2); this script is
the evidence for the other rows of 8d.  The oracle is used here as the checker only.
"""
import argparse
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import c_oracle as co  # noqa: E402
from sccav_cbf_b200 import ops, scenarios as sc  # noqa: E402
from sccav_cbf_b200.rollout import ClosedLoopRollout  # noqa: E402


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(1.0, np.abs(b))


def sample_batch(batch, idx):
    """The scenarios `idx` of a batch as a batch of their own (what the oracle re-runs)."""
    sub = sc.ScenarioBatch(batch.name + "_sample", np.ascontiguousarray(batch.state[:, idx]), list(batch.slot_desc),
                           None if batch.obst is None else np.ascontiguousarray(batch.obst[:, :, idx]), batch.course,
                           dict(batch.params), T=batch.T,
                           alpha=None if batch.alpha is None else np.ascontiguousarray(batch.alpha[idx]),
                           R=None if batch.R is None else np.ascontiguousarray(batch.R[:, idx]),
                           target_speed=None if batch.target_speed is None else np.ascontiguousarray(batch.target_speed[idx]))
    return sub


def oracle_run(b):
    return co.rollout(co.default_params(**b.params), b.slot_desc, b.state, b.obst, b.course, b.T,
                      alpha=b.alpha, R=b.R, target_speed=b.target_speed)


FLAGS = 0


def run_config(name, batch, dtype, reps, n_sample, dev):
    t0 = time.time()
    if FLAGS:
        batch.params = dict(batch.params, flags=FLAGS)
    cl = ClosedLoopRollout(batch, dtype=dtype, device=dev, pin=False)
    for _ in range(2):
        res = cl.run()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cl.reset()
        e0.record()
        res = ops.rollout(cl.params, cl.slot_desc, cl.d_state, cl.d_obst, cl.course, cl.T, alpha=cl.d_alpha, R=cl.d_R,
                          target_speed=cl.d_tspeed, record_stride=0, out=cl.out)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms_med = statistics.median(ms)
    g = {k: v.cpu().numpy() for k, v in res.items() if k != "obst"}
    steps = g["steps"].astype(np.int64)
    solves = float(steps.sum()) * batch.M
    line = {
        "config": name, "dtype": "f64" if dtype == torch.float64 else "f32", "vehicles": batch.N, "rows_per_vehicle": batch.M,
        "slots": sorted(set(int(d) for d in batch.slot_desc)), "T": batch.T, "ms_per_rollout": ms_med,
        "solves_per_s": solves / (ms_med * 1e-3), "vehicle_steps_per_s": float(steps.sum()) / (ms_med * 1e-3),
        "active_step_frac": float(g["n_active"].sum()) / max(1.0, float(steps.sum())),
        "infeasible_step_frac": float(g["n_infeasible"].sum()) / max(1.0, float(steps.sum())),
        "launch": ops.rollout_launch_info(batch.slot_desc, batch.N, 0 if batch.course is None else len(batch.course[0]), dtype),
    }
    # ---- invariants that hold at any size
    inv = {}
    if not batch.params.get("terminate"):
        inv["steps_eq_T"] = bool((steps == batch.T).all())
    else:
        inv["steps_le_T"] = bool((steps <= batch.T).all() and (steps >= 1).all())
        inv["distinct_step_counts"] = int(len(np.unique(steps)))
    # (an infeasible step whose least-violation candidate is u_ref itself has an empty active set)
    inv["counters_consistent"] = bool((g["n_infeasible"] <= steps).all() and (g["n_active"] <= steps).all()
                                      and (g["n_active"] >= 0).all() and (g["n_infeasible"] >= 0).all())
    inv["finite_state_frac"] = float(np.isfinite(g["state"]).all(axis=0).mean())
    if batch.course is not None:
        inv["target_idx_in_range"] = bool((g["target_idx"] >= 0).all() and (g["target_idx"] < len(batch.course[0])).all())
    # never-infeasible vehicles keep (discrete-time) safety: h_min stays above a small negative margin
    # (distance-like barriers only: the collision-cone h is a velocity-scaled quantity)
    ok = g["n_infeasible"] == 0
    if ok.any() and all((int(d) & 0x7f) in (0, 3) for d in batch.slot_desc):
        inv["h_min_ge_-0.05_frac_of_feasible"] = float((g["h_min"][ok] >= -0.05).mean())
    line["invariants"] = inv
    # ---- a sample from inside the full-size batch, re-run alone by the CPU oracle
    if n_sample > 0:
        rng = np.random.default_rng(12345)
        idx = np.sort(rng.choice(batch.N, size=min(n_sample, batch.N), replace=False))
        r = oracle_run(sample_batch(batch, idx))
        same = np.ones(len(idx), dtype=bool)
        for k in ("steps", "target_idx", "n_active", "n_infeasible"):
            same &= g[k][idx] == r[k]
        err = relerr(g["state"][:, idx], r["state"]).max(axis=0)
        tol = 1e-6 if dtype == torch.float64 else 5e-2
        line["oracle_sample"] = {
            "n": int(len(idx)), "identical_bookkeeping_frac": float(same.mean()),
            "state_relerr_median": float(np.median(err)), "state_relerr_le_tol_frac": float((err <= tol).mean()), "tol": tol,
        }
    line["wall_s"] = time.time() - t0
    del cl
    torch.cuda.empty_cache()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="2,3,4,5")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--sample", type=int, default=512)
    ap.add_argument("--scale", type=float, default=1.0, help="shrink every N (smoke runs)")
    ap.add_argument("--out", default="")
    ap.add_argument("--flags", type=int, default=0, help="sccav_params.flags for every config (5 = prepared rows + fused steer)")
    a = ap.parse_args()
    global FLAGS
    FLAGS = a.flags
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    want = set(a.configs.split(","))
    lines = []

    def n(x):
        return max(1024, int(x * a.scale))

    def emit(line):
        line["flags"] = a.flags
        lines.append(line)
        print(json.dumps(line), flush=True)

    if "2" in want:
        b = sc.config2(n_total=n(65536), M=8, T=1000)
        emit(run_config("config2: 65,536 vehicles x 8 static ellipses x 1000 steps", b, torch.float64, a.reps, a.sample, dev))
    if "3" in want:
        b = sc.config3(n_total=n(262144), M=16, T=600)
        emit(run_config("config3: 262,144 vehicles x 16 seeker circles (radial-dynamic TV-CBF) x 600 steps", b, torch.float64, a.reps, a.sample, dev))
        b.T = 1000
        emit(run_config("config3: same, 1000 steps", b, torch.float64, a.reps, 0, dev))
    if "4" in want:
        b = sc.config4(n_total=n(1048576), M=8, T=1000)
        l64 = run_config("config4: 1,048,576 vehicles x (8 ellipses + 2 lanes) x 1000 steps", b, torch.float64, a.reps, a.sample, dev)
        emit(l64)
        l32 = run_config("config4: same, fp32 variant", b, torch.float32, a.reps, a.sample, dev)
        emit(l32)
    if "5" in want:
        per_gpu = n(16777216 // 8)
        b = sc.config5(n_total=16777216, T=300, lo=3 * per_gpu, hi=4 * per_gpu)
        emit(run_config("config5: 2,097,152 of 16,777,216 sweep scenarios (rank 3 of 8), cone barrier, T<=300", b, torch.float64, a.reps, a.sample, dev))
    if a.out:
        with open(a.out, "w") as f:
            for line in lines:
                f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
