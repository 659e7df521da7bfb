#!/usr/bin/env python
"""bench.py -- closed-loop CBF-QP solves/sec (vehicles x obstacles x steps) on 1/2/4/8 B200.

    python bench.py --gpus N --steps K --warmup W
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the CPU restatement of the reference loop, host cores

A "step" is one pass of the hot path over one batch: ONE persistent closed-loop rollout
(Stanley nominal -> barrier rows -> 2-variable QP -> bicycle plant, T = 1,000 timesteps) of
BASELINE config #2 (65,536 vehicles x 8 static ellipses per GPU, synthetic, seed 0).  Scenarios are
independent, so N GPUs = N shards of the scenario axis, no collective on the path (weak scaling:
65,536 vehicles per GPU).

One JSON line on stdout (rank 0):
  value   whole-job solves/s with inputs resident in HBM (CUDA events around each rollout launch)
  e2e     the same metric through ClosedLoopRollout.run_from_host(): pinned host inputs -> H2D ->
          kernel -> D2H of the per-vehicle results, every step
  roofline            the rollout kernel against the MEASURED fp64 FMA peak of this GPU (persistent
                      form: state in registers, ~0.07 B/solve of DRAM traffic -- neither HBM- nor
                      tensor-bound, SURVEY 8d regime ii).  Algorithmic flops are those of the algorithm
                      as implemented: 7 per way-point distance evaluation /
                      capsule test actually made (counted by the kernel, n_evals) + 34 M + 27 per
                      vehicle-step; the 9 transcendental calls per vehicle-step are listed, not counted
  roofline_operator   the per-timestep fused filter kernel (K1+K2) against the measured HBM peak
                      (regime i, 64.6 B/solve)
  cpu_baseline        oracle/oracle.c ("port"; the reference itself cannot run here: no cvxopt)
                      timed on the box's host cores on a bounded sample of the same workload
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "closed-loop CBF-QP solves/sec (vehicles x obstacles x steps)"
UNIT = "solves/s"
P_COURSE = 2034


def exhaustive_flops_per_solve(P: int, M: int) -> float:
    """SURVEY 8(d) regime ii for the REFERENCE's algorithm (exhaustive argmin over all P way-points):
    F = 49 + (7 P + 25) / M fp64 flop per solve.  Reported for continuity only."""
    return 49.0 + (7.0 * P + 25.0) / M


FLOP_PER_EVAL = 7.0            # dx, dy, dx*dx, dy*dy, +, compare (+1); capsule tests are counted at the same rate
FLOP_PER_ROW = 34.0            # static ellipse row with hoisted coefficients (29) + reference-point feasibility test (5)
FLOP_PER_VEHICLE_STEP = 27.0   # P speed loop 2 + Stanley law 8 + update_com plant 17
TRANSCENDENTALS_PER_VEHICLE_STEP = 9   # sincos(yaw), sincos(yaw+pi/2), 4 atan2, 3 tan


def rollout_flops(evals: float, vehicle_steps: float, M: int) -> float:
    """Algorithmic fp64 flops of one rollout launch for the algorithm AS IMPLEMENTED."""
    return FLOP_PER_EVAL * evals + vehicle_steps * (FLOP_PER_ROW * M + FLOP_PER_VEHICLE_STEP)


def operator_bytes_per_solve(M: int, esize: int = 8) -> float:
    """SURVEY 8(d) regime i: 7 obstacle fields per (vehicle, obstacle) + per vehicle state 4 +
    u_ref 2 read, u 2 + mask (4 B) + status (1 B) written."""
    return 7 * esize + ((4 + 2 + 2) * esize + 5) / M


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 8 and parts[0].isdigit() and int(parts[0]) == self.index:
                self.rows.append((time.perf_counter(), parts))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [p for (t, p) in self.rows if t0 <= t <= t1] or [p for (_, p) in self.rows]
        sm, mx, pw, reasons = [], [], [], set()
        for p in rows:
            try:
                sm.append(float(p[1])); mx.append(float(p[2])); pw.append(float(p[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def csrc_sha() -> str:
    """Hash of the kernel sources: profiles/ncu_metrics.json records the one it was captured from."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "sccav_cbf_b200", "csrc")
    for f in sorted(os.listdir(d)):
        if f.endswith((".cu", ".cuh", ".h")):
            h.update(f.encode()); h.update(open(os.path.join(d, f), "rb").read())
    h.update(open(os.path.join(ROOT, "include", "sccav_cbf.h"), "rb").read())
    return h.hexdigest()[:16]


NCU_STALE = [None]


def committed_ncu(kernel: str):
    """Numbers of `kernel` from the committed ncu --set full capture (profiles/ncu_metrics.json):
    dram bytes per launch ("traffic"), fp64-pipe / issue utilisation.  None if absent -- or if the kernels have
    changed since the capture (the JSON records the hash of csrc/ it was taken from): stale numbers are not quoted."""
    path = os.path.join(ROOT, "profiles", "ncu_metrics.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            NCU_STALE[0] = d.get("_csrc_sha") != csrc_sha()
            if NCU_STALE[0]:
                return None
            return d.get(kernel)
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------ extra legs
def _time_rollout(cl, reps, flush):
    """CUDA-event time of `reps` rollout launches of a resident batch (L2 flushed before each)."""
    import torch
    cl.run()
    torch.cuda.synchronize()
    ms = []
    res = None
    for _ in range(reps):
        cl.reset()
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); res = cl.launch(); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return ms, res


def _sample_batch(batch, idx):
    """The scenarios `idx` of a batch as a batch of their own (what the oracle re-runs)."""
    import numpy as np
    from sccav_cbf_b200 import scenarios as sc
    return sc.ScenarioBatch(batch.name + "_sample", np.ascontiguousarray(batch.state[:, idx]), list(batch.slot_desc),
                            None if batch.obst is None else np.ascontiguousarray(batch.obst[:, :, idx]), batch.course,
                            dict(batch.params), T=batch.T,
                            alpha=None if batch.alpha is None else np.ascontiguousarray(batch.alpha[idx]),
                            R=None if batch.R is None else np.ascontiguousarray(batch.R[:, idx]),
                            target_speed=None if batch.target_speed is None else np.ascontiguousarray(batch.target_speed[idx]))


def _oracle_sample(batch, res, n_sample, f32=False):
    """A sample of vehicles taken from INSIDE the full-size batch, re-run alone by the CPU oracle (checker only):
    fraction with identical integer bookkeeping, fraction with the final state within tolerance."""
    import numpy as np
    from oracle import c_oracle as co
    rng = np.random.default_rng(12345)
    idx = np.sort(rng.choice(batch.N, size=min(n_sample, batch.N), replace=False))
    sub = _sample_batch(batch, idx)
    prm = dict(sub.params)
    prm.pop("flags", None)                       # the oracle's arithmetic is always the reference's order
    r = co.rollout(co.default_params(**prm), sub.slot_desc, sub.state, sub.obst, sub.course, sub.T,
                   alpha=sub.alpha, R=sub.R, target_speed=sub.target_speed)
    g = {k: res[k][..., idx].cpu().numpy() for k in ("steps", "target_idx", "n_active", "n_infeasible", "state")}
    same = np.ones(len(idx), dtype=bool)
    for k in ("steps", "target_idx", "n_active", "n_infeasible"):
        same &= g[k] == r[k]
    err = (np.abs(g["state"].astype(np.float64) - r["state"]) / np.maximum(1.0, np.abs(r["state"]))).max(axis=0)
    tol = 5e-2 if f32 else 1e-6
    return float(same.mean()), float((err <= tol).mean()), tol


def run_config_leg(prefix, batch, dtype, flags, dev, flush, sh, reps=2, sample=512, check=True):
    """One BASELINE.json configuration at its full per-GPU size: whole-job solves/s (CUDA events, max over ranks)
    + the oracle-sample parity fractions of rank 0's shard.  Returns FLAT keys (the driver keeps scalars)."""
    import torch
    from sccav_cbf_b200.rollout import ClosedLoopRollout
    if flags:
        batch.params = dict(batch.params, flags=flags)
    cl = ClosedLoopRollout(batch, dtype=dtype, device=dev, pin=False)
    sh.barrier()
    ms, res = _time_rollout(cl, reps, flush)
    ms_rank = min(ms)
    solves = sh.sum(float(res["steps"].sum().item()) * batch.M)
    ms_job = sh.max(ms_rank)
    out = {prefix + "_value": solves / (ms_job * 1e-3), prefix + "_ms": ms_job, prefix + "_vehicles_total": int(sh.sum(float(batch.N))),
           prefix + "_rows_per_vehicle": batch.M, prefix + "_active_step_frac": sh.sum(float(res["n_active"].sum().item())) / max(1.0, sh.sum(float(res["steps"].sum().item())))}
    if check and sh.rank == 0:
        same, ok, tol = _oracle_sample(batch, res, sample, f32=(dtype == torch.float32))
        out[prefix + "_oracle_identical_bookkeeping_frac"] = same
        out[prefix + "_oracle_state_within_tol_frac"] = ok
        out[prefix + "_oracle_sample"] = min(sample, batch.N)
    del cl
    torch.cuda.empty_cache()
    return out


def python_loop_rate(seconds, M, T):
    """What a reference user experiences minus cvxopt's IPM and matplotlib: the scalar Python restatement of the
    loop (oracle/oracle.py, structured like stanley_controller_ellipse.py:630-830), 1 core, whole vehicles until
    `seconds` are spent."""
    import numpy as np
    from oracle import oracle as o
    from sccav_cbf_b200 import scenarios as sc
    b = sc.config2(n_total=65536, M=M, T=T, seed=0, lo=0, hi=16)
    course = tuple(np.asarray(c) for c in b.course)
    t0 = time.perf_counter()
    solves, n = 0.0, 0
    while n < b.N and (n == 0 or time.perf_counter() - t0 < seconds):
        r = o.rollout(b.state[:, n], [int(d) & 0x3f for d in b.slot_desc], [b.obst[m, :, n] for m in range(M)], course, T)
        solves += r["steps"] * M
        n += 1
    dt = time.perf_counter() - t0
    return solves / dt, n, dt


def solve_cbf_latency(dev, n, reps=200):
    """Per-call latency of the class API as the reference uses it: one DBM_CBF_2DS.solve_cbf per tick on a list of
    8 ellipses (cbf/cbf.py:166-220), batch size n; microseconds per call, result synchronised every call."""
    import torch
    from sccav_cbf_b200 import DBM_CBF_2DS, Ellipse2D
    from sccav_cbf_b200.euclid import Vector2
    g = torch.Generator(device="cpu"); g.manual_seed(7)
    cbf = DBM_CBF_2DS(alpha=1.0)
    cbf.set_model_params(1.45, 1.45)
    for k in range(8):
        if n == 1:
            e = Ellipse2D(3.0 + 0.2 * k, 1.5, Vector2(12.0 + 5.0 * k, 4.0 - k), theta=0.1 * k, buffer=0.5)
        else:
            cx = (12.0 + 5.0 * k + torch.rand(n, generator=g, dtype=torch.float64)).to(dev)
            cy = (4.0 - k + torch.rand(n, generator=g, dtype=torch.float64)).to(dev)
            e = Ellipse2D(3.0 + 0.2 * k, 1.5, Vector2(cx, cy), theta=0.1 * k, buffer=0.5)
        cbf.obstacle_list2d[k] = e
    if n == 1:
        s = [0.0, 5.0, 0.35, 10.0]
        u_ref = [0.3, 0.05]
    else:
        s = torch.tensor([[0.0], [5.0], [0.35], [10.0]], dtype=torch.float64, device=dev).repeat(1, n)
        u_ref = torch.tensor([[0.3], [0.05]], dtype=torch.float64, device=dev).repeat(1, n)
    for _ in range(10):
        cbf.update_state(s)
        u = cbf.solve_cbf(u_ref)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        cbf.update_state(s)
        u = cbf.solve_cbf(u_ref)
        torch.cuda.synchronize()
    return 1e6 * (time.perf_counter() - t0) / reps


# ------------------------------------------------------------------------------------------ reference arm
def cpu_rollout_rate(n_vehicles: int, M: int, T: int, threads: int, seed: int = 0):
    """oracle/oracle.c closed loop on the first n_vehicles scenarios of config #2."""
    from oracle import c_oracle as co
    from sccav_cbf_b200 import scenarios as sc
    b = sc.config2(n_total=65536, M=M, T=T, seed=seed, lo=0, hi=n_vehicles)
    prm = co.default_params(**b.params)
    t0 = time.perf_counter()
    res = co.rollout(prm, b.slot_desc, b.state, b.obst, b.course, T, nthreads=threads)
    dt = time.perf_counter() - t0
    solves = float(res["steps"].sum()) * M
    return solves / dt, dt, solves


def run_reference(args):
    """`--impl reference`: the reference's own CPU loop cannot run here (cvxopt / euclid are not
    installable: no network), so this arm times the oracle port of that loop (oracle/oracle.c,
    all host threads) on the same config, metric and unit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import c_oracle as co
    threads = co.num_threads()
    M, T = args.obstacles, args.T
    rate, dt, _ = cpu_rollout_rate(max(threads, 8), M, T, threads)            # probe (also warms the page cache)
    n = int(max(threads, min(65536, rate * args.ref_seconds / (M * T))))
    for _ in range(args.warmup if args.warmup < 2 else 1):                     # CPU warm-up: one pass is enough
        cpu_rollout_rate(n, M, T, threads)
    t_tot, s_tot = 0.0, 0.0
    for _ in range(args.steps):
        r, d, s = cpu_rollout_rate(n, M, T, threads)
        t_tot += d; s_tot += s
    value = s_tot / t_tot
    sample = "%d of 65536 vehicles x %d ellipses x %d steps per step (config #2 prefix), %d host threads" % (n, M, T, threads)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args), "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "reference loop itself not runnable: cvxopt/euclid absent and not installable (no network)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_name(args):
    return ("config2: %d vehicles/GPU x %d static ellipses x %d-step closed-loop rollout "
            "(DBM + Stanley + update_com, P=%d course points)" % (args.vehicles, args.obstacles, args.T, P_COURSE))


# ------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--vehicles", type=int, default=65536, help="vehicles per GPU")
    ap.add_argument("--obstacles", type=int, default=8)
    ap.add_argument("--T", type=int, default=1000, help="timesteps per rollout")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--ref-seconds", type=float, default=4.0, help="CPU seconds per reference step")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU seconds for the cpu_baseline sample")
    ap.add_argument("--rows", default="fast", choices=["fast", "prepared", "canonical"],
                    help="arithmetic of the rollout: 'canonical' = the reference's operation order at every step; 'prepared' = "
                         "SCCAV_FLAG_PREPARED_ROWS (obstacles ingested once per launch, rows evaluated in their prepared form); "
                         "'fast' = prepared rows + SCCAV_FLAG_FUSED_STEER (beta = clamp(beta*) instead of beta* -> delta -> clip -> "
                         "beta).  All three are fp64 and compute the same functions -- a few ulp apart; the canonical timing and the "
                         "fraction of vehicles with identical integer bookkeeping are always reported beside the timed mode")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-operator", action="store_true", help="skip the regime-(i) operator roofline leg")
    ap.add_argument("--no-configs", action="store_true", help="skip the BASELINE configs 3 / 4 / 5 / 1M-vehicle-target legs")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the process to the CPUs of its GPU's NUMA node")
    ap.add_argument("--config-scale", type=float, default=1.0, help="shrink the vehicle counts of the extra config legs (smoke runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch

    from sccav_cbf_b200 import ops, scenarios as sc
    from sccav_cbf_b200.dist import Shards, bind_to_gpu_numa_node
    from sccav_cbf_b200.rollout import ClosedLoopRollout

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback on the product path)")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    all_cpus = os.sched_getaffinity(0)
    numa_cpus = None if args.no_numa_bind else bind_to_gpu_numa_node(local)   # before any pinned buffer is allocated
    sh = Shards(backend="nccl", device=dev)
    rank, world = sh.rank, sh.world

    dtype = torch.float64 if args.dtype == "f64" else torch.float32
    esize = 8 if args.dtype == "f64" else 4
    M, T, NV = args.obstacles, args.T, args.vehicles
    n_total = NV * world
    lo, hi = sc.shard_range(n_total, rank, world)
    batch = sc.config2(n_total=n_total, M=M, T=T, seed=0, lo=lo, hi=hi)
    ROW_FLAGS = {"canonical": 0, "prepared": 1, "fast": 5}       # SCCAV_FLAG_PREPARED_ROWS = 1, SCCAV_FLAG_FUSED_STEER = 4
    if ROW_FLAGS[args.rows]:
        batch.params = dict(batch.params, flags=ROW_FLAGS[args.rows])
    cl = ClosedLoopRollout(batch, dtype=dtype, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)       # > 126 MB L2

    # ---- warm-up
    for _ in range(args.warmup):
        cl.run()
    torch.cuda.synchronize()

    # ---- timed region 1: inputs resident in HBM, one rollout launch per step
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    sh.barrier()
    l0 = ops.launch_count()
    t_wall0 = time.perf_counter()
    for e0, e1 in evs:
        flush.fill_(1)                    # L2 flush between timed iterations (outside the event pair)
        e0.record()
        res = cl.run()
        e1.record()
    sh.barrier()
    t_wall1 = time.perf_counter()
    launches = ops.launch_count() - l0
    clocks = sampler.stop(t_wall0, t_wall1)
    ms_steps = [e0.elapsed_time(e1) for e0, e1 in evs]
    ms_total = sh.max(sum(ms_steps))
    vehicle_steps_rank = float(res["steps"].sum().item())
    evals_rank = float(res["n_evals"].to(torch.float64).sum().item())
    solves_per_step_rank = vehicle_steps_rank * M
    solves_per_step = sh.sum(solves_per_step_rank)
    value = solves_per_step * args.steps / (ms_total * 1e-3)
    checksum = sh.sum(float(res["n_active"].sum().item()))

    # ---- the other row mode, for the record (untimed region; 3 launches, same events)
    other = None
    fp32_variant = None
    if rank == 0 or world > 1:
        b2 = sc.config2(n_total=n_total, M=M, T=T, seed=0, lo=lo, hi=hi)
        if args.rows == "canonical":
            b2.params = dict(b2.params, flags=ROW_FLAGS["fast"])
        cl2 = ClosedLoopRollout(b2, dtype=dtype, device=dev, pin=False)
        cl2.run()
        ms2 = []
        for _ in range(3):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r2 = cl2.run(); e1.record()
            torch.cuda.synchronize()
            ms2.append(e0.elapsed_time(e1))
        same = (r2["steps"] == res["steps"]) & (r2["target_idx"] == res["target_idx"]) & (r2["n_active"] == res["n_active"]) \
            & (r2["n_infeasible"] == res["n_infeasible"])
        other = {"rows": "canonical" if args.rows != "canonical" else "fast", "ms_per_step": statistics.mean(ms2),
                 "value_this_rank": float(r2["steps"].sum().item()) * M / (statistics.mean(ms2) * 1e-3),
                 "identical_bookkeeping_frac_vs_timed_mode": float(same.double().mean().item())}
        del cl2
        # ---- the fp32 variant of the same rollout (north_star: fp64 by default, fp32 reported): timing and how far
        # it drifts from the fp64 run of the timed mode (same flags)
        if args.dtype == "f64":
            b3 = sc.config2(n_total=n_total, M=M, T=T, seed=0, lo=lo, hi=hi)
            b3.params = dict(batch.params)
            cl3 = ClosedLoopRollout(b3, dtype=torch.float32, device=dev, pin=False)
            cl3.run()
            ms3 = []
            for _ in range(3):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); r3 = cl3.run(); e1.record()
                torch.cuda.synchronize()
                ms3.append(e0.elapsed_time(e1))
            same3 = (r3["steps"] == res["steps"]) & (r3["target_idx"] == res["target_idx"]) & (r3["n_active"] == res["n_active"]) \
                & (r3["n_infeasible"] == res["n_infeasible"])
            dpos = (r3["state"][:2].double() - res["state"][:2].double()).norm(dim=0)
            fp32_variant = {"dtype": "f32", "rows": args.rows, "ms_per_step": statistics.mean(ms3),
                            "value_this_rank": float(r3["steps"].sum().item()) * M / (statistics.mean(ms3) * 1e-3),
                            "identical_bookkeeping_frac_vs_f64": float(same3.double().mean().item()),
                            "final_position_diff_m_median": float(dpos.median().item()),
                            "final_position_diff_m_p99": float(torch.quantile(dpos, 0.99).item())}
            del cl3

    # ---- timed region 2: end to end through the C-ABI host entry point (sccav_rollout_host_*):
    #      pinned HOST buffers in, H2D + kernel + D2H inside the call, HOST results out -- every step
    for _ in range(2):
        cl.run_from_host()
    sh.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        hres = cl.run_from_host()
    torch.cuda.synchronize()
    e2e_s = sh.max(time.perf_counter() - t0)
    sh.barrier()
    e2e_sync_value = solves_per_step * args.steps / e2e_s
    assert int(hres["steps"].sum().item()) == int(vehicle_steps_rank)
    # ---- the same through the PIPELINED host API (sccav_pipeline_*): every step still uploads ALL of its inputs (states
    #      and obstacles, pinned host memory) and downloads its results; the course is resident in the handle, and with two
    #      submissions in flight the copies of step i+1 / i-1 run beside the kernel of step i
    pipe = cl.pipeline(depth=2)
    cl.run_pipelined(pipe, 2)
    sh.barrier()
    t0 = time.perf_counter()
    pres = cl.run_pipelined(pipe, args.steps)
    e2e_pipe_s = sh.max(time.perf_counter() - t0)
    sh.barrier()
    assert int(pres["steps"].sum().item()) == int(vehicle_steps_rank)
    pipe.close()
    e2e_value = solves_per_step * args.steps / e2e_pipe_s

    # ---- the other BASELINE.json configurations at their full per-GPU sizes (flat keys: whole-job solves/s, ms,
    #      oracle-sample parity of rank 0's shard).  Every rank runs its shard; no collective on the path.
    flat = {}
    if other is not None:
        canon = other if other["rows"] == "canonical" else None
        if canon is not None:
            flat["canonical_ms_per_step"] = sh.max(canon["ms_per_step"])
            flat["canonical_value"] = solves_per_step / (flat["canonical_ms_per_step"] * 1e-3)
        flat["identical_bookkeeping_frac"] = other["identical_bookkeeping_frac_vs_timed_mode"]
    if fp32_variant is not None:
        flat["fp32_ms_per_step"] = sh.max(fp32_variant["ms_per_step"])
        flat["fp32_value"] = solves_per_step / (flat["fp32_ms_per_step"] * 1e-3)
        flat["fp32_identical_bookkeeping_frac_vs_f64"] = fp32_variant["identical_bookkeeping_frac_vs_f64"]
    if not args.no_configs and args.dtype == "f64":
        fl = ROW_FLAGS[args.rows]
        sc_ = args.config_scale

        def per_gpu(n):
            return max(1024, int(n * sc_))
        # config 3: 262,144 vehicles x 16 moving circles (radial-dynamic TV-CBF, seekers), 600 frames
        n3 = per_gpu(262144)
        lo3, hi3 = sc.shard_range(n3 * world, rank, world)
        # (config 3's fast mode adds SCCAV_FLAG_SEEKER_DIRECT = 16: seeker headings as normalised offsets)
        flat.update(run_config_leg("config3", sc.config3(n_total=n3 * world, M=16, T=600, lo=lo3, hi=hi3), dtype, (fl | 16) if fl else 0, dev, flush, sh))
        flat["config3_flags"] = (fl | 16) if fl else 0
        # config 4: 1,048,576 vehicles x (8 ellipses + 2 lanes), fp64 and fp32
        n4 = per_gpu(1048576)
        lo4, hi4 = sc.shard_range(n4 * world, rank, world)
        b4 = sc.config4(n_total=n4 * world, M=8, T=1000, lo=lo4, hi=hi4)
        flat.update(run_config_leg("config4_f64", b4, dtype, fl, dev, flush, sh))
        flat.update(run_config_leg("config4_f32", b4, torch.float32, fl, dev, flush, sh))
        del b4
        # config 5, weak: 2,097,152 sweep scenarios per GPU (shard `rank` of 8 x that many)
        n5 = per_gpu(16777216 // 8)
        b5 = sc.config5(n_total=8 * n5, T=300, lo=rank * n5, hi=(rank + 1) * n5)
        flat.update(run_config_leg("config5", b5, dtype, fl, dev, flush, sh))
        del b5
        # config 5, STRONG: the whole 16,777,216-scenario sweep split over the ranks of this job
        n5s = per_gpu(16777216)
        lo5, hi5 = sc.shard_range(n5s, rank, world)
        flat.update(run_config_leg("config5_strong", sc.config5(n_total=n5s, T=300, lo=lo5, hi=hi5), dtype, fl, dev, flush, sh, check=False))
        # the north-star target shard: 1,048,576 vehicles x 8 ellipses over 8 GPUs = 131,072 vehicles per GPU
        nt = per_gpu(131072)
        lot, hit = sc.shard_range(nt * world, rank, world)
        flat.update(run_config_leg("target1m", sc.config2(n_total=nt * world, M=8, T=1000, seed=0, lo=lot, hi=hit), dtype, fl, dev, flush, sh))
        # Monte-Carlo over ROADS: 64 roads x 1,024 vehicles x 8 ellipses x 1000 steps in ONE launch (sccav_rollout_roads_*)
        # against 64 launches with one road each (what the road-per-launch interface costs); rank 0 only
        if rank == 0:
            n_roads, per_road = 64, max(32, int(1024 * sc_))
            (rcx, rcy, rcyaw, rnp), rnph, rstate, robst = sc.roads(n_roads, per_road, M=8, seed=4, dtype=dtype, device=dev)
            rprm = ops.make_params(flags=fl)
            rsd = [0x40] * 8                                              # SLOT_ELLIPSE | SLOT_STATIC
            d_rs = torch.from_numpy(rstate).to(device=dev, dtype=dtype); d_ro = torch.from_numpy(robst).to(device=dev, dtype=dtype)
            rout = {}
            ops.rollout(rprm, rsd, d_rs, d_ro, (rcx, rcy, rcyaw), 1000, course_np=rnp, out=rout)
            torch.cuda.synchronize()
            rms = []
            for _ in range(2):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); rr = ops.rollout(rprm, rsd, d_rs, d_ro, (rcx, rcy, rcyaw), 1000, course_np=rnp, out=rout); e1.record()
                torch.cuda.synchronize()
                rms.append(e0.elapsed_time(e1))
            rsolves = float(rr["steps"].sum().item()) * 8
            singles = [(rcx[c, :rnph[c]].contiguous(), rcy[c, :rnph[c]].contiguous(), rcyaw[c, :rnph[c]].contiguous(),
                        d_rs[:, c * per_road:(c + 1) * per_road].contiguous(), d_ro[:, :, c * per_road:(c + 1) * per_road].contiguous())
                       for c in range(n_roads)]
            souts = [{} for _ in range(n_roads)]
            for rep in range(2):
                flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for c, (a_, b_, c_, s_, o_) in enumerate(singles):
                    ops.rollout(rprm, rsd, s_, o_, (a_, b_, c_), 1000, out=souts[c])
                e1.record()
                torch.cuda.synchronize()
                sms = e0.elapsed_time(e1)
            same = all(torch.equal(souts[c]["state"], rr["state"][:, c * per_road:(c + 1) * per_road]) for c in range(n_roads))
            flat.update({"roads64_value": rsolves / (min(rms) * 1e-3), "roads64_ms": min(rms), "roads64_vehicles": n_roads * per_road,
                         "roads64_one_launch_per_road_ms": sms, "roads64_bit_identical_to_single_road_launches": bool(same)})
            del singles, souts, rout, d_rs, d_ro
        flat["config_legs_flags"] = fl
    # ---- roofline of the dominant kernel (rollout): fp64 FMA peak measured on this GPU, now
    peak_tf = ops.measure_fma_peak(dtype)
    flops_per_launch = rollout_flops(evals_rank, vehicle_steps_rank, M)
    ms_launch = sum(ms_steps) / len(ms_steps)
    ach_tf = flops_per_launch / (ms_launch * 1e-3) / 1e12
    tname = "double" if args.dtype == "f64" else "float"
    ncu_roll = committed_ncu("rollout_kernel<%s>" % tname) or {}
    info = ops.rollout_launch_info(batch.slot_desc, batch.N, P_COURSE, dtype)
    hbm_bytes = batch.N * ((4 + M * 7) * esize + M * 6 * esize * 2 + 4 * esize + 5 * 4 + 4 * esize)
    roofline = {
        "bound": "fp64" if args.dtype == "f64" else "fp32", "kernel": "rollout_kernel<%s>" % tname,
        "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach_tf / peak_tf if peak_tf else None,
        "peak_source": "measured live: sccav_measure_fma_peak (unrolled FMA chains, FMA = 2 flop; the reference-order fp64 "
                       "arithmetic is compiled without FMA contraction for parity -- only the prepared ellipse rows use fma)",
        "algorithmic_flops_per_launch": flops_per_launch,
        "algorithmic_flops_per_solve": flops_per_launch / solves_per_step_rank,
        "flops_model": "7 x n_evals (counted by the kernel) + vehicle_steps x (34 M + 27); %d transcendental calls per "
                       "vehicle-step not counted" % (TRANSCENDENTALS_PER_VEHICLE_STEP - (4 if args.rows == "fast" else 0)),
        "nearest_search_evals_per_vehicle_step": evals_rank / max(vehicle_steps_rank, 1.0),
        "exhaustive_scan_flops_per_solve": exhaustive_flops_per_solve(P_COURSE, M),
        "traffic": ncu_roll.get("dram_bytes"),
        "fp64_pipe_active_pct_ncu": ncu_roll.get("fp64_pipe_active_pct"),
        "issue_active_pct_ncu": ncu_roll.get("issue_active_pct"),
        "tensor_pipe_active_pct_ncu": ncu_roll.get("tensor_pipe_active_pct"),
        "hbm_gbs_achieved": hbm_bytes / (ms_launch * 1e-3) / 1e9,
        "launch": info,
        "note": "not HBM- or tensor-bound: state lives in registers for the whole rollout; DRAM traffic is the one-off "
                "read of the inputs (about 0.07 B/solve); the limiter is instruction issue / fp64 latency at 14 warps per SM",
    }

    # ---- regime (i): per-timestep fused operator against the HBM roofline (inputs >> L2)
    roofline_op = None
    if not args.no_operator:
        hbm_peak, hbm_src = measured_peaks()
        n_op = 2 * 1024 * 1024
        gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)
        st = torch.empty((4, n_op), dtype=dtype, device=dev)
        st[0].uniform_(-20, 120, generator=gen); st[1].uniform_(-45, 15, generator=gen)
        st[2].uniform_(-3.2, 3.2, generator=gen); st[3].uniform_(2, 12, generator=gen)
        ob = torch.zeros((M, 8, n_op), dtype=dtype, device=dev)
        ob[:, 0].uniform_(-20, 120, generator=gen); ob[:, 1].uniform_(-45, 15, generator=gen)
        ob[:, 2].uniform_(2.5, 6.5, generator=gen); ob[:, 3].uniform_(1.5, 3.5, generator=gen)
        ob[:, 4].uniform_(-3.1, 3.1, generator=gen)
        ur = torch.zeros((2, n_op), dtype=dtype, device=dev); ur[1].uniform_(-0.3, 0.3, generator=gen)
        prm = ops.make_params()
        for _ in range(3):
            ops.filter_step(prm, batch.slot_desc, st, ob, ur)
        oevs = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.filter_step(prm, batch.slot_desc, st, ob, ur); e1.record()
            oevs.append((e0, e1))
        torch.cuda.synchronize()
        oms = statistics.mean(e0.elapsed_time(e1) for e0, e1 in oevs)
        obytes = operator_bytes_per_solve(M, esize) * n_op * M
        ncu_op = committed_ncu("filter_step_kernel<%s>" % tname) or {}
        roofline_op = {
            "bound": "hbm", "kernel": "filter_step_kernel<%s>" % tname,
            "achieved": obytes / (oms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
            "frac": obytes / (oms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src,
            "algorithmic_bytes_per_solve": operator_bytes_per_solve(M, esize), "solves_per_s": n_op * M / (oms * 1e-3),
            "workload": "%d vehicles x %d ellipses, inputs %.0f MB (> L2)" % (n_op, M, obytes / 1e6),
            "traffic": ncu_op.get("dram_bytes"),
        }
        # the same solve on prepared obstacles (ingest once with sccav_prepare_obstacles_*, solve every tick):
        # 6 fields per static (vehicle, obstacle) instead of 7, no sincos / division per row
        sdp, obp = ops.prepare_obstacles([d | 0x40 for d in batch.slot_desc], ob)
        for _ in range(3):
            ops.filter_step(prm, sdp, st, obp, ur)
        pevs = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.filter_step(prm, sdp, st, obp, ur); e1.record()
            pevs.append((e0, e1))
        torch.cuda.synchronize()
        pms = statistics.mean(e0.elapsed_time(e1) for e0, e1 in pevs)
        pbps = 6 * esize + ((4 + 2 + 2) * esize + 5) / M
        ncu_st = committed_ncu("filter_step_staged_kernel<%s>" % tname) or {}
        u_p, _, st_p, _ = ops.filter_step(prm, sdp, st, obp, ur)
        roofline_op["prepared"] = {
            "kernel": "filter_step_staged_kernel<%s, ELLIPSE_PREP, 6> (static slots)" % tname, "algorithmic_bytes_per_solve": pbps,
            "achieved": pbps * n_op * M / (pms * 1e-3) / 1e9, "frac": pbps * n_op * M / (pms * 1e-3) / 1e9 / hbm_peak,
            "solves_per_s": n_op * M / (pms * 1e-3), "ms": pms, "traffic": ncu_st.get("dram_bytes"),
            "fp64_pipe_active_pct_ncu": ncu_st.get("fp64_pipe_active_pct"), "issue_active_pct_ncu": ncu_st.get("issue_active_pct"),
            "active_frac": float((st_p == 1).double().mean().item()), "infeasible_frac": float((st_p == 2).double().mean().item()),
        }
        # the same prepared solve with beta in / beta out (SCCAV_FLAG_BETA_IO): callers that integrate their plant in beta
        # (State.update_com) skip the delta <-> beta conversions of cbf.py:175,216 -- two tan + two atan2 per solve
        from sccav_cbf_b200 import _native as _nv
        prm_b = ops.make_params(flags=_nv.FLAG_BETA_IO)
        ur_b = ur.clone()
        ur_b[1] = torch.atan2(prm.lr * torch.tan(ur[1]), torch.full_like(ur[1], prm.lf + prm.lr))
        for _ in range(3):
            u_b, _, st_b, _ = ops.filter_step(prm_b, sdp, st, obp, ur_b)
        bevs = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.filter_step(prm_b, sdp, st, obp, ur_b); e1.record()
            bevs.append((e0, e1))
        torch.cuda.synchronize()
        bms = statistics.mean(e0.elapsed_time(e1) for e0, e1 in bevs)
        beta_of_u = torch.atan2(prm.lr * torch.tan(u_p[1]), torch.full_like(u_p[1], prm.lf + prm.lr))
        roofline_op["prepared_beta_io"] = {
            "kernel": roofline_op["prepared"]["kernel"] + ", SCCAV_FLAG_BETA_IO", "algorithmic_bytes_per_solve": pbps,
            "achieved": pbps * n_op * M / (bms * 1e-3) / 1e9, "frac": pbps * n_op * M / (bms * 1e-3) / 1e9 / hbm_peak,
            "solves_per_s": n_op * M / (bms * 1e-3), "ms": bms,
            "statuses_equal_delta_io": bool(torch.equal(st_b, st_p)),
            "max_abs_beta_diff_vs_delta_io": float((u_b[1] - beta_of_u).abs().max().item()),
        }
        del ur_b, u_b, st_b, beta_of_u
        # the same batch with the colliding pairs removed: an obstacle whose ellipse already contains its vehicle
        # (h < 0.05 -- a crash, not a tick the filter is meant for) is moved 1 km away.  The random batch above
        # keeps them (5.7 % of its problems have contradictory rows), which makes it QP-heavy on purpose.
        h_all = ops.barrier_partials(batch.slot_desc, st, ob)[:, 0]
        ob_cf = ob.clone()
        ob_cf[:, 0] = torch.where(h_all < 0.05, ob[:, 0] + 1000.0, ob[:, 0])
        _, obp_cf = ops.prepare_obstacles([d | 0x40 for d in batch.slot_desc], ob_cf)
        for _ in range(3):
            _, _, st_cf, _ = ops.filter_step(prm, sdp, st, obp_cf, ur)
        cevs = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.filter_step(prm, sdp, st, obp_cf, ur); e1.record()
            cevs.append((e0, e1))
        torch.cuda.synchronize()
        cms = statistics.mean(e0.elapsed_time(e1) for e0, e1 in cevs)
        roofline_op["prepared_collision_free"] = {
            "kernel": roofline_op["prepared"]["kernel"], "algorithmic_bytes_per_solve": pbps,
            "achieved": pbps * n_op * M / (cms * 1e-3) / 1e9, "frac": pbps * n_op * M / (cms * 1e-3) / 1e9 / hbm_peak,
            "solves_per_s": n_op * M / (cms * 1e-3), "ms": cms,
            "active_frac": float((st_cf == 1).double().mean().item()), "infeasible_frac": float((st_cf == 2).double().mean().item()),
        }
        del h_all, ob_cf, obp_cf
        # the same batch through the CLASS API (what a user of the reference calls every tick): DBM_CBF_2DS with 8
        # Ellipse2D objects holding [N] tensors; the first call packs and ingests them, later calls reuse the image
        try:
            from sccav_cbf_b200 import DBM_CBF_2DS, Ellipse2D
            from sccav_cbf_b200.euclid import Vector2
            cbf = DBM_CBF_2DS(alpha=1.0)
            cbf.set_model_params(1.45, 1.45)
            for m_ in range(M):
                cbf.obstacle_list2d[m_] = Ellipse2D(ob[m_, 2], ob[m_, 3], Vector2(ob[m_, 0], ob[m_, 1]), theta=ob[m_, 4], buffer=0)
            cbf.update_state(st)
            for _ in range(3):
                u_c = cbf.solve_cbf(ur)
            aevs = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); cbf.update_state(st); u_c = cbf.solve_cbf(ur); e1.record()
                aevs.append((e0, e1))
            torch.cuda.synchronize()
            ams = statistics.mean(e0.elapsed_time(e1) for e0, e1 in aevs)
            roofline_op["class_api"] = {
                "call": "DBM_CBF_2DS.update_state + solve_cbf, 8 Ellipse2D with [N] tensor fields (cached ingested image)",
                "algorithmic_bytes_per_solve": pbps, "achieved": pbps * n_op * M / (ams * 1e-3) / 1e9,
                "frac": pbps * n_op * M / (ams * 1e-3) / 1e9 / hbm_peak, "ms": ams,
                "max_abs_diff_vs_operator": float((u_c - u_p).abs().max().item())}
            del cbf, u_c
        except Exception as exc:
            roofline_op["class_api"] = {"error": repr(exc)[:300]}
        ops.prepare_obstacles(batch.slot_desc, ob, out=obp)
        ievs = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.prepare_obstacles(batch.slot_desc, ob, out=obp); e1.record()
            ievs.append((e0, e1))
        torch.cuda.synchronize()
        ims = statistics.mean(e0.elapsed_time(e1) for e0, e1 in ievs)
        roofline_op["ingest"] = {"kernel": "prepare_obstacles_vec2_kernel<%s> (two vehicles per thread, 16-byte accesses)" % tname, "bytes_per_slot": 16 * esize,
                                 "achieved": 16 * esize * n_op * M / (ims * 1e-3) / 1e9,
                                 "frac": 16 * esize * n_op * M / (ims * 1e-3) / 1e9 / hbm_peak, "ms": ims}
        del st, ob, ur, obp, u_p, st_p

    # ---- per-call latency of the class API (one solve_cbf per tick, as the reference uses it)
    if rank == 0:
        try:
            flat["solve_cbf_latency_us_n1"] = solve_cbf_latency(dev, 1)
            flat["solve_cbf_latency_us_n1024"] = solve_cbf_latency(dev, 1024)
        except Exception as exc:          # keep the bench line even if the class API leg fails
            flat["solve_cbf_latency_error"] = repr(exc)[:200]

    # ---- CPU baseline (rank 0, N = 1 only): the oracle port on the box's host cores, bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import c_oracle as co
        os.sched_setaffinity(0, all_cpus)          # the CPU arm gets every host core back
        threads = co.num_threads()
        rate, _, _ = cpu_rollout_rate(max(threads, 8), M, T, threads)
        n = int(max(threads, min(NV, rate * args.cpu_seconds / (M * T))))
        rate, dt_cpu, _ = cpu_rollout_rate(n, M, T, threads)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": "first %d of %d vehicles x %d ellipses x %d steps (%.1f s), oracle/oracle.c, gcc -O2, %d threads"
                         % (n, NV, M, T, dt_cpu, threads),
               "note": "reference loop not runnable here (cvxopt/euclid absent, no network); the port solves each QP exactly "
                       "instead of paying cvxopt's python-callback interior-point iterations, so it flatters the reference"}
        # what a reference USER runs is a Python loop: the scalar Python restatement (no cvxopt IPM, no matplotlib), 1 core
        py_rate, py_n, py_dt = python_loop_rate(4.0, M, T)
        cpu["python_loop_value"] = py_rate
        cpu["python_loop_sample"] = "%d vehicles x %d ellipses x %d steps in %.1f s, oracle/oracle.py, 1 core" % (py_n, M, T, py_dt)
        flat["cpu_python_loop_value"] = py_rate
        flat["cpu_port_value"] = rate

    if roofline_op is not None:
        flat["operator_canonical_hbm_frac"] = roofline_op["frac"]
        flat["operator_prepared_hbm_frac"] = roofline_op["prepared"]["frac"]
        flat["operator_prepared_ms"] = roofline_op["prepared"]["ms"]
        flat["operator_prepared_beta_io_hbm_frac"] = roofline_op["prepared_beta_io"]["frac"]
        flat["operator_prepared_beta_io_ms"] = roofline_op["prepared_beta_io"]["ms"]
        if "frac" in roofline_op.get("class_api", {}):
            flat["class_api_hbm_frac"] = roofline_op["class_api"]["frac"]
            flat["class_api_ms"] = roofline_op["class_api"]["ms"]
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": workload_name(args), "vehicles_total": n_total, "obstacles": M, "timesteps": T,
                       "solves_per_step": solves_per_step, "parallelism": "scenario shards x%d, no collective" % world,
                       "l2": "flushed (256 MB write) between timed iterations; inputs 36 MB/GPU",
                       "rows": args.rows, "other_row_mode": other, "fp32_variant": fp32_variant,
                       "seed": 0, "active_step_checksum": checksum},
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": cl.h2d_bytes() - sum(t.numel() * t.element_size() for t in (cl.h_course or ())),
                    "d2h_bytes_per_step": cl.d2h_bytes(), "ms_per_step": 1e3 * e2e_pipe_s / args.steps,
                    "path": "sccav_pipeline_submit/wait_%s (C-ABI, pinned host pointers, 2 submissions in flight): H2D of the step's "
                            "states + obstacles, rollout kernel, D2H of its results -- every step; the course is resident" % args.dtype,
                    "synchronous_value": e2e_sync_value, "synchronous_ms_per_step": 1e3 * e2e_s / args.steps,
                    "synchronous_path": "sccav_rollout_host_%s: one blocking call per step, H2D + kernel + D2H in sequence" % args.dtype},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "roofline_operator": roofline_op,
            "cpu_baseline": cpu,
            "wall_s_timed_region": t_wall1 - t_wall0,
            "host_cpu_affinity": ("%d CPUs of the GPU's NUMA node (NVML ideal affinity)" % len(numa_cpus)) if numa_cpus else "unchanged",
            "vs_reference_note": "N GPUs over ONE host's CPU threads when n_gpus > 1 (the CPU arm does not scale with --gpus)",
        }
        flat["e2e_sync_value"] = e2e_sync_value
        flat["e2e_over_value"] = e2e_value / value
        line.update(flat)
        line["ncu_metrics_stale"] = NCU_STALE[0]
        print(json.dumps(line), flush=True)
    sh.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
