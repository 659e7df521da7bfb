#!/usr/bin/env python
"""Summarise ncu artefacts brought back in gpurun_out/ into small text files for profiles/.

    python scripts/ncu_summary.py launches gpurun_out/<tag>/launches.csv  > profiles/<tag>_launches.txt
    python scripts/ncu_summary.py report   gpurun_out/<tag>/prof_x.ncu-rep > profiles/<tag>_prof_x.txt

`report` needs the ncu CLI (present in the build container; no GPU needed to read a report).
"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = [
    "gpu__time_duration.sum",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__inst_executed.avg.per_cycle_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_wait",
    "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_selected",
    "smsp__pcsamp_warps_issue_stalled_branch_resolving", "smsp__pcsamp_warps_issue_stalled_mio_throttle",
    "smsp__pcsamp_warps_issue_stalled_lg_throttle", "smsp__pcsamp_warps_issue_stalled_barrier",
    "smsp__pcsamp_warps_issue_stalled_dispatch_stall", "smsp__pcsamp_warps_issue_stalled_no_instructions",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = defaultdict(lambda: [0, 0.0])
    order = []
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1.0)
        name = r[ki]
        if name not in agg:
            order.append(name)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none : per-kernel totals (cold-cache, serialised)")
    print("# source: %s ; %d launches, %.3f ms total" % (path, sum(v[0] for v in agg.values()), tot))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%10.3f ms  %5.1f%%  n=%-4d avg %9.4f ms  %s" % (v[1], 100 * v[1] / tot, v[0], v[1] / v[0], k[:110]))


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    rows = [r for r in rows if r]
    hdr, units = rows[0], rows[1]
    print("# ncu --set full --clock-control none : selected raw metrics, one block per captured launch")
    print("# source: %s" % path)
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(units, vals)))
        print("\n== %s  grid %s block %s" % (d.get("Kernel Name", ("", "?"))[1][:100], d.get("Grid Size", ("", "?"))[1], d.get("Block Size", ("", "?"))[1]))
        for k in KEYS:
            for h in hdr:
                if h == k or (k.startswith("smsp__pcsamp") and h == k):
                    u, v = d[h]
                    print("%-86s %-12s %s" % (h, u, v))


if __name__ == "__main__" and sys.argv[1] in ("launches", "report"):
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])


def source(path, top=45):
    """Hot source lines (needs -lineinfo): samples, executed warp instructions, thread efficiency."""
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur = None
    recs = []
    hdr = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or cur is None or r[0] in ("Function Name", "Kernel Name") or not r[0].isdigit():
            continue
        d = dict(zip(hdr, r))
        def num(k):
            try:
                return int(d.get(k) or 0)
            except ValueError:          # ncu prints "-" where a line has no such counter
                return 0
        recs.append((cur, int(r[0]), r[1].strip()[:90], num("# Samples"), num("Instructions Executed"),
                     num("Thread Instructions Executed"), num("L1 Wavefronts Shared"),
                     num("stall_no_inst"), num("stall_wait"), num("stall_short_sb"), num("stall_long_sb")))
    ts = sum(x[3] for x in recs) or 1
    ti = sum(x[4] for x in recs) or 1
    print("# ncu source page, hottest CUDA source lines of %s" % path)
    print("# total samples %d, executed warp instructions %d" % (ts, ti))
    byfile = defaultdict(lambda: [0, 0])
    for x in recs:
        byfile[x[0]][0] += x[3]
        byfile[x[0]][1] += x[4]
    for f, v in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
        print("  file %-28s samples %5.1f%%  warp-inst %5.1f%%" % (f, 100 * v[0] / ts, 100 * v[1] / ti))
    print("%-22s %6s %6s %5s %9s %6s %6s %6s %6s  %s" % ("file:line", "samp%", "inst%", "thr/w", "smem_wf", "noinst", "wait", "shortsb", "longsb", "source"))
    for x in sorted(recs, key=lambda x: -x[3])[:top]:
        eff = x[5] / x[4] if x[4] else 0
        print("%-22s %6.2f %6.2f %5.1f %9d %6d %6d %6d %6d  %s" % ("%s:%d" % (x[0], x[1]), 100 * x[3] / ts, 100 * x[4] / ti, eff, x[6], x[7], x[8], x[9], x[10], x[2]))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "source":
    source(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 45)


def metrics(paths, out_json):
    """profiles/ncu_metrics.json: per kernel, the few numbers bench.py quotes next to its live timings."""
    import json
    import os
    import re
    res = {}
    if os.path.exists(out_json):
        res = json.load(open(out_json))
    for path in paths:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = [r for r in csv.reader(out.splitlines()) if r]
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            d = dict(zip(hdr, zip(units, vals)))
            name = d["Kernel Name"][1]
            m = re.match(r"(?:void )?(?:sccav::)?(\w+)<(double|float)", name)
            key = "%s<%s>" % (m.group(1), m.group(2)) if m else name

            def val(k, scale={"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}):
                u, v = d[k]
                return float(v.replace(",", "")) * scale.get(u, 1.0)
            res[key] = {
                "source": path, "duration_ms": val("gpu__time_duration.sum"),
                "dram_bytes": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                "fp64_pipe_active_pct": val("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "tensor_pipe_active_pct": val("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                "dram_throughput_pct": val("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                "registers_per_thread": val("launch__registers_per_thread"),
                "grid": d["Grid Size"][1], "block": d["Block Size"][1],
            }
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    res["_csrc_sha"] = bench.csrc_sha()            # bench.py refuses to quote these numbers once csrc/ has changed
    json.dump(res, open(out_json, "w"), indent=1, sort_keys=True)
    print(json.dumps(res, indent=1, sort_keys=True))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "metrics":
    metrics(sys.argv[3:], sys.argv[2])
