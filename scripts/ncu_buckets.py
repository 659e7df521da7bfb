#!/usr/bin/env python
"""Executed-instruction / sample shares of a rollout-kernel ncu capture by code region (needs -lineinfo)."""
import collections
import csv
import subprocess
import sys


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    cur = hdr = None
    recs = []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or cur is None or not r[0].isdigit():
            continue
        d = dict(zip(hdr, r))

        def num(k):
            try:
                return int(d.get(k) or 0)
            except ValueError:
                return 0
        recs.append((cur, r[1], num("# Samples"), num("Instructions Executed"), num("Thread Instructions Executed")))
    ti = sum(x[3] for x in recs) or 1
    ts = sum(x[2] for x in recs) or 1
    b = collections.OrderedDict()
    for f, src, s, i, th in recs:
        if f == "course_index.cuh":
            k = "nearest way-point search"
        elif f == "kernels.cuh":
            k = "kernel body (plant, bookkeeping, staging)"
        elif f == "path.cuh":
            if "::sincos" in src or "::tan(" in src or "::atan2" in src or "::atan(" in src or "::sqrt" in src:
                k = "libm (sincos / tan / atan2)"
            elif "rv.A0" in src or "rv.A1" in src or "rv.b(" in src or "qp_check" in src or "rows[(pitch" in src:
                k = "QP (row reads)"
            else:
                k = "path.cuh other (rows, QP arithmetic, Stanley law)"
        else:
            k = f
        v = b.setdefault(k, [0, 0, 0])
        v[0] += s; v[1] += i; v[2] += th
    print("# %s: %d samples, %d executed warp instructions" % (path, ts, ti))
    for k, (s, i, th) in b.items():
        print("%-52s samples %5.1f%%  warp-inst %5.1f%%  threads/inst %4.1f" % (k, 100 * s / ts, 100 * i / ti, th / max(i, 1)))


if __name__ == "__main__":
    main(sys.argv[1])
