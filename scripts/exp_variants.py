#!/usr/bin/env python
"""Developer tool: build several variants of libsccav_cbf.so (extra -D flags) under build/variants/
and time the two headline kernels with each on the GPU box.

    python scripts/exp_variants.py build  name1:-DFOO name2:-DBAR=1,-DBAZ ...   (here; nvcc cross-compiles)
    python scripts/exp_variants.py run                                           (under gpurun)

`run` executes scripts/microbench.py once per variant in a fresh process with SCCAV_CBF_LIB set.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "build", "variants")


def build(specs):
    from sccav_cbf_b200 import build as b
    os.makedirs(VDIR, exist_ok=True)
    procs = []
    for spec in specs:
        name, _, flags = spec.partition(":")
        flags = [f for f in flags.split(",") if f]
        d = os.path.join(VDIR, name)
        os.makedirs(d, exist_ok=True)
        objs = []
        for src, extra in b.UNITS:
            o = os.path.join(d, src.replace(".cu", ".o"))
            objs.append(o)
            cmd = [b._nvcc()] + b.ARCH + [c for c in b.COMMON if not c.startswith("--use_fast_math")] + extra + flags + \
                  ["-c", os.path.join(b.CSRC, src), "-o", o]
            procs.append((name, subprocess.Popen(cmd)))
    for name, p in procs:
        if p.wait() != 0:
            raise SystemExit("variant %s failed to compile" % name)
    for spec in specs:
        name = spec.partition(":")[0]
        d = os.path.join(VDIR, name)
        objs = [os.path.join(d, src.replace(".cu", ".o")) for src, _ in b.UNITS]
        lib = os.path.join(VDIR, "lib_%s.so" % name)
        subprocess.run([b._nvcc()] + b.ARCH + ["-shared", "-cudart", "static", "-o", lib] + objs, check=True)
        for o in objs:
            os.remove(o)
        print("built", lib)


def run(extra):
    libs = sorted(f for f in os.listdir(VDIR) if f.startswith("lib_") and f.endswith(".so"))
    for f in libs:
        env = dict(os.environ, SCCAV_CBF_LIB=os.path.join(VDIR, f))
        print("=== %s" % f[4:-3], flush=True)
        subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "microbench.py")] + extra, env=env)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build(sys.argv[2:])
    else:
        run(sys.argv[2:])
