"""Build recipe for libsccav_cbf.so (hand-written sm_100a CUDA + the C-ABI), in-tree.

    python -m sccav_cbf_b200.build          # or __graft_entry__.build()

nvcc cross-compiles without a GPU.  The fp64 translation unit is compiled with -fmad=false so
that its arithmetic rounds exactly like the oracle's (index / active-set parity); the fp32 unit
keeps FMA contraction.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsccav_cbf.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--use_fast_math=false"]
UNITS = [
    ("capi_common.cu", []),
    ("capi_f64.cu", ["-fmad=false"]),
    ("capi_f32.cu", ["-fmad=true"]),
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build libsccav_cbf.so")
    return exe


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "sccav_cbf.h"))
    headers.append(os.path.abspath(__file__))
    objs = []
    for src, extra in UNITS:
        s = os.path.join(CSRC, src)
        o = os.path.join(CSRC, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + ARCH + [c for c in COMMON if not c.startswith("--use_fast_math")] + extra + os.environ.get("SCCAV_NVCC_EXTRA", "").split() + ["-c", s, "-o", o]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd), flush=True)
            subprocess.run(cmd, check=True)
    if force or _stale(LIB, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-cudart", "static", "-o", LIB] + objs
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv or "--verbose" in sys.argv))
