#!/bin/bash
set -u
OUT=gpurun_out/${1:-s4q}; mkdir -p "$OUT"
timeout 600 python -m pytest tests -m gpu -q -x > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?"; tail -4 "$OUT/pytest_gpu.log"
timeout 300 python scripts/microbench.py --no-operator 2>&1 | tee "$OUT/micro.log"
