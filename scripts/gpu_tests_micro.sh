#!/bin/bash
# GPU parity tests, then the rollout microbench (developer visit; gpu_check.sh is the full one)
set -u
OUT=gpurun_out/${1:-tm}; mkdir -p "$OUT"
timeout 600 python -m pytest tests -m gpu -q -x > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?"; tail -4 "$OUT/pytest_gpu.log"
timeout 300 python scripts/microbench.py --no-operator 2>&1 | tee "$OUT/micro.log"
