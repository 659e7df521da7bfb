#!/bin/bash
set -u
OUT=gpurun_out/${1:-s4k}; mkdir -p "$OUT"
timeout 600 python -m pytest tests -m gpu -q -x > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?"; tail -15 "$OUT/pytest_gpu.log"
