#!/bin/bash
set -u
OUT=gpurun_out/${1:-s4i}; mkdir -p "$OUT"
timeout 600 python -m pytest tests -m gpu -q -x > "$OUT/pytest_gpu.log" 2>&1; echo "pytest exit $?"; tail -4 "$OUT/pytest_gpu.log"
echo "== defaults"; timeout 300 python scripts/microbench.py --no-rollout 2>&1 | tee "$OUT/micro_default.log"
echo "== direct, coop"; SCCAV_K12_PIPE=0 SCCAV_K12_QP=coop timeout 300 python scripts/microbench.py --no-rollout 2>&1 | tee "$OUT/micro_direct_coop.log"
echo "== direct, thread"; SCCAV_K12_PIPE=0 SCCAV_K12_QP=thread timeout 300 python scripts/microbench.py --no-rollout 2>&1 | tee "$OUT/micro_direct_thread.log"
