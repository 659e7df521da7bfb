#!/bin/bash
# A/B of the staged filter-step kernel's QP form (SCCAV_K12_QP=thread|coop), operator microbench only
set -u
OUT=gpurun_out/${1:-s4h}; mkdir -p "$OUT"
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "filter or qp or k1_k2 or prepared" > "$OUT/pytest.log" 2>&1; tail -3 "$OUT/pytest.log"
for qp in thread coop; do
  echo "== staged, QP=$qp"
  SCCAV_K12_QP=$qp timeout 300 python scripts/microbench.py --no-rollout 2>&1 | tee "$OUT/micro_$qp.log"
  SCCAV_K12_QP=$qp timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "filter or qp or k1_k2 or prepared" 2>&1 | tail -1
done
