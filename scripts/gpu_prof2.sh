#!/bin/bash
# ncu --set full capture of the prepared-operator kernel (SPEC 2) through scripts/microbench.py
set -u
TAG=${1:-p2}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:filter_step_kernel -s 24 -c 1 \
    -f -o "$OUT/prof_filter_prep" python scripts/microbench.py --no-rollout > "$OUT/ncu_filter_prep.log" 2>&1
echo "ncu filter prep exit $?"
tail -3 "$OUT/ncu_filter_prep.log"
