#!/bin/bash
# ncu --set full capture of the rollout kernel on prepared rows (SPEC 2) through scripts/microbench.py
set -u
TAG=${1:-p3}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 8 -c 1 \
    -f -o "$OUT/prof_rollout_prep" python scripts/microbench.py --no-operator > "$OUT/ncu_rollout_prep.log" 2>&1
echo "ncu rollout prep exit $?"
tail -3 "$OUT/ncu_rollout_prep.log"
