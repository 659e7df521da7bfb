#!/bin/bash
# ncu --set full captures of the two headline kernels through scripts/microbench.py (one launch each).
# Usage (under gpurun): bash scripts/gpu_prof.sh <tag> [rollout|operator|both]
set -u
TAG=${1:-p}; WHAT=${2:-both}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
if [ "$WHAT" != "operator" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 2 -c 1 \
    -f -o "$OUT/prof_rollout" python scripts/microbench.py --no-operator > "$OUT/ncu_rollout.log" 2>&1
echo "ncu rollout exit $?"
fi
if [ "$WHAT" != "rollout" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:filter_step_kernel -s 3 -c 1 \
    -f -o "$OUT/prof_filter" python scripts/microbench.py --no-rollout > "$OUT/ncu_filter.log" 2>&1
echo "ncu filter exit $?"
fi
