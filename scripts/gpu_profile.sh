#!/bin/bash
# Profile visit: ncu launch list of the bench command + --set full captures of the rollout / filter kernels.
set -u
TAG=${1:-p}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
echo "== ncu launch list (same command as the bench, fewer steps)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file "$OUT/launches.csv" python bench.py --steps 2 --warmup 3 --no-cpu --no-configs > "$OUT/bench_under_ncu.log" 2>&1
echo "ncu list exit $?"
echo "== ncu --set full: rollout kernel (the timed instance)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 3 -c 1 \
    -f -o "$OUT/prof_rollout" python bench.py --steps 1 --warmup 3 --no-cpu --no-operator --no-configs > "$OUT/ncu_rollout.log" 2>&1
echo "ncu rollout exit $?"
echo "== ncu --set full: rollout kernel, canonical rows"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 3 -c 1 \
    -f -o "$OUT/prof_rollout_canonical" python bench.py --rows canonical --steps 1 --warmup 3 --no-cpu --no-operator --no-configs > "$OUT/ncu_rollout_c.log" 2>&1
echo "ncu rollout canonical exit $?"
echo "== ncu --set full: filter kernels"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:filter_step_kernel -s 3 -c 1 \
    -f -o "$OUT/prof_filter" python bench.py --steps 1 --warmup 3 --no-cpu --no-configs --T 10 > "$OUT/ncu_filter.log" 2>&1
echo "ncu filter exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:filter_step_staged -s 3 -c 1 \
    -f -o "$OUT/prof_filter_staged" python bench.py --steps 1 --warmup 3 --no-cpu --no-configs --T 10 > "$OUT/ncu_filter_staged.log" 2>&1
echo "ncu filter staged exit $?"
ls -la "$OUT"
