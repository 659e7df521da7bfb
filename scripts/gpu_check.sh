#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), ncu launch list, one --set full capture.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_check.sh [tag]
# Everything is written under gpurun_out/<tag>/ ; copy what should be judged into profiles/.
set -u
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > "$OUT/gpu.csv" 2>&1

echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q > "$OUT/pytest_gpu.log" 2>&1
echo "pytest exit $?"; tail -5 "$OUT/pytest_gpu.log"

echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1
echo "smoke exit $?"; tail -2 "$OUT/smoke.log"

echo "== bench (ours)"
timeout 900 python bench.py --steps 5 --warmup 3 > "$OUT/bench.json" 2> "$OUT/bench.err"
echo "bench exit $?"; tail -c 3000 "$OUT/bench.json"; tail -3 "$OUT/bench.err"

echo "== bench (reference arm)"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"
echo "ref exit $?"; tail -c 1500 "$OUT/bench_ref.json"

echo "== ncu launch list (same command, fewer steps)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file "$OUT/launches.csv" python bench.py --steps 2 --warmup 3 --no-cpu > "$OUT/bench_under_ncu.log" 2>&1
echo "ncu list exit $?"

echo "== ncu --set full: rollout kernel + fused operator kernel"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 3 -c 1 \
    -f -o "$OUT/prof_rollout" python bench.py --steps 1 --warmup 3 --no-cpu --no-operator > "$OUT/ncu_rollout.log" 2>&1
echo "ncu rollout exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:filter_step_kernel -s 3 -c 1 \
    -f -o "$OUT/prof_filter" python bench.py --steps 1 --warmup 3 --no-cpu --T 10 > "$OUT/ncu_filter.log" 2>&1
echo "ncu filter exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:filter_step_staged -s 3 -c 1 \
    -f -o "$OUT/prof_filter_staged" python bench.py --steps 1 --warmup 3 --no-cpu --T 10 > "$OUT/ncu_filter_staged.log" 2>&1
echo "ncu filter staged exit $?"
ls -la "$OUT"
