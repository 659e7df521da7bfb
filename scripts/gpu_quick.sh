#!/bin/bash
# Quick GPU visit: parity tests, bench (ours), one ncu --set full capture of the rollout kernel.
set -u
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
echo "== pytest -m gpu"
timeout 400 python -m pytest tests -m gpu -q -x > "$OUT/pytest_gpu.log" 2>&1
echo "pytest exit $?"; tail -4 "$OUT/pytest_gpu.log"
echo "== bench (ours)"
timeout 600 python bench.py --steps 5 --warmup 3 ${BENCH_ARGS:-} > "$OUT/bench.json" 2> "$OUT/bench.err"
echo "bench exit $?"; python - "$OUT/bench.json" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value %.4g %s  ms/step %.3f  e2e %.4g  roofline %s  op %s  cpu %s" % (d["value"], d["unit"], d["ms_per_step"], d["e2e"]["value"],
          {k: d["roofline"][k] for k in ("achieved","peak","frac")}, d["roofline_operator"] and {k: d["roofline_operator"][k] for k in ("achieved","frac")}, d["cpu_baseline"] and d["cpu_baseline"]["value"]))
    print("clocks", d["clocks"])
except Exception as e:
    print("bench parse failed", e)
P
tail -3 "$OUT/bench.err"
if [ "${NCU:-1}" = "1" ]; then
echo "== ncu --set full: rollout kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 3 -c 1 \
    -f -o "$OUT/prof_rollout" python bench.py --steps 1 --warmup 3 --no-cpu --no-operator > "$OUT/ncu_rollout.log" 2>&1
echo "ncu rollout exit $?"
fi
