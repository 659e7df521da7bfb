#!/bin/bash
# GPU visit: parity tests, smoke, the full default bench line (wall-clocked), optional ncu capture.
set -u
TAG=${1:-b}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
if [ "${TESTS:-1}" = "1" ]; then
echo "== pytest -m gpu"
timeout 600 python -m pytest tests -m gpu -q -x > "$OUT/pytest_gpu.log" 2>&1
echo "pytest exit $?"; tail -4 "$OUT/pytest_gpu.log"
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1
echo "smoke exit $?"; tail -2 "$OUT/smoke.log"
fi
echo "== bench (ours)"
START=$(date +%s)
timeout 1200 python bench.py ${BENCH_ARGS:-} > "$OUT/bench.json" 2> "$OUT/bench.err"
echo "bench exit $?  wall $(( $(date +%s) - START )) s"
python - "$OUT/bench.json" <<'P'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value %.4g  ms/step %.3f  e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
    for k,v in d.items():
        if not isinstance(v,(dict,list)) and k not in ("metric","unit","data"): print("  ", k, v)
    print("roofline", {k: d["roofline"][k] for k in ("achieved","peak","frac")})
    print("clocks", d["clocks"])
except Exception as e:
    print("bench parse failed", e)
P
tail -5 "$OUT/bench.err"
if [ "${NCU:-0}" = "1" ]; then
echo "== ncu --set full: rollout kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 3 -c 1 \
    -f -o "$OUT/prof_rollout" python bench.py --steps 1 --warmup 3 --no-cpu --no-operator --no-configs > "$OUT/ncu_rollout.log" 2>&1
echo "ncu rollout exit $?"
fi
