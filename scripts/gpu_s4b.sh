#!/bin/bash
# new-kernel visit: parity tests of the QP shortcut / pipelined K12, A/B microbench, ncu capture of the pipelined operator
set -u
TAG=${1:-s4b}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
timeout 600 python -m pytest tests -m gpu -q -x > "$OUT/pytest_gpu.log" 2>&1
echo "pytest exit $?"; tail -15 "$OUT/pytest_gpu.log"
echo "== microbench (pipe on)"
timeout 300 python scripts/microbench.py > "$OUT/micro_pipe1.log" 2>&1; cat "$OUT/micro_pipe1.log"
echo "== microbench operator (pipe off)"
SCCAV_K12_PIPE=0 timeout 300 python scripts/microbench.py --no-rollout > "$OUT/micro_pipe0.log" 2>&1; cat "$OUT/micro_pipe0.log"
if [ "${NCU:-1}" = "1" ]; then
# microbench launches the operator 10x canonical, 10x prepared, 10x prepared static: capture the last of each
for spec in ${NCU_SPECS:-9:canon 19:prep 29:prepstatic}; do
  skip=${spec%%:*}; name=${spec##*:}
  SCCAV_K12_PIPE=${NCU_PIPE:-1} timeout 300 ncu --set full --clock-control none --import-source on -k regex:filter_step -s $skip -c 1 \
      -f -o "$OUT/prof_filter_$name" python scripts/microbench.py --no-rollout > "$OUT/ncu_filter_$name.log" 2>&1
  echo "ncu $name exit $?"
done
fi
