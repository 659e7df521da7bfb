#!/bin/bash
# run selected GPU tests one by one under their own timeout (a hang costs one test, not the visit)
OUT=gpurun_out/${1:-each}; mkdir -p $OUT; shift
for t in "$@"; do
  echo "=== $t"
  timeout 150 python -X faulthandler -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$t" > $OUT/$t.log 2>&1
  echo "exit $?"; grep -E "^E  |passed|failed|Error" $OUT/$t.log | head -12
done
