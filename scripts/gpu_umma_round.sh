#!/bin/bash
# K10 bring-up on one GPU box.   gpurun --timeout 600 -- 'bash scripts/gpu_umma_round.sh [cases...]'
set -u
mkdir -p gpurun_out
T0=$SECONDS
CASES="${@:-tiny one_tile k768 ragged kd_prefix}"
: > gpurun_out/umma_debug.jsonl
for c in $CASES; do
  echo "== [$((SECONDS-T0)) s] case $c"
  timeout 150 python scripts/debug_umma.py $c > gpurun_out/umma_case.out 2> gpurun_out/umma_case.err; rc=$?
  echo "exit $rc"; cat gpurun_out/umma_case.out | cut -c1-1500; cat gpurun_out/umma_case.out >> gpurun_out/umma_debug.jsonl
  if [ $rc -ne 0 ]; then tail -8 gpurun_out/umma_case.err | cut -c1-600; fi
done
echo "== [$((SECONDS-T0)) s] pytest K10"
timeout 300 python -m pytest tests/test_gpu_umma.py -m gpu -q -x --timeout=200 -p no:cacheprovider 2>&1 | tail -25 | cut -c1-400
echo "== [$((SECONDS-T0)) s] done"
