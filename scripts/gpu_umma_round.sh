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
if [ "${UMMA_FULL:-0}" = "1" ]; then
echo "== [$((SECONDS-T0)) s] pytest -m gpu (all)"; timeout 900 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider 2>&1 | tail -15 | cut -c1-300
echo "== [$((SECONDS-T0)) s] smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== [$((SECONDS-T0)) s] bench (ours)"; timeout 600 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; cat gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
echo "== [$((SECONDS-T0)) s] ncu: K10 at 1M x 768, 1024 queries"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:umma_filter_kernel -s 1 -c 1 -o gpurun_out/prof_umma \
    python bench.py --rows 1000000 --steps 3 --warmup 3 --batch-queries 1024 --no-cpu-baseline > gpurun_out/ncu_umma.log 2>&1; echo "exit $?"
echo "== [$((SECONDS-T0)) s] done"; ls -la gpurun_out | head -40
fi
