#!/bin/bash
# round 2: K10 with thresholds from the group's minima -- suite, sanitizer on the K10 tests, A/B, bench
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu";  timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -5
echo "== A/B group_min (filter kernel ms)"
for c in "1000000 128 1024" "1000000 128 64" "4000000 32 1024" "1000000 768 1024" "10000000 128 1024"; do
  for gm in 1 0; do python scripts/k10_role_cycles.py $c $gm 2>&1 | tail -1 | cut -c1-700; done; done | tee gpurun_out/r02_K10_role_cycles_ab.jsonl | cut -c1-200
export SVDB_ARENA=malloc
CS=/usr/local/cuda/bin/compute-sanitizer
echo "== memcheck K10"; timeout 1800 $CS --tool memcheck --error-exitcode 9 python -m pytest -q -x --timeout=3000 -p no:cacheprovider -m gpu tests/test_gpu_umma.py > gpurun_out/san_mem_umma.log 2>&1; echo "exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san_mem_umma.log | tail -2
echo "== racecheck K10"; timeout 2400 $CS --tool racecheck --error-exitcode 9 python -m pytest -q -x --timeout=3000 -p no:cacheprovider -m gpu tests/test_gpu_umma.py -k "not accumulator" > gpurun_out/san_race_umma.log 2>&1; echo "exit $?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/san_race_umma.log | tail -2
unset SVDB_ARENA
echo "== bench (ours)";   timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; tail -2 gpurun_out/r02_bench_n1_final.err; cut -c1-300 gpurun_out/r02_bench_n1_final.json
echo "== other configs"; rm -f gpurun_out/r02_extra_final.jsonl; timeout 1500 python scripts/bench_extra.py c2 c5 --out=gpurun_out/r02_extra_final2.jsonl 2>&1 | cut -c1-260 | grep "queries_per_call\|parity" | head
