#!/bin/bash
# round 2, last 1-GPU run: suite, smoke, sanitizer over the K10 tests, bench arms, K10 A/B on the final code
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu";  timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -4
echo "== smoke";          timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
export SVDB_ARENA=malloc
CS=/usr/local/cuda/bin/compute-sanitizer
echo "== memcheck K10"; timeout 1800 $CS --tool memcheck --error-exitcode 9 python -m pytest -q -x --timeout=3000 -p no:cacheprovider -m gpu tests/test_gpu_umma.py -k "not decreasing" > gpurun_out/san_mem_umma.log 2>&1; echo "exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san_mem_umma.log | tail -2
echo "== racecheck K10"; timeout 2400 $CS --tool racecheck --error-exitcode 9 python -m pytest -q -x --timeout=3000 -p no:cacheprovider -m gpu tests/test_gpu_umma.py -k "not accumulator and not decreasing" > gpurun_out/san_race_umma.log 2>&1; echo "exit $?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/san_race_umma.log | tail -2
unset SVDB_ARENA
echo "== bench (ours)";   timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/b1.err; tail -1 gpurun_out/b1.err; cut -c1-260 gpurun_out/r02_bench_n1_final.json
echo "== bench (reference arm)"; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_final.json 2>/dev/null; cut -c1-160 gpurun_out/r02_bench_reference_final.json
echo "== other configs"; timeout 1500 python scripts/bench_extra.py c2 c5 --out=gpurun_out/r02_extra_c2_c5_n1_final.jsonl 2>&1 | cut -c1-230 | grep "queries_per_call\|parity" | head
