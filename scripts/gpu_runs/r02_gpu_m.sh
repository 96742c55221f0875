#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu";  timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -5
echo "== smoke";          timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench (ours)";   timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; tail -2 gpurun_out/r02_bench_n1_final.err; cut -c1-300 gpurun_out/r02_bench_n1_final.json
echo "== bench (reference arm)"; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_final.json 2>/dev/null; cut -c1-200 gpurun_out/r02_bench_reference_final.json
echo "== ncu: launch list of bench.py's timed region"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r02_launches_bench_10Mx768_final.csv \
    python bench.py --steps 20 --warmup 3 --batch-queries 0 --no-cpu-baseline --no-fp64-scan --no-parity-check > gpurun_out/ncu_launch_bench.log 2>&1; echo "exit $?"
