#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== plane scan tests"; timeout 900 python -m pytest tests/test_gpu_shadow_scan.py -m gpu -q -x --timeout=600 -p no:cacheprovider 2>&1 | tail -15
echo "== pytest -m gpu";  timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -5
echo "== bench (overlap on)";   timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_overlap.json 2> gpurun_out/b1.err; tail -2 gpurun_out/b1.err; cut -c1-300 gpurun_out/r02_bench_n1_overlap.json
echo "== bench (overlap off)";  timeout 900 python bench.py --steps 20 --warmup 5 --batch-queries 0 --no-fp64-scan --no-parity-check --no-cpu-baseline --opt scan.overlap_steps=0 > gpurun_out/r02_bench_n1_no_overlap.json 2> gpurun_out/b2.err; tail -2 gpurun_out/b2.err; cut -c1-300 gpurun_out/r02_bench_n1_no_overlap.json
echo "== small shard: 1.25M x 768, overlap on / off (200 steps)"
timeout 600 python bench.py --rows 1250000 --steps 200 --warmup 10 --batch-queries 0 --no-fp64-scan --no-parity-check --no-cpu-baseline > gpurun_out/r02_bench_1250k_overlap.json 2>/dev/null; cut -c1-260 gpurun_out/r02_bench_1250k_overlap.json
timeout 600 python bench.py --rows 1250000 --steps 200 --warmup 10 --batch-queries 0 --no-fp64-scan --no-parity-check --no-cpu-baseline --opt scan.overlap_steps=0 > gpurun_out/r02_bench_1250k_no_overlap.json 2>/dev/null; cut -c1-260 gpurun_out/r02_bench_1250k_no_overlap.json
