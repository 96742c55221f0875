#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -60 gpurun_out/pytest_gpu.log
echo "== bench full"
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench full exit $?"
tail -3 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
echo "== ncu of the mma kernel (small)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_mma_kernel -c 1 -o gpurun_out/prof_mma python bench.py --rows 1000000 --steps 3 --warmup 3 --batch-queries 1024 --no-cpu-baseline > gpurun_out/ncu_mma.log 2>&1; echo "ncu exit $?"
