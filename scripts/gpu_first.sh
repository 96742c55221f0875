#!/bin/bash
# First GPU contact: parity tests, smoke, a small and then the full bench. Logs -> gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== pytest gpu" 
timeout 1200 python -m pytest tests -m gpu -q --timeout=600 -x -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
echo "== bench small"
timeout 600 python bench.py --rows 1000000 --steps 5 --warmup 3 --batch-queries 64 --cpu-sample-rows 20000 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "bench small exit $?"
tail -3 gpurun_out/bench_small.err; cat gpurun_out/bench_small.json
