#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -60 gpurun_out/pytest_gpu.log
echo "== extra c1"
rm -f gpurun_out/extra.jsonl
timeout 600 python scripts/bench_extra.py c1 > gpurun_out/extra_c1.log 2>&1; echo "extra exit $?"
tail -8 gpurun_out/extra_c1.log
echo "== bench full"
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench full exit $?"
tail -3 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
