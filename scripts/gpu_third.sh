#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
echo "== bench full"
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench full exit $?"
tail -3 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
echo "== extra configs"
rm -f gpurun_out/extra.jsonl
timeout 900 python scripts/bench_extra.py c1 c2 c4 > gpurun_out/extra.log 2>&1; echo "extra exit $?"
tail -40 gpurun_out/extra.log
echo "== ncu launch list, full size"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_full.csv python bench.py --steps 3 --warmup 3 --batch-queries 16 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1; echo "ncu list exit $?"
echo "== ncu full capture of the scan kernel at full size"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:scan_wide_kernel -s 4 -c 1 -o gpurun_out/prof_scan_full python bench.py --steps 3 --warmup 3 --batch-queries 0 --no-cpu-baseline > gpurun_out/ncu_full_bench.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out
