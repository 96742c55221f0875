#!/bin/bash
# Parity suite (all tests), full-size bench, ncu launch list + full capture of the scan kernel.
set -u
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -60 gpurun_out/pytest_gpu.log
echo "== bench full"
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv 2>&1 &
SMI=$!
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench full exit $?"
kill $SMI
tail -3 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
echo "== bench variant 1 (LDG)"
timeout 600 python bench.py --opt scan.variant=1 --no-cpu-baseline --batch-queries 0 > gpurun_out/bench_ldg.json 2> gpurun_out/bench_ldg.err; echo "exit $?"
cat gpurun_out/bench_ldg.json
echo "== ncu launch list (2M rows to keep it short)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --rows 2000000 --steps 3 --warmup 3 --batch-queries 16 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1; echo "ncu list exit $?"
echo "== ncu full capture of the scan kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_wide_kernel -s 3 -c 2 -o gpurun_out/prof_scan python bench.py --rows 2000000 --steps 3 --warmup 3 --batch-queries 0 --no-cpu-baseline > gpurun_out/ncu_full_bench.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out
