#!/bin/bash
# One GPU box, one call: parity suite, smoke, both bench arms, the other configs, and the ncu captures
# that profiles/ keeps.   gpurun --timeout 2400 -- 'bash scripts/gpu_round.sh'
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu";  timeout 1200 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -6
echo "== smoke";          timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (ours)";   timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
echo "== bench (reference arm)"; timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cut -c1-300 gpurun_out/bench_reference.json
echo "== other configs";  rm -f gpurun_out/extra.jsonl; timeout 900 python scripts/bench_extra.py c1 c2 c4 lat > gpurun_out/extra.log 2>&1; tail -30 gpurun_out/extra.log | cut -c1-260
echo "== host call latency"; timeout 300 python scripts/host_call_latency.py 2>&1 | tail -20
echo "== ncu: launch list of bench.py (the timed region: bench.py brackets it with cudaProfilerStart/Stop)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/launches_full.csv \
    python bench.py --steps 20 --warmup 3 --batch-queries 0 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1; echo "exit $?"
echo "== ncu: full capture of the scan kernel (one launch at full size)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:scan_wide_kernel -s 4 -c 1 -o gpurun_out/prof_scan_full \
    python bench.py --steps 3 --warmup 3 --batch-queries 0 --no-cpu-baseline > gpurun_out/ncu_full_bench.log 2>&1; echo "exit $?"
echo "== ncu: K2, K10, K11 (bench.py stops the profiler after its timed region, so these come from the bring-up scripts)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_mma_kernel -c 1 -o gpurun_out/prof_mma -f \
    python scripts/debug_umma.py time_1M > gpurun_out/ncu_mma.log 2>&1; echo "exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:umma_filter_kernel -c 1 -o gpurun_out/prof_umma -f \
    python scripts/debug_umma.py time_1M > gpurun_out/ncu_umma.log 2>&1; echo "exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:scan_shadow_kernel -s 2 -c 1 -o gpurun_out/prof_shadow -f \
    python scripts/debug_shadow.py 2000000 768 > gpurun_out/ncu_shadow.log 2>&1; echo "exit $?"
echo "== which kernel for which batch size"; timeout 600 python scripts/sweep_batch_paths.py > gpurun_out/sweep_batch_paths.jsonl 2> gpurun_out/sweep_batch_paths.err; cut -c1-400 gpurun_out/sweep_batch_paths.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:compare_kernel -s 12 -c 1 -o gpurun_out/prof_compare \
    python scripts/bench_extra.py c4 --out=gpurun_out/tmp.jsonl > gpurun_out/ncu_compare.log 2>&1; echo "exit $?"
ls -la gpurun_out
