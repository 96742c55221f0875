#!/bin/bash
# round 2: parity suite, both bench arms, launch list + ncu capture of the K12 scan, batch-path sweep
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu";  timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -15
echo "== smoke";          timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench (ours)";   timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -3 gpurun_out/r02_bench_n1.err; cut -c1-700 gpurun_out/r02_bench_n1.json
echo "== bench (reference arm)"; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; cut -c1-600 gpurun_out/r02_bench_reference.json
echo "== ncu: launch list of bench.py's timed region"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r02_launches_bench_10Mx768.csv \
    python bench.py --steps 20 --warmup 3 --batch-queries 0 --no-cpu-baseline --no-fp64-scan --no-parity-check > gpurun_out/ncu_launch_bench.log 2>&1; echo "exit $?"
echo "== ncu: full capture of the K12 scan (one launch at full size)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:scan_plane_kernel -s 4 -c 1 -o gpurun_out/r02_prof_scan_plane -f \
    python bench.py --steps 3 --warmup 3 --batch-queries 0 --no-cpu-baseline --no-fp64-scan --no-parity-check > gpurun_out/ncu_full_bench.log 2>&1; echo "exit $?"
echo "== which kernel for which batch size"
for shape in "2000000 768" "1000000 128" "4000000 32"; do
  timeout 600 python scripts/sweep_batch_paths.py $shape >> gpurun_out/r02_sweep_batch_paths.jsonl 2>> gpurun_out/sweep_batch_paths.err
done
cut -c1-420 gpurun_out/r02_sweep_batch_paths.jsonl
ls -la gpurun_out | tail -12
