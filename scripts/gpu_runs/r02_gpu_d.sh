#!/bin/bash
# round 2: parity suite with K13 as the default single-query scan, both bench arms, launch list + ncu capture of K13
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu";  timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -6
echo "== smoke";          timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench (ours)";   timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; tail -2 gpurun_out/r02_bench_n1_final.err; cut -c1-300 gpurun_out/r02_bench_n1_final.json
echo "== bench (reference arm)"; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_final.json 2>/dev/null; cut -c1-200 gpurun_out/r02_bench_reference_final.json
echo "== ncu: launch list of bench.py's timed region"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r02_launches_bench_10Mx768_K13.csv \
    python bench.py --steps 20 --warmup 3 --batch-queries 0 --no-cpu-baseline --no-fp64-scan --no-parity-check > gpurun_out/ncu_launch_bench.log 2>&1; echo "exit $?"
echo "== ncu: full capture of the K13 scan (one launch at full size)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:scan_plane8_kernel -s 4 -c 1 -o gpurun_out/r02_prof_scan_plane8 -f \
    python bench.py --steps 3 --warmup 3 --batch-queries 0 --no-cpu-baseline --no-fp64-scan --no-parity-check > gpurun_out/ncu_full_bench.log 2>&1; echo "exit $?"
echo "== other configs (2, 4, 5 on one GPU) with their parity checks"
rm -f gpurun_out/r02_extra_final.jsonl; timeout 1500 python scripts/bench_extra.py c2 c4 c5 lat --out=gpurun_out/r02_extra_final.jsonl 2>&1 | cut -c1-330 | tail -24
