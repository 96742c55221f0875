#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"
timeout 900 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -6
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench full (with cpu baseline)"
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench exit $?"; cat gpurun_out/bench_full.json
echo "== reference arm"
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref exit $?"; cat gpurun_out/bench_reference.json | cut -c1-900
echo "== extra c2 c4"
timeout 600 python scripts/bench_extra.py c2 c4 lat --out=gpurun_out/extra_v2.jsonl 2>&1 | cut -c1-300
echo "== ncu compare kernel"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:compare_kernel -s 12 -c 1 -o gpurun_out/prof_compare python scripts/bench_extra.py c4 --out=gpurun_out/tmp.jsonl > gpurun_out/ncu_compare.log 2>&1; echo ncu exit $?; tail -3 gpurun_out/ncu_compare.log
