#!/bin/bash
# round 2, late: parity suite + smoke + both bench arms on the final routing (k-dependent plane choice, drop-in changes)
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu";  timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -6
echo "== smoke";          timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench (ours)";   timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; tail -2 gpurun_out/r02_bench_n1_final.err; cut -c1-300 gpurun_out/r02_bench_n1_final.json
echo "== bench (reference arm)"; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_final.json 2>/dev/null; cut -c1-200 gpurun_out/r02_bench_reference_final.json
