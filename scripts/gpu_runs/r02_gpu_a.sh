#!/bin/bash
# round 2, first GPU call: parity suite on the new scan path (K12 + fused tail), smoke, one bench line
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu";  timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 -p no:cacheprovider 2>&1 | tail -15
echo "== smoke";          timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench (ours)";   timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; tail -5 gpurun_out/r02_bench_a.err; cut -c1-3000 gpurun_out/r02_bench_a.json
