#!/bin/bash
# round 2: final routing (K13 up to k = 16 and on short rows, K12 up to k = 24) -- suite, bench, other configs
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu";  timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -6
echo "== smoke";          timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench (ours)";   timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; tail -2 gpurun_out/r02_bench_n1_final.err; cut -c1-300 gpurun_out/r02_bench_n1_final.json
echo "== other configs (2, 4, 5 on one GPU) with their parity checks"
rm -f gpurun_out/r02_extra_final.jsonl; timeout 1500 python scripts/bench_extra.py c2 c4 c5 lat --out=gpurun_out/r02_extra_final.jsonl 2>&1 | cut -c1-330 | tail -30
echo "== sweep (which path for how many queries)"; SWEEP_MAX_NQ=1 timeout 600 python scripts/sweep_batch_paths.py 1000000 128 10 2>&1 | cut -c1-400 | tail -9
