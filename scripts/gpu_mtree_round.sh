#!/bin/bash
# One GPU box, one call: the median-tree parity tests first (fail fast), the K6 vs K9 A/B, then the whole GPU suite,
# smoke, the headline bench and one ncu capture of K9.   gpurun --timeout 1100 -- 'bash scripts/gpu_mtree_round.sh'
set -u
mkdir -p gpurun_out
T0=$SECONDS
echo "== pytest median tree";  timeout 400 python -m pytest tests/test_gpu_median_tree.py -m gpu -q -x --timeout=300 -p no:cacheprovider 2>&1 | tail -25
echo "== [$((SECONDS-T0)) s] A/B K6 vs K9"; rm -f gpurun_out/mtree.jsonl; timeout 420 python scripts/bench_mtree.py small big > gpurun_out/mtree.log 2>&1; tail -40 gpurun_out/mtree.log | cut -c1-1200
echo "== [$((SECONDS-T0)) s] pytest -m gpu (all)"; timeout 900 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider 2>&1 | tail -15
echo "== [$((SECONDS-T0)) s] thin-path tests with the median tree switched off (K6 as before)"
SVDB_MTREE=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=300 -p no:cacheprovider -k "ties or tree or grids or golden or distinct" 2>&1 | tail -5
echo "== [$((SECONDS-T0)) s] smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
if [ "${1:-}" != "quick" ]; then
echo "== [$((SECONDS-T0)) s] bench (ours)"; timeout 600 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; cat gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
echo "== [$((SECONDS-T0)) s] ncu: K9 on 10M thin rows, one launch of a 65536-query call"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mtree_nearest_kernel -s 6 -c 1 -o gpurun_out/prof_mtree \
    python scripts/bench_mtree.py ncu --out=gpurun_out/tmp.jsonl > gpurun_out/ncu_mtree.log 2>&1; echo "exit $?"
echo "== [$((SECONDS-T0)) s] huge (100M thin rows)"; timeout 300 python scripts/bench_mtree.py huge > gpurun_out/mtree_huge.log 2>&1; tail -5 gpurun_out/mtree_huge.log | cut -c1-1200
fi
echo "== [$((SECONDS-T0)) s] done"; ls -la gpurun_out
