#!/bin/bash
# End-of-round run on one GPU box: whole GPU suite, smoke, headline bench (K10 and K2 batches side by side), K10 timing
# cases, one ncu --set full capture of the K10 kernel.   gpurun --timeout 900 -- 'bash scripts/gpu_final_round.sh'
set -u
mkdir -p gpurun_out
T0=$SECONDS
echo "== pytest -m gpu (all)"; timeout 900 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider 2>&1 | tail -8 | cut -c1-300
echo "== [$((SECONDS-T0)) s] smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== [$((SECONDS-T0)) s] bench (ours)"; timeout 600 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; cat gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
: > gpurun_out/umma_debug.jsonl
for c in time_1M time_1M_128; do
  echo "== [$((SECONDS-T0)) s] case $c"; timeout 150 python scripts/debug_umma.py $c 2> gpurun_out/umma_case.err | tee -a gpurun_out/umma_debug.jsonl | cut -c1-1500
done
echo "== [$((SECONDS-T0)) s] ncu: K10 at 1M x 768, 1024 queries"
timeout 240 ncu --set full --clock-control none --import-source on -k regex:umma_filter_kernel -c 1 -o gpurun_out/prof_umma -f \
    python scripts/debug_umma.py time_1M > gpurun_out/ncu_umma.log 2>&1; echo "exit $?"
echo "== [$((SECONDS-T0)) s] done"
