#!/bin/bash
# usage: gpu_mtree_multi.sh N -- the thin-store paths on N GPUs after the median tree became AUTO's choice:
# sharded parity check (incl. replicated thin stores) and the replicated-vs-row-shards thin bench
set -u
N=${1:-2}
mkdir -p gpurun_out
echo "== median tree tests (1 GPU)"; timeout 300 python -m pytest tests/test_gpu_median_tree.py -m gpu -q -x --timeout=300 -p no:cacheprovider 2>&1 | tail -5
echo "== A/B (1 GPU, 10M rows)"; timeout 300 python scripts/bench_mtree.py ncu --out=gpurun_out/mtree_v2.jsonl 2>&1 | tail -3 | cut -c1-900
echo "== sharded parity check on $N GPUs"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/check_sharded.py > gpurun_out/check_sharded_$N.json 2> gpurun_out/check_sharded_$N.err
echo "check exit $?"; cat gpurun_out/check_sharded_$N.json | cut -c1-1500; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/check_sharded_$N.err | tail -15
echo "== thin kd-points on $N GPUs: replicated kd log vs row shards"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29536 scripts/bench_thin_multi.py > gpurun_out/thin_multi_$N.json 2> gpurun_out/thin_multi_$N.err
echo "thin exit $?"; cat gpurun_out/thin_multi_$N.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/thin_multi_$N.err | tail -5
