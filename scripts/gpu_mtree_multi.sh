#!/bin/bash
# usage: gpu_mtree_multi.sh N -- the thin-store paths on N GPUs after the median tree became AUTO's choice:
# GPU suite on one GPU, then the sharded parity check (incl. thin row shards and replicated thin stores) and the
# replicated-vs-row-shards thin bench on N GPUs
set -u
N=${1:-2}
mkdir -p gpurun_out
echo "== median tree tests (1 GPU)"; timeout 300 python -m pytest tests/test_gpu_median_tree.py -m gpu -q -x --timeout=300 -p no:cacheprovider 2>&1 | tail -5
echo "== A/B (1 GPU)"; rm -f gpurun_out/mtree_v3.jsonl; timeout 400 python scripts/bench_mtree.py small big --out=gpurun_out/mtree_v3.jsonl 2>&1 | tail -30 | cut -c1-2500
echo "== pytest -m gpu (all)"; timeout 900 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider 2>&1 | tail -15
echo "== sharded parity check on $N GPUs"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/check_sharded.py > gpurun_out/check_sharded_$N.json 2> gpurun_out/check_sharded_$N.err
echo "check exit $?"; cat gpurun_out/check_sharded_$N.json | cut -c1-600; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/check_sharded_$N.err | tail -15
echo "== thin kd-points on $N GPUs: replicated kd log vs row shards"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29536 scripts/bench_thin_multi.py > gpurun_out/thin_multi_$N.json 2> gpurun_out/thin_multi_$N.err
echo "thin exit $?"; cat gpurun_out/thin_multi_$N.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/thin_multi_$N.err | tail -5
