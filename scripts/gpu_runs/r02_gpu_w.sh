#!/bin/bash
# round 2, final 2-GPU run: suite on a 2-GPU box, sharded parity check incl. two-query calls, bench
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu";  timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -4
echo "== smoke";          timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
N=2
echo "== sharded parity check on $N GPUs"
SVDB_CHECK_OUT=gpurun_out/r02_check_sharded_n$N.json timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/check_sharded.py > gpurun_out/check_sharded_$N.out 2> gpurun_out/check_sharded_$N.err
echo "check exit $?"; cut -c1-400 gpurun_out/r02_check_sharded_n$N.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/check_sharded_$N.err | tail -8
echo "== bench on $N GPUs"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench exit $?"; cut -c1-300 gpurun_out/r02_bench_n$N.json
echo "== bench on 1 GPU"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/b1.err; cut -c1-300 gpurun_out/r02_bench_n1_final.json
