#!/bin/bash
# usage: gpu_multi.sh N  -- sharded parity check + bench on N GPUs of one box
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
echo "== sharded parity check on $N GPUs"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/check_sharded.py > gpurun_out/check_sharded_$N.json 2> gpurun_out/check_sharded_$N.err
echo "check exit $?"; cat gpurun_out/check_sharded_$N.json; tail -5 gpurun_out/check_sharded_$N.err
echo "== bench on $N GPUs"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench exit $?"; cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
