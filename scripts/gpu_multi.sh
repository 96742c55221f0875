#!/bin/bash
# usage: gpu_multi.sh N [quick]  -- sharded parity check + bench on N GPUs of one box
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
if [ "${2:-}" != "quick" ]; then
  echo "== single-GPU parity suite"
  timeout 1200 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -15
  echo "== single-GPU bench"
  timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_n1_again.json 2> gpurun_out/bench_n1_again.err; cat gpurun_out/bench_n1_again.json | cut -c1-400
fi
echo "== sharded parity check on $N GPUs"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/check_sharded.py > gpurun_out/check_sharded_$N.json 2> gpurun_out/check_sharded_$N.err
echo "check exit $?"; cat gpurun_out/check_sharded_$N.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/check_sharded_$N.err | tail -15
echo "== bench on $N GPUs (p2p exchange)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench exit $?"; cat gpurun_out/bench_n$N.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/bench_n$N.err | tail -8
echo "== bench on $N GPUs (nccl exchange)"
SVDB_EXCHANGE=nccl timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --steps 20 --warmup 3 --batch-queries 0 > gpurun_out/bench_n${N}_nccl.json 2> gpurun_out/bench_n${N}_nccl.err
echo "bench exit $?"; cat gpurun_out/bench_n${N}_nccl.json | cut -c1-330
echo "== thin kd-points on $N GPUs: replicated kd log vs row shards"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29536 scripts/bench_thin_multi.py > gpurun_out/thin_multi_$N.json 2> gpurun_out/thin_multi_$N.err
echo "thin exit $?"; cat gpurun_out/thin_multi_$N.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/thin_multi_$N.err | tail -5
