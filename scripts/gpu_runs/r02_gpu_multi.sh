#!/bin/bash
# usage: r02_gpu_multi.sh N  -- sharded parity check (vs single engine AND vs the CPU oracle) + bench on N GPUs of one box
set -u
N=${1:-2}
mkdir -p gpurun_out
echo "== sharded parity check on $N GPUs"
SVDB_CHECK_OUT=gpurun_out/r02_check_sharded_n$N.json timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/check_sharded.py > gpurun_out/check_sharded_$N.out 2> gpurun_out/check_sharded_$N.err
echo "check exit $?"; cut -c1-1500 gpurun_out/r02_check_sharded_n$N.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/check_sharded_$N.err | tail -15
echo "== bench on $N GPUs (p2p exchange fused into the scan)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench exit $?"; cut -c1-2500 gpurun_out/r02_bench_n$N.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/bench_n$N.err | tail -8
echo "== same, tail not fused (option scan.fuse_tail=0: scan, finalize, push, merge as four launches)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --steps 20 --warmup 5 --batch-queries 0 --no-fp64-scan --no-parity-check --opt scan.fuse_tail=0 > gpurun_out/r02_bench_n${N}_unfused.json 2> gpurun_out/bench_n${N}_unfused.err
echo "bench exit $?"; cut -c1-400 gpurun_out/r02_bench_n${N}_unfused.json
