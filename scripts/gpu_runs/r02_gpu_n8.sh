#!/bin/bash
# round 2, one 8-GPU box: sharded parity check, the headline at N = 8, BASELINE config 5 where it is defined
set -u
N=${1:-8}
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
echo "== sharded parity check on $N GPUs"
SVDB_CHECK_OUT=gpurun_out/r02_check_sharded_n$N.json timeout 600 bash -c "$(declare -f run); N=$N; run 29533 scripts/check_sharded.py" > gpurun_out/check_sharded_$N.out 2> gpurun_out/check_sharded_$N.err
echo "check exit $?"; cut -c1-600 gpurun_out/r02_check_sharded_n$N.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/check_sharded_$N.err | tail -5
echo "== bench on $N GPUs: headline (10M x 768)"
timeout 600 bash -c "$(declare -f run); N=$N; run 29534 bench.py --gpus $N --steps 20 --warmup 5" > gpurun_out/r02_bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench exit $?"; cut -c1-1800 gpurun_out/r02_bench_n$N.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/bench_n$N.err | tail -4
echo "== config 5, kd_dim = 128: 100M x 128 row-sharded over $N GPUs"
timeout 900 bash -c "$(declare -f run); N=$N; run 29535 bench.py --gpus $N --steps 20 --warmup 5 --rows 100000000 --dim 128" > gpurun_out/r02_c5_k128_n$N.json 2> gpurun_out/c5_k128_n$N.err
echo "c5 exit $?"; cut -c1-1800 gpurun_out/r02_c5_k128_n$N.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/c5_k128_n$N.err | tail -4
echo "== config 5, kd_dim = 3 (the reference's default): GPU KD-tree builds + nearest, kd log replicated"
SVDB_OUT=gpurun_out/r02_c5_k3_n$N.json timeout 900 bash -c "$(declare -f run); N=$N; run 29536 scripts/c5_thin_sharded.py 100000000 128 1048576" > gpurun_out/c5_k3_n$N.out 2> gpurun_out/c5_k3_n$N.err
echo "c5 thin exit $?"; cut -c1-2500 gpurun_out/r02_c5_k3_n$N.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/c5_k3_n$N.err | tail -5
