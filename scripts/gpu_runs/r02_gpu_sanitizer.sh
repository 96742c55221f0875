#!/bin/bash
# round 2: compute-sanitizer over the plane scans, the 128-candidate tail, the overlapped launches and K10
set -u
mkdir -p gpurun_out
export SVDB_ARENA=malloc
CS=/usr/local/cuda/bin/compute-sanitizer
PY="python -m pytest -q -x --timeout=3000 -p no:cacheprovider -m gpu"
echo "== memcheck: tests/test_gpu_shadow_scan.py"
timeout 2400 $CS --tool memcheck --error-exitcode 9 $PY tests/test_gpu_shadow_scan.py > gpurun_out/san_mem_plane.log 2>&1; echo "exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san_mem_plane.log | tail -3
echo "== memcheck: tests/test_gpu_umma.py"
timeout 1800 $CS --tool memcheck --error-exitcode 9 $PY tests/test_gpu_umma.py > gpurun_out/san_mem_umma.log 2>&1; echo "exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san_mem_umma.log | tail -3
echo "== racecheck: tail with many candidates, short-row K13, overlapped launches, follow-inserts"
timeout 2400 $CS --tool racecheck --error-exitcode 9 $PY tests/test_gpu_shadow_scan.py -k "many_candidates or overlap or follows_inserts or few_queries or (vs_oracle and 3-1)" > gpurun_out/san_race_plane.log 2>&1; echo "exit $?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/san_race_plane.log | tail -3
echo "== synccheck: same subset"
timeout 1800 $CS --tool synccheck --error-exitcode 9 $PY tests/test_gpu_shadow_scan.py -k "many_candidates or overlap or few_queries" > gpurun_out/san_sync_plane.log 2>&1; echo "exit $?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san_sync_plane.log | tail -3
