#!/bin/bash
set -u
mkdir -p gpurun_out
B="python bench.py --steps 50 --warmup 10 --batch-queries 0 --no-fp64-scan --no-parity-check --no-cpu-baseline"
pr() { python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['ms_per_step'],4), 'iso', round(d['roofline']['isolated_launch_ms'],4), 'launches', d['gpu_launches'])"; }
for rows in 10000000 1250000; do
echo "rows $rows"
$B --rows $rows 2>/dev/null | pr "overlap on, 295 CTAs   "
SVDB_PDL_GRID_FULL=1 $B --rows $rows 2>/dev/null | pr "overlap on, 296 CTAs   "
SVDB_PDL_NOATTR=1 $B --rows $rows 2>/dev/null | pr "no attribute, 295 CTAs "
$B --rows $rows --opt scan.overlap_steps=0 2>/dev/null | pr "overlap off, 296 CTAs  "
$B --rows $rows 2>/dev/null | pr "overlap on, 295 CTAs   "
$B --rows $rows --opt scan.overlap_steps=0 2>/dev/null | pr "overlap off, 296 CTAs  "
done
