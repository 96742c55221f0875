#!/bin/bash
# A/B of scan launch options under bench conditions (same box, interleaved, 40 steps each)
run() { python bench.py --no-cpu-baseline --batch-queries 0 --steps 40 "$@" 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('%-60s step %.4f scan %.4f GB/s %.0f' % (' '.join(sys.argv[1:]), j['ms_per_step'], j['roofline']['avg_launch_ms'], j['roofline']['achieved']))" "$@"; }
for rep in 1 2; do
run
run --opt scan.assign=1
run --opt scan.assign=1 --opt scan.stages=4
run --opt scan.warps=4 --opt scan.ctas_per_sm=2 --opt scan.stages=2
run --opt scan.warps=4 --opt scan.ctas_per_sm=2 --opt scan.stages=2 --opt scan.assign=1
run --opt scan.warps=4 --opt scan.stages=6
run --opt scan.warps=16 --opt scan.stages=2
done
