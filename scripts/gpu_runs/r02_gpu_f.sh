#!/bin/bash
# round 2: K13 on short rows (packed), single-pass selection + 128 candidate slots in the tail, window sizes by plane and k
set -u
mkdir -p gpurun_out
echo "== plane scan tests"; timeout 900 python -m pytest tests/test_gpu_shadow_scan.py -m gpu -q -x --timeout=600 -p no:cacheprovider 2>&1 | tail -15
echo "== pytest -m gpu";  timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -6
echo "== window sizes 10M x 768"; timeout 900 python scripts/window_counts.py 10000000 768 > gpurun_out/r02_window_counts_10Mx768.jsonl 2> gpurun_out/wc.err; tail -3 gpurun_out/wc.err; cut -c1-420 gpurun_out/r02_window_counts_10Mx768.jsonl
echo "== window sizes 4M x 128"; timeout 600 python scripts/window_counts.py 4000000 128 > gpurun_out/r02_window_counts_4Mx128.jsonl 2> gpurun_out/wc2.err; tail -3 gpurun_out/wc2.err; cut -c1-420 gpurun_out/r02_window_counts_4Mx128.jsonl
echo "== tail breakdown K13 k=1, k=10"; timeout 300 python scripts/tail_breakdown.py 1250000 768 3 1 2>&1 | tail -1 | cut -c1-1200; timeout 300 python scripts/tail_breakdown.py 1250000 768 3 10 2>&1 | tail -1 | cut -c1-1200
