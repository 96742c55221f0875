#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== plane scan tests"; timeout 900 python -m pytest tests/test_gpu_shadow_scan.py -m gpu -q --timeout=600 -p no:cacheprovider 2>&1 | tail -25
echo "== pytest -m gpu";  timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -8
echo "== c5 / c2 few queries"; timeout 1200 python scripts/bench_extra.py c2 c5 --out=gpurun_out/r02_extra_h.jsonl 2>&1 | cut -c1-300 | grep queries_per_call
