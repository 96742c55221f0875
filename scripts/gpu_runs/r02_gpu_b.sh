#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/r02_debug1.py 2>&1 | tail -60
echo "== pytest -m gpu (all)";  timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -40
