#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu (chunk 16)"
timeout 900 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -8
echo "== pytest compare tests with chunk 32 / 64"
SVDB_CMP_CHUNK=32 timeout 600 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider -k "compare or metrics or million" 2>&1 | tail -4
SVDB_CMP_CHUNK=64 timeout 600 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider -k "compare or metrics or million" 2>&1 | tail -4
for ch in 16 32 64; do
  echo "== c4 compare with chunk $ch"
  SVDB_CMP_CHUNK=$ch timeout 600 python scripts/bench_extra.py c4 --out=gpurun_out/extra_c4_ch$ch.jsonl 2>&1 | grep -v single_pair | cut -c1-330
done
echo "== thin exact scan"
timeout 600 python scripts/bench_extra.py thin --out=gpurun_out/extra_thin.jsonl 2>&1 | cut -c1-420
