#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest -m gpu";  timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -5
echo "== sweep 20M x 128"; SWEEP_MAX_NQ=1 timeout 900 python scripts/sweep_batch_paths.py 20000000 128 10 > gpurun_out/r02_sweep_20Mx128.jsonl 2>gpurun_out/sw.err; tail -2 gpurun_out/sw.err
echo "== sweep 2M x 768"; SWEEP_MAX_NQ=1 timeout 900 python scripts/sweep_batch_paths.py 2000000 768 10 > gpurun_out/r02_sweep_2Mx768.jsonl 2>gpurun_out/sw.err; tail -2 gpurun_out/sw.err
echo "== sweep 1M x 128"; SWEEP_MAX_NQ=1 timeout 900 python scripts/sweep_batch_paths.py 1000000 128 10 > gpurun_out/r02_sweep_1Mx128.jsonl 2>gpurun_out/sw.err; tail -2 gpurun_out/sw.err
python - <<'PY'
import json
for f in ("20Mx128","2Mx768","1Mx128"):
    for l in open(f"gpurun_out/r02_sweep_{f}.jsonl"):
        d=json.loads(l); print(f, d["nq"], {k[:-3]:round(v,3) for k,v in d.items() if k.endswith("_ms")}, all(v for k,v in d.items() if k.endswith("identical")))
PY
