#!/bin/bash
# round 2: K13 with two queries per pass -- tests, sweep
set -u
mkdir -p gpurun_out
echo "== plane scan tests"; timeout 900 python -m pytest tests/test_gpu_shadow_scan.py -m gpu -q -x --timeout=600 -p no:cacheprovider 2>&1 | tail -8
echo "== pytest -m gpu";  timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider 2>&1 | tail -4
for c in "2000000 768" "10000000 768" "1000000 256"; do SWEEP_MAX_NQ=1 timeout 400 python scripts/sweep_batch_paths.py $c 3 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)
    if d['nq'] in (1,2,3,4,8): print(d['rows'],d['dim'],d['nq'],{k[:-3]:round(v,3) for k,v in d.items() if k.endswith('_ms') and k[:3] in ('K13','K12','K10')}, all(v for k,v in d.items() if k.endswith('identical')))
"; done
