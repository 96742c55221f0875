"""SURVEY.md s8f row 4 (JSON edge): what the O(D^2) `cJSON_GetArrayItem(json, i)` loop of the reference's nearest
handler (src/compare_handler.c:371-372) costs per request, and what integration/f3_f4_handlers.patch (one walk over the
array's child list) leaves.  Both handler builds run in-process over the reference's OWN L1 code (tests/c/fake_http
stands in for libmicrohttpd / cJSON: its arrays are linked lists exactly like cJSON's), on a 16-row store, so the
request time is the handler's parse + reply, not the search.  CPU only.

    python scripts/f4_json_walk_timing.py > profiles/r02_f4_json_walk_timing.jsonl
"""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRV = {"unpatched": os.path.join(ROOT, "oracle", "_ref", "handler_driver_ref"),
       "patched": os.path.join(ROOT, "oracle", "_ref", "handler_driver_patched_ref")}


def main():
    rng = np.random.Generator(np.random.PCG64(1))
    for D in (128, 768, 1536, 4096):
        nreq = 400 if D <= 1536 else 150
        with tempfile.TemporaryDirectory() as tmp:
            base = os.path.join(tmp, "base.txt")
            full = os.path.join(tmp, "full.txt")
            rows = ["POST /vector - " + json.dumps({"uuid": f"00000000-0000-0000-0000-{i:012d}", "vector": list(rng.random(D))})
                    for i in range(16)]
            open(base, "w").write("\n".join(rows) + "\n")
            reqs = ["POST /nearest - " + json.dumps(list(rng.random(D))) for _ in range(nreq)]
            open(full, "w").write("\n".join(rows + reqs) + "\n")
            line = {"D": D, "requests": nreq, "what": "POST /nearest through the reference's handler, reference L1, 16-row store, kd_dim 3"}
            for name, drv in DRV.items():
                best = {}
                for script in (base, full):
                    ts = []
                    for _ in range(5):
                        t0 = time.perf_counter()
                        subprocess.run([drv, script, "3", str(D)], stdout=subprocess.DEVNULL, check=True)
                        ts.append(time.perf_counter() - t0)
                    best[script] = min(ts)
                line[name + "_us_per_request"] = (best[full] - best[base]) / nreq * 1e6
            line["speedup"] = line["unpatched_us_per_request"] / line["patched_us_per_request"]
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
