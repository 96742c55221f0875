#!/usr/bin/env python
"""Sweep the K1 launch shape (warps x stages x tile rows x CTAs/SM) on one GPU; prints GB/s per point."""
import itertools, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "simple-vector-db_b200")):
    sys.path.insert(0, p)
import torch
from svdb import binding as B
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from bench_extra import fill, DEV, PEAK

def main():
    n, D = int(sys.argv[1]), int(sys.argv[2])
    torch.cuda.set_device(0)
    torch.cuda.set_stream(torch.cuda.Stream(device=DEV))
    out = []
    with B.Engine(D, D, reserve_rows=n) as e:
        e.set_stream(torch.cuda.current_stream().cuda_stream)
        fill(e, n, D, seed=1)
        q = torch.rand((1, D), dtype=torch.float64, device=DEV)
        for warps, stages, tr, cps in itertools.product((4, 8, 12, 16), (2, 3, 4, 6), (0, 1, 2, 4), (1, 2)):
            for o, v in (("scan.warps", warps), ("scan.stages", stages), ("scan.tile_rows", tr), ("scan.ctas_per_sm", cps)):
                e.set_option(o, v)
            try:
                ms = e.time_scan(q.data_ptr(), 1, D, 1, 5)
            except B.SvdbError as ex:
                continue
            gbs = n * D * 8 / ms / 1e6
            out.append({"warps": warps, "stages": stages, "tile_rows": tr, "ctas_per_sm": cps, "ms": ms, "GBps": gbs, "frac": gbs / PEAK})
    out.sort(key=lambda r: -r["GBps"])
    for r in out[:12]:
        print(json.dumps(r))
    print("worst", json.dumps(out[-1]))
    with open(f"gpurun_out/sweep_scan_{n}x{D}.jsonl", "w") as f:
        for r in out:
            f.write(json.dumps(r) + "\n")

if __name__ == "__main__":
    main()
