"""K11 A/B on a GPU box: single-query scan over the fp64 rows (K1) against the scan over the split-bf16 shadow (K11),
scan kernel timed alone with CUDA events (option profile.scan_events).   python scripts/debug_shadow.py [rows] [dim]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "simple-vector-db_b200")]
from svdb import binding as B  # noqa: E402


def main():
    import torch
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
    D = int(sys.argv[2]) if len(sys.argv) > 2 else 768
    g = torch.Generator(device="cuda").manual_seed(5)
    out = {"rows": n, "dim": D}
    with B.Engine(D, D) as e:
        step = 250_000
        for lo in range(0, n, step):
            m = min(step, n - lo)
            part = torch.rand((m, D), dtype=torch.float64, device="cuda", generator=g)
            torch.cuda.synchronize()
            e.insert_device(part.data_ptr(), m, D)
        Q = np.random.default_rng(6).random((4, D))
        for nq, k in ((1, 1), (2, 10)):
            res = {}
            for label, on in (("k1_fp64_rows", 0), ("k11_shadow", 1)):
                e.set_option("scan.shadow", on)
                e.set_option("profile.scan_events", 0)
                ans = e.nearest(Q[:nq], k)
                e.set_option("profile.scan_events", 1)
                e.take_scan_time()
                for _ in range(5):
                    e.nearest(Q[:nq], k)
                ms, launches = e.take_scan_time()
                res[label] = {"scan_ms": ms / max(1, launches), "seq": ans[2].tolist(), "dist_bits": ans[1].view(np.uint64).tolist()}
            same = res["k1_fp64_rows"]["seq"] == res["k11_shadow"]["seq"] and res["k1_fp64_rows"]["dist_bits"] == res["k11_shadow"]["dist_bits"]
            kp = -(-D // 64) * 64
            out[f"nq{nq}_k{k}"] = {"k1_scan_ms": res["k1_fp64_rows"]["scan_ms"], "k11_scan_ms": res["k11_shadow"]["scan_ms"],
                                   "k1_gbs": n * D * 8 / res["k1_fp64_rows"]["scan_ms"] / 1e6,
                                   "k11_gbs_of_shadow_bytes": n * kp * 4 / res["k11_shadow"]["scan_ms"] / 1e6,
                                   "speedup": res["k1_fp64_rows"]["scan_ms"] / res["k11_shadow"]["scan_ms"], "identical": same}
        out["exact_reruns"] = e.stats()["exact_reruns"]
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
