"""Which kernel should answer a call of nq queries?  Times one device-resident call (CUDA events, queries in HBM) through
each path on the same store and prints one JSON line per nq: K1 (fp64 rows, passes of <= 8 queries), K11 (hi + lo bf16
planes, passes of <= 8), K12 (hi plane, passes of <= 2), K13 (one-byte plane, one pass per query, k <= 16), K2 (FP64 DMMA),
K10 (tcgen05).  Answers of every path are
compared with K1's; `hbm_pass_ms` is one pass over the fp64 rows at the measured HBM peak.

    python scripts/sweep_batch_paths.py [rows] [dim] [k]          # defaults 2000000 768 10
AUTO's thresholds (nearest.mma_min_queries, nearest.umma_min_queries, scan.plane) are set from this sweep
(profiles/r02_sweep_batch_paths_*.jsonl)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "simple-vector-db_b200")]
from svdb import binding as B  # noqa: E402

PATHS = {           # name: (scan.plane, nearest.mma_min_queries, nearest.umma_min_queries)
    "K1_fp64_rows": (0, 0, 0),
    "K11_hi_lo_planes": (1, 0, 0),
    "K12_hi_plane": (2, 0, 0),
    "K13_byte_plane": (3, 0, 0),
    "K2_dmma": (0, 1, 0),
    "K10_tcgen05": (0, 0, 1),
}
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:  # noqa: BLE001
    PEAK = 6650.0


def main():
    import torch
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
    D = int(sys.argv[2]) if len(sys.argv) > 2 else 768
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    g = torch.Generator(device="cuda").manual_seed(5)
    with B.Engine(D, D) as e:
        for lo in range(0, n, 250_000):
            m = min(250_000, n - lo)
            part = torch.rand((m, D), dtype=torch.float64, device="cuda", generator=g)
            torch.cuda.synchronize()
            e.insert_device(part.data_ptr(), m, D)
            del part
        e.set_stream(torch.cuda.current_stream().cuda_stream)
        e.set_option("nearest.umma_min_kd_dim", 1)
        e.set_option("scan.plane8_max_queries", 64)
        Qall = torch.rand((1024, D), dtype=torch.float64, device="cuda", generator=g)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for nq in (1, 2, 3, 4, 8, 16, 32, 64, 65, 128, 256, 1024):
            if nq > 64 and os.environ.get("SWEEP_MAX_NQ"):
                break
            Q = Qall[:nq].contiguous()
            out = torch.zeros((nq, k, 4), dtype=torch.int64, device="cuda")
            line = {"rows": n, "dim": D, "k": k, "nq": nq, "hbm_pass_ms": n * D * 8 / PEAK / 1e6}
            want = None
            for name, (shadow, mma, umma) in PATHS.items():
                if name in ("K1_fp64_rows", "K11_hi_lo_planes") and nq > 256 or name in ("K12_hi_plane", "K13_byte_plane") and nq > 64:
                    continue                                   # too many passes: not a contender
                e.set_option("scan.plane", shadow)
                e.set_option("nearest.mma_min_queries", mma)
                e.set_option("nearest.umma_min_queries", umma)
                e.nearest_device(Q.data_ptr(), nq, D, k, out.data_ptr())          # warm-up (builds the shadow once)
                torch.cuda.synchronize()
                reps = 3 if nq >= 256 else 10
                ev0.record()
                for _ in range(reps):
                    e.nearest_device(Q.data_ptr(), nq, D, k, out.data_ptr())
                ev1.record()
                torch.cuda.synchronize()
                line[name + "_ms"] = ev0.elapsed_time(ev1) / reps
                res = out.clone()
                if want is None:
                    want = res
                line[name + "_identical"] = bool(torch.equal(res[..., :3], want[..., :3]))
            best = min((v, kname) for kname, v in line.items() if kname.endswith("_ms"))
            line["fastest"] = best[1][:-3]
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
