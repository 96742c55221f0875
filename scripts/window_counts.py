"""How many rows does the re-rank window of a single-plane scan hold, by plane and k?  The coarser the plane, the wider
the window finalize derives from its measured error (tail.cuh: window()), and a window that overflows the tail's
FIN_NC = 128 candidate slots costs an fp64 scan on top.  For plane in (K12 bf16 hi, K13 one byte) and k in
1..24: per-call time (CUDA events), candidates found (dbg[19] of option scan.tail_debug), fp64 re-runs, and the
answers compared with K1's.  Sets the defaults of scan.plane_max_k / scan.plane8_max_k.
    python scripts/window_counts.py [rows] [dim]         # one JSON line per (plane, k)"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "simple-vector-db_b200")]
from svdb import binding as B  # noqa: E402


def main():
    import torch
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    D = int(sys.argv[2]) if len(sys.argv) > 2 else 768
    NQ = 24
    g = torch.Generator(device="cuda").manual_seed(5)
    with B.Engine(D, D) as e:
        for lo in range(0, n, 250_000):
            m = min(250_000, n - lo)
            part = torch.rand((m, D), dtype=torch.float64, device="cuda", generator=g)
            torch.cuda.synchronize()
            e.insert_device(part.data_ptr(), m, D)
            del part
        e.set_stream(torch.cuda.current_stream().cuda_stream)
        e.set_option("nearest.umma_min_queries", 0)
        e.set_option("nearest.mma_min_queries", 0)
        Q = torch.rand((NQ, D), dtype=torch.float64, device="cuda", generator=g)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for k in (1, 2, 4, 6, 10, 16, 24):
            out = torch.zeros((1, k, 4), dtype=torch.int64, device="cuda")
            want = []
            e.set_option("scan.plane", 0)
            e.set_option("scan.tail_debug", 0)
            t_k1 = []
            for i in range(NQ):
                ev0.record()
                e.nearest_device(Q[i:i + 1].data_ptr(), 1, D, k, out.data_ptr())
                ev1.record()
                torch.cuda.synchronize()
                t_k1.append(ev0.elapsed_time(ev1))
                want.append(out.clone())
            for plane in (1, 2, 3):
                e.set_option("scan.plane", plane)
                e.set_option("scan.plane_max_k", 24)
                e.set_option("scan.plane8_max_k", 24)
                e.set_option("scan.tail_debug", 1)
                e.nearest_device(Q[0:1].data_ptr(), 1, D, k, out.data_ptr())       # builds the plane
                torch.cuda.synchronize()
                r0 = e.stats()["fp64_reruns"]
                found, ms, same, unsafe = [], [], 0, 0
                for i in range(NQ):
                    ev0.record()
                    e.nearest_device(Q[i:i + 1].data_ptr(), 1, D, k, out.data_ptr())
                    ev1.record()
                    torch.cuda.synchronize()
                    ms.append(ev0.elapsed_time(ev1))
                    t = e.debug_tail_times(296)
                    found.append(int(t[19]))
                    flags = out[0, :, 3].cpu().numpy()
                    bad = bool((flags & B.CAND_UNSAFE).any()) if hasattr(B, "CAND_UNSAFE") else False
                    unsafe += bad
                    same += bool(torch.equal(out[..., :3], want[i][..., :3])) or bad
                # the device entry point returns unproven answers flagged; the host entry points re-answer them
                print(json.dumps({"rows": n, "dim": D, "plane": plane, "k": k, "queries": NQ,
                                  "ms_per_call_median": float(np.median(ms)), "K1_ms_per_call_median": float(np.median(t_k1)),
                                  "candidates_median": float(np.median(found)), "candidates_max": int(max(found)),
                                  "candidates": found, "unsafe_flagged": int(unsafe),
                                  "identical_to_K1_or_flagged": int(same),
                                  "fp64_reruns": int(e.stats()["fp64_reruns"] - r0)}), flush=True)


if __name__ == "__main__":
    main()
