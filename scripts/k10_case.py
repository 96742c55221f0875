"""One K10 (tcgen05) batch call on a synthetic store, timed by CUDA events; the shape ncu captures of umma_filter_kernel use.
    python scripts/k10_case.py [rows] [dim] [queries] [k]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "simple-vector-db_b200")]
from svdb import binding as B  # noqa: E402


def main():
    import torch
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    D = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    nq = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    k = int(sys.argv[4]) if len(sys.argv) > 4 else 10
    g = torch.Generator(device="cuda").manual_seed(5)
    with B.Engine(D, D) as e:
        for lo in range(0, n, 250_000):
            m = min(250_000, n - lo)
            part = torch.rand((m, D), dtype=torch.float64, device="cuda", generator=g)
            torch.cuda.synchronize()
            e.insert_device(part.data_ptr(), m, D)
            del part
        e.set_stream(torch.cuda.current_stream().cuda_stream)
        e.set_option("nearest.umma_min_queries", 1)
        e.set_option("nearest.umma_min_kd_dim", 1)
        Q = torch.rand((nq, D), dtype=torch.float64, device="cuda", generator=g)
        out = torch.zeros((nq, k, 4), dtype=torch.int64, device="cuda")
        for _ in range(3):
            e.nearest_device(Q.data_ptr(), nq, D, k, out.data_ptr())
        torch.cuda.synchronize()
        e.set_option("profile.scan_events", 1)
        e.take_scan_time()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(10):
            e.nearest_device(Q.data_ptr(), nq, D, k, out.data_ptr())
        ev1.record()
        torch.cuda.synchronize()
        filt, launches = e.take_scan_time()
        kp = -(-D // 64) * 64
        print(json.dumps({"rows": n, "dim": D, "queries": nq, "k": k, "call_ms": ev0.elapsed_time(ev1) / 10,
                          "filter_kernel_ms": filt / max(1, launches), "shadow_bytes": n * kp * 4,
                          "shadow_pass_ms_at_7TBs": n * kp * 4 / 7.0e9, "tiles_per_cta": n / 128 / 148}))


if __name__ == "__main__":
    main()
