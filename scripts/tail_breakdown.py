"""Where does the time of a single-query step go once the scan is done?  Runs single-query device calls with option
scan.tail_debug = 1 and prints, from the %globaltimer stamps the fused tail leaves (kernels.h: TailArgs::dbg), the
spread of the CTAs' finish times and the phases of the tail.   python scripts/tail_breakdown.py [rows] [dim] [plane] [k]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "simple-vector-db_b200")]
from svdb import binding as B  # noqa: E402


def main():
    import torch
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
    D = int(sys.argv[2]) if len(sys.argv) > 2 else 768
    plane = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    k = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    g = torch.Generator(device="cuda").manual_seed(5)
    with B.Engine(D, D) as e:
        for lo in range(0, n, 250_000):
            m = min(250_000, n - lo)
            part = torch.rand((m, D), dtype=torch.float64, device="cuda", generator=g)
            torch.cuda.synchronize()
            e.insert_device(part.data_ptr(), m, D)
            del part
        e.set_stream(torch.cuda.current_stream().cuda_stream)
        e.set_option("scan.plane", plane)
        e.set_option("scan.tail_debug", 1)
        Q = torch.rand((32, D), dtype=torch.float64, device="cuda", generator=g)
        out = torch.zeros((1, k, 4), dtype=torch.int64, device="cuda")
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        recs = []
        for i in range(32):
            q = Q[i:i + 1]
            ev0.record()
            e.nearest_device(q.data_ptr(), 1, D, k, out.data_ptr())
            ev1.record()
            torch.cuda.synchronize()
            t = e.debug_tail_times(296).astype(np.int64)
            if i < 8:
                continue                                   # warm-up (shadow build, clocks)
            start, ctas = t[6], t[32:]
            recs.append({"event_us": ev0.elapsed_time(ev1) * 1e3, "first_cta_done_us": (ctas.min() - start) / 1e3,
                         "median_cta_done_us": (np.median(ctas) - start) / 1e3, "p90_cta_done_us": (np.percentile(ctas, 90) - start) / 1e3,
                         "last_cta_done_us": (t[0] - start) / 1e3, "lists_merged_us": (t[1] - t[0]) / 1e3,
                         "rerank_us": (t[2] - t[1]) / 1e3, "kernel_span_us": (t[2] - start) / 1e3,
                         "a_ticket_to_finalize_entry_us": (t[9] - t[0]) / 1e3, "b_lists_in_smem_us": (t[10] - t[9]) / 1e3,
                         "c_lists_merged_per_warp_us": (t[11] - t[10]) / 1e3, "d_query_norms_us": (t[12] - t[11]) / 1e3,
                         "e_cross_warp_merge_us": (t[1] - t[12]) / 1e3, "f_window_us": (t[13] - t[1]) / 1e3,
                         "g_rows_fetched_squared_us": (t[14] - t[13]) / 1e3, "h_serial_chain_rank_emit_us": (t[2] - t[14]) / 1e3})
        keys = recs[0].keys()
        print(json.dumps({"rows": n, "dim": D, "plane": plane, "k": k, "calls": len(recs),
                          **{kk: float(np.median([r[kk] for r in recs])) for kk in keys}}), flush=True)


if __name__ == "__main__":
    main()
