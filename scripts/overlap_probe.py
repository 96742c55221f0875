"""What does a chain of back-to-back K13 launches look like from the inside?  %globaltimer stamps of every CTA's start and
finish in the LAST launch of a chain (option scan.tail_debug), with and without scan.overlap_steps.
    python scripts/overlap_probe.py [rows] [dim] [chain]"""
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
    chain = int(sys.argv[3]) if len(sys.argv) > 3 else 12
    g = torch.Generator(device="cuda").manual_seed(5)
    with B.Engine(D, D) as e:
        for lo in range(0, n, 250_000):
            m = min(250_000, n - lo)
            part = torch.rand((m, D), dtype=torch.float64, device="cuda", generator=g)
            torch.cuda.synchronize()
            e.insert_device(part.data_ptr(), m, D)
            del part
        e.set_stream(torch.cuda.current_stream().cuda_stream)
        Q = torch.rand((chain, D), dtype=torch.float64, device="cuda", generator=g)
        out = torch.zeros((chain, 1, 4), dtype=torch.int64, device="cuda")
        e.set_option("scan.tail_debug", 1)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for overlap in (1, 0, 1, 0):
            e.set_option("scan.overlap_steps", overlap)
            for rep in range(3):
                torch.cuda.synchronize()
                ev0.record()
                for i in range(chain):
                    e.nearest_device(Q[i].data_ptr(), 1, D, 1, out[i].data_ptr())
                ev1.record()
                torch.cuda.synchronize()
            ncta = 295 if overlap else 296
            t = e.debug_tail_times(3 * ncta).astype(np.int64)
            done, start, smid = t[32:32 + ncta], t[32 + ncta:32 + 2 * ncta], t[32 + 2 * ncta:32 + 3 * ncta]
            s0 = start.min()
            dur = done - start
            per_sm = np.bincount(smid, minlength=148)
            slow = np.argsort(dur)[-8:]
            fast = np.argsort(dur)[:8]
            # duration of a CTA vs how many CTAs of this launch share its SM, and vs its SM's TPC partner
            by_share = {int(c): float(np.median(dur[per_sm[smid] == c])) / 1e3 for c in np.unique(per_sm[smid])}
            print(json.dumps({"overlap": overlap, "ctas_per_sm_histogram": np.bincount(per_sm).tolist(), "median_duration_us_by_ctas_on_the_sm": by_share,
                              "slowest": [(int(b), int(smid[b]), round(float(dur[b]) / 1e3, 1)) for b in slow],
                              "fastest": [(int(b), int(smid[b]), round(float(dur[b]) / 1e3, 1)) for b in fast]}), flush=True)
            print(json.dumps({"rows": n, "dim": D, "overlap": overlap, "chain": chain, "ms_per_step": ev0.elapsed_time(ev1) / chain,
                              "last_launch": {"cta_start_spread_us": float(start.max() - s0) / 1e3,
                                              "cta_start_p50_us": float(np.median(start) - s0) / 1e3,
                                              "cta_start_p90_us": float(np.percentile(start, 90) - s0) / 1e3,
                                              "cta_done_first_us": float(done.min() - s0) / 1e3, "cta_done_p50_us": float(np.median(done) - s0) / 1e3,
                                              "cta_done_last_us": float(done.max() - s0) / 1e3,
                                              "cta_duration_min_us": float(dur.min()) / 1e3, "cta_duration_p50_us": float(np.median(dur)) / 1e3,
                                              "cta_duration_max_us": float(dur.max()) / 1e3,
                                              "tail_end_us": float(t[2] - s0) / 1e3}}), flush=True)


if __name__ == "__main__":
    main()
