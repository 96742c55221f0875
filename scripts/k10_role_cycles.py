"""Where do the cycles of a K10 (tcgen05) batch call go?  CTA 0 counts, with clock64, what each of its three roles spends
waiting and working (csrc/umma_filter.cu, option umma.debug_keys): the TMA producer (waiting for a free stage), the MMA issuer
(waiting for the epilogue / for loads) and epilogue warp 4 (waiting for the accumulator, tcgen05.ld + thresholds, the
compares, appending survivors, barriers + pruning).  One JSON line, kilocycles.
    python scripts/k10_role_cycles.py [rows] [dim] [queries] [umma.group_min] [umma.sparse_checks]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "simple-vector-db_b200")]
from svdb import binding as B  # noqa: E402

NAMES = ["prod_wait_empty", "prod_total", "mma_wait_epilogue", "mma_wait_loads", "mma_total", "epi_wait_acc", "epi_chunks",
         "epi_bar_prune", "tiles", "epi_chunks_ldtm_thr", "epi_chunks_compare", "epi_chunks_append"]


def main():
    import torch
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    D = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    nq = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
    gm = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    sp = int(sys.argv[5]) if len(sys.argv) > 5 else 1
    k = 10
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
        e.set_option("umma.group_min", gm)
        e.set_option("umma.sparse_checks", sp)
        Q = torch.rand((nq, D), dtype=torch.float64, device="cuda", generator=g)
        out = torch.zeros((nq, k, 4), dtype=torch.int64, device="cuda")
        for _ in range(3):
            e.nearest_device(Q.data_ptr(), nq, D, k, out.data_ptr())
        e.set_option("umma.debug_keys", 1)
        e.nearest_device(Q.data_ptr(), nq, D, k, out.data_ptr())
        torch.cuda.synchronize()
        v = e.debug_filter_cycles()[:12]
        print(json.dumps({"rows": n, "dim": D, "nq": nq, "group_min": gm, "sparse_checks": sp, "kilocycles": {a: float(b) for a, b in zip(NAMES, v)}}))


if __name__ == "__main__":
    main()
