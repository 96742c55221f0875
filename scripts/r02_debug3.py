import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "simple-vector-db_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from svdb import binding as B
from oracle import binding as OB
from oracle.binding import PortDB
port = OB.load_port()
def oracle_topk(rows, K, Q, k):
    db = PortDB(port, rows.shape[1], K)
    for r in rows: db.insert(r)
    out = [db.topk(q, k) for q in Q]
    db.close(); return out
for rep in range(2):
  for K in (256, 768, 1000):
    for kind in ["uniform", "normal", "offset", "heavy_tail", "constant_columns"]:
        rng = np.random.Generator(np.random.PCG64(K + len(kind)))
        n = 6000
        if kind == "uniform": rows, Q = rng.random((n, K)), rng.random((12, K))
        elif kind == "normal": rows, Q = rng.standard_normal((n, K)), rng.standard_normal((12, K))
        elif kind == "offset": rows, Q = 1000.0 + rng.random((n, K)), 1000.0 + rng.random((12, K))
        elif kind == "heavy_tail": rows, Q = rng.standard_cauchy((n, K)).clip(-1e6, 1e6), rng.standard_cauchy((12, K)).clip(-1e6, 1e6)
        else:
            rows, Q = rng.random((n, K)), rng.random((12, K)); rows[:, ::3] = 0.5; Q[:, ::3] = 0.5
        want = oracle_topk(rows, K, Q, 5)
        with B.Engine(K, K) as e:
            e.insert(rows)
            e.set_option("scan.plane", 3)
            e.set_option("nearest.umma_min_queries", 0)
            for i in range(12):
                s0 = e.stats()
                idx, dist, seq = e.nearest(Q[i:i + 1], 5)
                s1 = e.stats()
                ok = np.array_equal(seq[0].astype(np.int64), want[i][0])
                if not ok:
                    print("MISMATCH", rep, K, kind, "call", i, "got", seq[0], "want", want[i][0], "plane_last", s1["scan_plane_last"],
                          "fp64+", s1["fp64_reruns"] - s0["fp64_reruns"], "exact+", s1["exact_reruns"] - s0["exact_reruns"],
                          "launched+", s1["kernels_launched"] - s0["kernels_launched"], flush=True)
            st = e.stats()
            print(rep, K, kind, "done fp64", st["fp64_reruns"], "exact", st["exact_reruns"], "plane_last", st["scan_plane_last"], flush=True)
