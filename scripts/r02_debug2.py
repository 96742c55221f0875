import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "simple-vector-db_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from svdb import binding as B
K, n = 768, 6000
for rep in range(3):
    for kind in ("uniform", "normal"):
        rng = np.random.Generator(np.random.PCG64(K + len(kind)))
        rows, Q = (rng.random((n, K)), rng.random((12, K))) if kind == "uniform" else (rng.standard_normal((n, K)), rng.standard_normal((12, K)))
        with B.Engine(K, K) as e:
            e.insert(rows)
            e.set_option("scan.plane", 3)
            e.set_option("nearest.umma_min_queries", 0)
            idx, dist, seq = e.nearest(Q[0:1], 5)
            lo, step, err, raw = e.debug_plane8(n * K)
            u = np.clip(np.rint((rows - lo) / step), 0, 255).astype(np.uint8)
            raw = raw.reshape(n, K)
            badrows = np.nonzero((raw != u).any(1))[0]
            d = ((rows - Q[0]) ** 2).sum(1)
            print(rep, kind, "lo", lo, rows.min(), "step", step, (rows.max() - rows.min()) / 255, "err", err, "bad rows", len(badrows), badrows[:10],
                  "got", seq[0], "want", np.argsort(d)[:5], flush=True)
            if len(badrows):
                r = badrows[0]
                print("   row", r, "raw", raw[r, :12], "exp", u[r, :12], "n mismatching coords", int((raw[r] != u[r]).sum()))
