"""debug: which stage of test_topk_vs_oracle[4099-20-13-10-21] disagrees with the oracle"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "simple-vector-db_b200"), os.path.join(ROOT, "tests")]
import numpy as np
from svdb import binding as B, synth
from oracle import binding as OB
port = OB.load_port()
from oracle.binding import PortDB

def oracle_topk(rows, K, Q, k):
    db = PortDB(port, rows.shape[1], K)
    for r in rows:
        db.insert(r)
    out = [db.topk(q, k) for q in Q]
    db.close()
    return out

def check(tag, got, want, k):
    idx, dist, seq = got
    bad = []
    for i, (wseq, widx, wd) in enumerate(want):
        m = len(wseq)
        if not np.array_equal(seq[i, :m].astype(np.int64), wseq):
            bad.append((i, seq[i, :m].tolist(), list(map(int, wseq))))
    print(tag, "OK" if not bad else f"BAD {len(bad)}: {bad[:2]}", flush=True)

n, D, K, k, seed = 4099, 20, 13, 10, 21
rows = synth.uniform_rows(seed, n, D)
Q = synth.uniform_rows(seed + 50, 13, D)
want = oracle_topk(rows, K, Q, k)
for fuse, graphs in ((1, 1), (1, 0), (0, 1)):
    with B.Engine(D, K) as e:
        e.insert(rows)
        e.set_option("scan.fuse_tail", fuse)
        e.set_option("host.graphs", graphs)
        check(f"fuse={fuse} g={graphs} first(K2)", e.nearest(Q, k), want, k)
        e.set_option("nearest.mma_min_queries", 0)
        check(f"fuse={fuse} g={graphs} K1 4+4+4+1", e.nearest(Q, k), want, k)
        for opts in ({"scan.variant": 1}, {"scan.variant": 0, "scan.nq_per_pass": 1},
                     {"scan.nq_per_pass": 8, "scan.warps": 4, "scan.stages": 2},
                     {"scan.tile_rows": 1, "scan.ctas_per_sm": 2}, {"scan.force_exact": 1}):
            for name, v in opts.items():
                e.set_option(name, v)
            check(f"fuse={fuse} g={graphs} {opts}", e.nearest(Q, k), want, k)
            check(f"fuse={fuse} g={graphs} {opts} again", e.nearest(Q, k), want, k)
