import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "simple-vector-db_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
from svdb import binding as B
kind, K = "offset", 256
rng = np.random.Generator(np.random.PCG64(K + len(kind)))
n = 6000
rows, Q = 1000.0 + rng.random((n, K)), 1000.0 + rng.random((12, K))
with B.Engine(K, K) as e:
    e.insert(rows); e.flush()
    e.set_option("nearest.umma_min_queries", 0)
    e.set_option("scan.plane", 3)
    qd = torch.from_numpy(Q).cuda()
    out = torch.zeros((1, 5, 4), dtype=torch.int64, device="cuda")
    e.nearest_device(qd[0:1].data_ptr(), 1, K, 5, out.data_ptr(), 0)
    torch.cuda.synchronize()
    lo, step, err, raw = e.debug_plane8(2 * K)
    print("lo", lo, "step", step, "err", err, "expected lo", rows.min(), "step", (rows.max() - rows.min()) / 255)
    u = np.clip(np.rint((rows[:2] - lo) / step), 0, 255).astype(np.uint8)
    print("bytes match:", np.array_equal(raw.reshape(2, K), u), raw[:16], u[0, :16])
    e.set_option("scan.tail_debug", 1)
    for plane in (3, 0):
        e.set_option("scan.plane", plane)
        e.nearest_device(qd[0:1].data_ptr(), 1, K, 5, out.data_ptr(), 0)
        torch.cuda.synchronize()
        t = e.debug_tail_times(296)
        print("plane", plane, "E", t[16:17].view(np.float64), "lim", t[17:18].view(np.float64), "dk", t[18:19].view(np.float64), "found", t[19], "bound", t[20:21].view(np.float64))
