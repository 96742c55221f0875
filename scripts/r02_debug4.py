import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "simple-vector-db_b200")]
import numpy as np, torch
from svdb import binding as B
from oracle import binding as OB
from oracle.binding import PortDB
n, D, K, levels = 60_000, 768, 768, 2
g = torch.Generator().manual_seed(n + levels)
rows = torch.randint(0, levels, (n, D), generator=g).to(torch.float64) / 2
Q = torch.randint(0, levels, (16, D), generator=g).to(torch.float64) / 2
Q[:, 0] += 0.25
rows_np, Qn = rows.numpy(), Q.numpy()
port = OB.load_port()
db = PortDB(port, D, K)
for r in rows_np: db.insert(r)
want = np.array([db.nearest(q) for q in Qn])
d = ((rows_np[:, None, :] - Qn[None, :, :]) ** 2).sum(-1)
print("ties at min per query:", [(int((d[:, j] == d[:, j].min()).sum())) for j in range(16)])
with B.Engine(D, K) as e:
    e.insert(rows_np)
    for plane in (3, 2, 0):
        e.set_option("scan.plane", plane)
        got16 = e.nearest(Qn, 1)[0][:, 0]
        got1 = np.array([e.nearest(Qn[i:i+1], 1)[0][0, 0] for i in range(16)])
        st = e.stats()
        print("plane", plane, "16-query call ok:", np.array_equal(got16, want), "single calls ok:", np.array_equal(got1, want),
              "bad single:", np.nonzero(got1 != want)[0], "fp64", st["fp64_reruns"], "exact", st["exact_reruns"], "tree", st["tree_reruns"], flush=True)
        if not np.array_equal(got1, want):
            i = int(np.nonzero(got1 != want)[0][0])
            print("   q", i, "got", got1[i], "want", want[i], "d got", d[got1[i], i], "d want", d[want[i], i])
        if not np.array_equal(got16, want):
            i = int(np.nonzero(got16 != want)[0][0])
            print("   16: q", i, "got", got16[i], "want", want[i], "d got", d[got16[i], i], "d want", d[want[i], i])
