"""K11 against K1 against a torch fp64 brute force on the device, 300k and 2M rows x 768: prints the three answers side by
side (the run that showed the shadow-built-under-a-discarded-capture bug: K11 returned rows 0..8).  GPU box only."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "simple-vector-db_b200")]
from svdb import binding as B
import torch
D = 768
for n in (300_000, 2_000_000):
    g = torch.Generator(device="cuda").manual_seed(5)
    with B.Engine(D, D) as e:
        parts = []
        for lo in range(0, n, 250_000):
            m = min(250_000, n - lo)
            part = torch.rand((m, D), dtype=torch.float64, device="cuda", generator=g)
            torch.cuda.synchronize()
            e.insert_device(part.data_ptr(), m, D)
            parts.append(part)
        Q = np.random.default_rng(6).random((2, D))
        out = {"n": n}
        for label, on in (("k1", 0), ("k11", 1)):
            e.set_option("scan.shadow", on)
            idx, dist, seq = e.nearest(Q, 3)
            out[label] = {"seq": seq.tolist(), "dist": dist.tolist(), "idx": idx.tolist()}
        q = torch.from_numpy(Q).cuda()
        best = []
        for qi in range(2):
            dmin, arg, base = None, None, 0
            ds = []
            for part in parts:
                d = ((part - q[qi]) ** 2).sum(1)
                ds.append(d)
            d = torch.cat(ds)
            v, i = torch.topk(d, 3, largest=False)
            best.append({"seq": i.tolist(), "dist": v.tolist()})
        out["torch"] = best
        out["reruns"] = e.stats()["exact_reruns"]
        print(json.dumps(out), flush=True)
