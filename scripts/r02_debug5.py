import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "simple-vector-db_b200")]
import numpy as np, torch, torch.distributed as dist
from svdb import binding as B
from svdb.sharded import ShardedIndex
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.cuda.set_stream(torch.cuda.Stream(device=dev))
n, D, K, levels = 60_000, 768, 768, 2
g = torch.Generator().manual_seed(n + levels)
rows = torch.randint(0, levels, (n, D), generator=g).to(torch.float64) / 2
Q = torch.randint(0, levels, (16, D), generator=g).to(torch.float64) / 2
Q[:, 0] += 0.25
Q = Q.pin_memory()
d = ((rows.numpy()[:, None, :] - Q.numpy()[None, :, :]) ** 2).sum(-1)
for plane in (3, 2, 0):
    idx = ShardedIndex(D, K, n, rank, world, local)
    idx.bind_current_stream()
    idx.ingest_device(rows[idx.lo:idx.hi].to(dev).contiguous())
    idx.engine.set_option("scan.plane", plane)
    ref = None
    if rank == 0:
        ref = B.Engine(D, K, device=local); ref.insert_device(rows.to(dev).data_ptr(), n, D)
    got16 = idx.nearest(Q, 1)
    ones = [idx.nearest(Q[i:i+1], 1) for i in range(16)]
    qd = Q.to(dev)
    devs = []
    for i in range(16):
        r = idx.nearest_device(qd[i:i+1], 1).clone(); torch.cuda.synchronize()
        devs.append(r.cpu().numpy().view(B.candidate_dtype).reshape(1, 1))
    if rank == 0:
        widx, wdist, wseq = ref.nearest(Q.numpy(), 1)
        for i in range(16):
            a, b, c = got16["seq"][i, 0], ones[i]["seq"][0, 0], devs[i]["seq"][0, 0]
            if a != wseq[i, 0] or b != wseq[i, 0]:
                print("plane", plane, "q", i, "want", wseq[i, 0], "16call", a, "single", b, "device", c, "flags dev", devs[i]["flags"][0, 0],
                      "ties at min:", np.nonzero(d[:, i] == d[:, i].min())[0], "lo/hi", idx.lo, idx.hi, flush=True)
        st = idx.engine.stats()
        print("plane", plane, "done; tie_events", st["tie_events"], "fp64", st["fp64_reruns"], "exact", st["exact_reruns"], flush=True)
        ref.close()
    idx.close()
dist.barrier(); dist.destroy_process_group()
