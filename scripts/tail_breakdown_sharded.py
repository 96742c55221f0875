"""The fused tail of a SHARDED single-query step, phase by phase (rank 0's %globaltimer stamps; option scan.tail_debug).
    python -m torch.distributed.run --nproc-per-node N ... scripts/tail_breakdown_sharded.py [rows_per_rank] [dim]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "simple-vector-db_b200")]
from svdb.sharded import ShardedIndex  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    per = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
    D = int(sys.argv[2]) if len(sys.argv) > 2 else 768
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    idx = ShardedIndex(D, D, per * world, rank, world, local)
    idx.bind_current_stream()
    g = torch.Generator(device=dev).manual_seed(5 + rank)
    for lo in range(0, per, 250_000):
        m = min(250_000, per - lo)
        idx.ingest_device(torch.rand((m, D), dtype=torch.float64, device=dev, generator=g))
        torch.cuda.synchronize()
    e = idx.engine
    e.set_option("scan.tail_debug", 1)
    gq = torch.Generator().manual_seed(9)
    Q = torch.rand((40, D), dtype=torch.float64, generator=gq).to(dev)
    recs = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(40):
        dist.barrier()
        torch.cuda.synchronize()
        ev0.record()
        idx.nearest_device(Q[i:i + 1], 1)
        ev1.record()
        torch.cuda.synchronize()
        t = e.debug_tail_times(296).astype(np.int64)
        if i >= 8:
            recs.append({"event_us": ev0.elapsed_time(ev1) * 1e3, "scan_last_cta_us": (t[0] - t[6]) / 1e3, "finalize_us": (t[2] - t[0]) / 1e3,
                         "push_us": (t[3] - t[2]) / 1e3, "wait_peers_us": (t[4] - t[3]) / 1e3, "merge_us": (t[5] - t[4]) / 1e3,
                         "kernel_span_us": (t[5] - t[6]) / 1e3})
    # back-to-back steps, as bench.py times them
    dist.barrier()
    torch.cuda.synchronize()
    ev0.record()
    for i in range(40):
        idx.nearest_device(Q[i:i + 1], 1)
    ev1.record()
    torch.cuda.synchronize()
    out = {"world": world, "rows_per_rank": per, "dim": D, "rank": rank, "back_to_back_us_per_step": ev0.elapsed_time(ev1) * 1e3 / 40,
           **{k: float(np.median([r[k] for r in recs])) for k in recs[0]}}
    allo = [None] * world
    dist.all_gather_object(allo, out)
    if rank == 0:
        for o in allo:
            print(json.dumps(o), flush=True)
    dist.barrier()
    idx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
