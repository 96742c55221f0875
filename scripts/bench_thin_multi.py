#!/usr/bin/env python
"""Thin kd-points (kd_dim = 3, the reference's default) on N GPUs: replicated kd log + tree, queries split.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29536 scripts/bench_thin_multi.py [rows] [queries_per_call]

Device-resident queries, CUDA events, max over ranks; one JSON line from rank 0.  Also runs the same call
with row shards (replicate_thin=False) for comparison: there every GPU scans its slice for every query.
"""
from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "simple-vector-db_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from svdb.sharded import ShardedIndex  # noqa: E402


def timed(idx, dq, k, iters, dev):
    for _ in range(3):
        idx.nearest_device(dq, k)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        idx.nearest_device(dq, k)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
    D, K = 16, 3
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    g = torch.Generator().manual_seed(5)
    out = {"bench": "thin_multi_gpu", "world": world, "rows": n, "dim": D, "kd_dim": K, "queries_per_call": nq}
    dq = torch.rand((nq, D), dtype=torch.float64, generator=g).to(dev)
    answers = {}
    for name, replicate, iters in (("replicated", True, 10), ("row_shards", False, 2)):
        idx = ShardedIndex(D, K, n, rank, world, local, replicate_thin=replicate)
        idx.bind_current_stream()
        gen = torch.Generator().manual_seed(1000 + rank)
        rows = torch.rand((idx.hi - idx.lo, D), dtype=torch.float64, generator=gen).to(dev)
        idx.ingest_device(rows)
        nqi = nq if replicate else min(nq, 1024)          # the scan path is ~100x slower per query: keep it short
        ms = timed(idx, dq[:nqi], 1, iters, dev)
        if replicate:
            # the same call answered by ONE replica alone (what a single GPU does), timed on every rank
            one = torch.zeros((nqi, 1, 4), dtype=torch.int64, device=dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for it in range(3 + iters):
                if it == 3:
                    e0.record()
                idx.engine.nearest_device(dq.data_ptr(), nqi, dq.stride(0), 1, one.data_ptr())
            e1.record()
            torch.cuda.synchronize()
            t1 = torch.tensor([e0.elapsed_time(e1) / iters], dtype=torch.float64, device=dev)
            dist.all_reduce(t1, op=dist.ReduceOp.MAX)
            out["one_replica_alone"] = {"queries": nqi, "ms_per_call": float(t1), "queries_per_s": nqi / (float(t1) / 1e3)}
        res = idx.nearest_device(dq[:1024], 1)
        torch.cuda.synchronize()
        answers[name] = res.clone()
        out[name] = {"queries": nqi, "ms_per_call": ms, "queries_per_s": nqi / (ms / 1e3)}
        idx.close()
    same = bool(torch.equal(answers["replicated"][..., :2], answers["row_shards"][..., :2]))     # dist bits + global seq
    out["replicated_equals_row_shards"] = same
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if same else 1)


if __name__ == "__main__":
    main()
