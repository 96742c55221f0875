#!/usr/bin/env python
"""BASELINE.json config 5 with the reference's default kd_dim (3): 100M x 128 fp64 rows over the GPUs of one box, the
GPU KD-tree builds (K5: the reference-shaped insertion-order tree; K8: the balanced median tree by radix-select
partitioning) and nearest through them, with the parity check of oracle/bigcheck.py.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29537 \
        scripts/c5_thin_sharded.py [rows] [dim] [queries_per_call]

Thin kd-points are REPLICATED (svdb/sharded.py): every rank generates its row-range shard (dim doubles per row), the
8*kd_dim bytes per row that /nearest reads are all-gathered once, every GPU builds the trees over ALL rows, and the
queries of a call are split over the ranks.  One JSON line from rank 0.
"""
from __future__ import annotations

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "simple-vector-db_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from svdb.sharded import ShardedIndex  # noqa: E402

CHUNK = 250_000
SEED = 4


def shard_chunks(dev, lo, hi, D):
    c = lo // CHUNK
    while c * CHUNK < hi:
        g = torch.Generator(device=dev).manual_seed(SEED * 1_000_003 + c)
        t = torch.rand((CHUNK, D), dtype=torch.float64, device=dev, generator=g)
        a, b = max(lo, c * CHUNK), min(hi, (c + 1) * CHUNK)
        yield a, t[a - c * CHUNK: b - c * CHUNK]
        del t
        c += 1


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
    D = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    nq = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
    K = 3
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def tmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    idx = ShardedIndex(D, K, n, rank, world, local)
    idx.bind_current_stream()
    barrier()
    t0 = time.perf_counter()
    for _, part in shard_chunks(dev, idx.lo, idx.hi, D):
        idx.ingest_device(part.contiguous())
    if world == 1:
        idx.engine.flush()
    barrier()
    ingest_s = tmax(time.perf_counter() - t0)          # generation + all-gather of the kd-points + K5 (insertion-order tree)
    st = idx.engine.stats()
    g = torch.Generator().manual_seed(SEED + 9)
    q_host = torch.rand((nq, D), dtype=torch.float64, generator=g).pin_memory()
    dq = q_host.to(dev)
    barrier()
    t0 = time.perf_counter()
    idx.nearest_device(dq[:1024], 1)                   # the first query builds the median tree (K8)
    barrier()
    first_call_s = tmax(time.perf_counter() - t0)
    st2 = idx.engine.stats()
    out = {"bench": "config5_thin_kd_dim3", "world": world, "rows": n, "dim": D, "kd_dim": K,
           "layout": "replicated kd log + trees on every GPU, queries split" if idx.replicated else "one GPU",
           "ingest_s_incl_generation_allgather_K5": ingest_s, "K5_tree_rounds": st["tree_rounds"],
           "first_call_s_incl_K8_median_build": first_call_s, "K8_levels": st2["mtree_levels"], "K8_builds": st2["mtree_builds"],
           "K8_bytes_model": n * (K * 8 + 16 * int(st2["mtree_levels"])), "hbm_gib_mapped_per_gpu": st2["hbm_bytes_mapped"] / 2**30}
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for label, m in (("big_call", nq), ("call_1024", 1024), ("single_query", 1)):
        for _ in range(3):
            idx.nearest_device(dq[:m], 1)
        barrier()
        iters = 10 if m > 1024 else 50
        ev0.record()
        for _ in range(iters):
            idx.nearest_device(dq[:m], 1)
        ev1.record()
        barrier()
        ms = tmax(ev0.elapsed_time(ev1) / iters)
        out[label] = {"queries": m, "ms_per_call": ms, "queries_per_s": m / (ms / 1e3)}
    # parity: 16 queries against an independent brute force over every shard + the CPU oracle
    from oracle import bigcheck
    npq = 16
    got = idx.nearest(q_host[:npq], 1)
    cand = bigcheck.brute_candidates(shard_chunks(dev, idx.lo, idx.hi, D), dq[:npq, :K].contiguous(), 64)
    if world > 1:
        allc = [None] * world
        dist.all_gather_object(allc, cand)
    else:
        allc = [cand]
    if rank == 0:
        v = bigcheck.verdict(allc, q_host[:npq, :K].numpy().copy(), got["index"], got["dist"], 1, n)
        out["parity_check"] = v
        print(json.dumps(out), flush=True)
        if os.environ.get("SVDB_OUT"):
            with open(os.environ["SVDB_OUT"], "w") as f:
                f.write(json.dumps(out) + "\n")
    barrier()
    idx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
