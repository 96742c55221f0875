#!/usr/bin/env python
"""A/B of the two tree paths for thin kd-points on one GPU: K6 (the reference-shaped tree, one thread per query)
against K8/K9 (balanced median tree, 32/16/8 lanes per query), same store, same queries, answers compared.

    python scripts/bench_mtree.py [small] [big] [huge] [--out=gpurun_out/mtree.jsonl]

Device-timed with CUDA events on the engine's stream after warm-up; the build is timed by wall clock around the
first query (it synchronizes)."""
from __future__ import annotations

import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "simple-vector-db_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

from svdb import binding as B  # noqa: E402

DEV = torch.device("cuda", 0)


def fill(e, n, D, seed, chunk=4_000_000):
    done = c = 0
    while done < n:
        m = min(chunk, n - done)
        g = torch.Generator(device=DEV).manual_seed(seed * 1_000_003 + c)
        t = torch.rand((m, D), dtype=torch.float64, device=DEV, generator=g)
        e.insert_device(t.data_ptr(), m, D)
        torch.cuda.synchronize()
        done += m
        c += 1
    torch.cuda.empty_cache()


def timed(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def case(name, n, D, K, nqs, out):
    with B.Engine(D, K, reserve_rows=n) as e:
        e.set_stream(torch.cuda.current_stream().cuda_stream)
        t0 = time.perf_counter()
        fill(e, n, D, seed=len(name))
        e.flush()
        torch.cuda.synchronize()
        ingest_s = time.perf_counter() - t0            # includes K5 (the reference-shaped tree) and the row copies
        q1 = torch.rand((1, D), dtype=torch.float64, device=DEV)
        r1 = torch.zeros((1, 4), dtype=torch.int64, device=DEV)
        e.set_option("nearest.mtree", 1)
        t0 = time.perf_counter()
        e.nearest_device(q1.data_ptr(), 1, D, 1, r1.data_ptr())
        torch.cuda.synchronize()
        build_ms = (time.perf_counter() - t0) * 1e3
        st = e.stats()
        rec = {"config": name, "rows": n, "kd_dim": K, "mtree_build_ms_wall": build_ms, "mtree_levels": st["mtree_levels"],
               "ingest_s_incl_K5": ingest_s, "K5_rounds": st["tree_rounds"]}
        print(json.dumps(rec), flush=True)
        out.append(rec)
        for nq in nqs:
            g = torch.Generator(device=DEV).manual_seed(nq)
            q = torch.rand((nq, D), dtype=torch.float64, device=DEV, generator=g)
            res6 = torch.zeros((nq, 4), dtype=torch.int64, device=DEV)
            res9 = torch.zeros((nq, 4), dtype=torch.int64, device=DEV)
            iters = 50 if nq <= 1024 else (10 if nq <= 65536 else 4)
            e.set_option("nearest.mtree", 0)
            ms6 = timed(lambda: e.nearest_device(q.data_ptr(), nq, D, 1, res6.data_ptr()), iters)
            rec = {"config": name, "rows": n, "kd_dim": K, "queries_per_call": nq, "K6_ms": ms6, "K6_qps": nq / ms6 * 1e3}
            e.set_option("nearest.mtree", 1)
            best = None
            for blk in (1, 3):                                      # split values in heap order / three levels per 64-byte block
                e.set_option("mtree.block_levels", blk)            # (the first warm-up call rebuilds the tree in that layout)
                for lanes in (32, 16, 8):
                    e.set_option("mtree.lanes", lanes)
                    ms9 = timed(lambda: e.nearest_device(q.data_ptr(), nq, D, 1, res9.data_ptr()), iters)
                    torch.cuda.synchronize()
                    tag = f"K9_blk{blk}_lanes{lanes}"
                    rec[f"{tag}_ms"] = ms9
                    rec[f"{tag}_qps"] = nq / ms9 * 1e3
                    rec[f"{tag}_identical_to_K6"] = bool(torch.equal(res6, res9))
                    best = ms9 if best is None else min(best, ms9)
            e.set_option("mtree.lanes", 0)
            ms9 = timed(lambda: e.nearest_device(q.data_ptr(), nq, D, 1, res9.data_ptr()), iters)
            rec["K9_default_ms"] = ms9
            rec["K9_default_qps"] = nq / ms9 * 1e3
            rec["best_speedup_vs_K6"] = ms6 / best
            rec["default_speedup_vs_K6"] = ms6 / ms9
            if nq <= 65536:                                         # top-10: K6's k-smallest traversal vs a warp per query
                k = 10
                r6 = torch.zeros((nq, k, 4), dtype=torch.int64, device=DEV)
                r9 = torch.zeros((nq, k, 4), dtype=torch.int64, device=DEV)
                e.set_option("nearest.mtree", 0)
                rec["K6_top10_ms"] = timed(lambda: e.nearest_device(q.data_ptr(), nq, D, k, r6.data_ptr()), max(2, iters // 5))
                e.set_option("nearest.mtree", 1)
                rec["K9_top10_ms"] = timed(lambda: e.nearest_device(q.data_ptr(), nq, D, k, r9.data_ptr()), max(2, iters // 5))
                torch.cuda.synchronize()
                rec["K9_top10_identical_to_K6"] = bool(torch.equal(r6, r9))
                rec["top10_speedup_vs_K6"] = rec["K6_top10_ms"] / rec["K9_top10_ms"]
            print(json.dumps(rec), flush=True)
            out.append(rec)


def main():
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["small", "big"]
    torch.cuda.set_device(0)
    torch.cuda.set_stream(torch.cuda.Stream(device=DEV))
    res = []
    if "small" in which:     # config 1 shapes: the reference's own default kd_dim = 3
        case("c1_10k_k3", 10_000, 16, 3, (1, 1024, 65536), res)
        case("c1b_100k_k3", 100_000, 16, 3, (1, 1024, 65536), res)
    if "big" in which:
        case("1M_k3", 1_000_000, 16, 3, (1024, 65536, 1048576), res)
        case("10M_k3", 10_000_000, 16, 3, (1, 1024, 65536, 1048576), res)
        case("10M_k8", 10_000_000, 16, 8, (65536,), res)
    if "ncu" in which:       # one shape for the profiler (scripts/gpu_mtree_round.sh)
        case("10M_k3", 10_000_000, 16, 3, (65536,), res)
    if "huge" in which:      # config 5's row count with thin kd-points (rows kept narrow: the kd log is what is searched)
        case("100M_k3", 100_000_000, 8, 3, (1024, 1048576), res)
    out = "gpurun_out/mtree.jsonl"
    for a in sys.argv[1:]:
        if a.startswith("--out="):
            out = a.split("=", 1)[1]
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    with open(out, "a") as f:
        for r in res:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
