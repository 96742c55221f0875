#!/usr/bin/env python
"""Measurements for the BASELINE.json configs that are not the bench.py headline (1 GPU).

    python scripts/bench_extra.py c1 c2 c4 c5 [--out gpurun_out/extra.jsonl]

Every number is device-timed with CUDA events on the stream the kernels run on, after
warm-up; GB/s figures divide ALGORITHMIC bytes by that time.
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

from svdb import binding as B  # noqa: E402

DEV = torch.device("cuda", 0)
PEAK = 6542.7
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def chunks(n: int, D: int, seed: int, chunk: int = 250_000):
    """The synthetic rows, regenerable from (seed, chunk number): the parity checks below never read rows back from the
    engine, they regenerate them."""
    done = 0
    c = 0
    while done < n:
        m = min(chunk, n - done)
        g = torch.Generator(device=DEV).manual_seed(seed * 1_000_003 + c)
        yield done, torch.rand((m, D), dtype=torch.float64, device=DEV, generator=g)
        done += m
        c += 1


def fill(e: B.Engine, n: int, D: int, seed: int, chunk: int = 250_000):
    for _, t in chunks(n, D, seed, chunk):
        e.insert_device(t.data_ptr(), t.shape[0], D)
        torch.cuda.synchronize()
        del t
    torch.cuda.empty_cache()


PLANE_BYTES = {0: 8, 1: 4, 2: 2, 3: 1}      # bytes per coordinate of the copy of the log a single-query scan streams (engine stat scan_plane_last)


def nearest_parity(e: B.Engine, n: int, D: int, K: int, seed: int, k: int, nq: int = 8):
    """ids + fp64 distance bits of the engine's top-k == the CPU oracle's on an independent brute force over the WHOLE
    store (oracle/bigcheck.py), for nq single queries and for the same queries as one batch call."""
    from oracle import bigcheck
    g = torch.Generator(device=DEV).manual_seed(seed + 4242)
    Q = torch.rand((nq, D), dtype=torch.float64, device=DEV, generator=g)
    Qh = Q.cpu().numpy()
    one = [e.nearest(Qh[i], k) for i in range(nq)]
    ids = np.stack([o[0][0] for o in one])
    dist = np.stack([o[1][0] for o in one])
    bi, bd, _ = e.nearest(Qh, k)
    cand = bigcheck.brute_candidates(chunks(n, D, seed), Q[:, :K].contiguous(), 64)
    v = bigcheck.verdict([cand], Qh[:, :K], ids, dist, k, n)
    vb = bigcheck.verdict([cand], Qh[:, :K], bi, bd, k, n)
    v["batch_call_ok"] = vb["ok"]
    v["ok"] = v["ok"] and vb["ok"]
    return v


def timed(fn, iters: int, warm: int = 3) -> float:
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


def nearest_case(name, n, D, K, k, nqs, iters=20, extra_opts=(), parity=False):
    out = []
    seed = hash(name) % 1000 if not parity else sum(map(ord, name)) % 1000      # str hashes are salted per process
    with B.Engine(D, K, reserve_rows=n) as e:
        e.set_stream(torch.cuda.current_stream().cuda_stream)
        t0 = time.perf_counter()
        fill(e, n, D, seed=seed)
        e.flush()
        torch.cuda.synchronize()
        ingest_s = time.perf_counter() - t0
        for o, v in extra_opts:
            e.set_option(o, v)
        for nq in nqs:
            q = torch.rand((nq, D), dtype=torch.float64, device=DEV)
            res = torch.zeros((nq, k, 4), dtype=torch.int64, device=DEV)
            ms = timed(lambda: e.nearest_device(q.data_ptr(), nq, D, k, res.data_ptr()), iters)
            # the scan launches of one call, timed by CUDA events on the engine's stream
            e.set_option("profile.scan_events", 1)
            e.take_scan_time()
            for _ in range(5):
                e.nearest_device(q.data_ptr(), nq, D, k, res.data_ptr())
            scan_total, launches = e.take_scan_time()
            e.set_option("profile.scan_events", 0)
            st = e.stats()
            plane = int(st["scan_plane_last"])
            passes = max(1, int(launches) // 5)
            scan_ms = scan_total / max(1, launches)
            kp = -(-K // 64) * 64
            algo = n * K * 8 if plane == 0 else n * kp * PLANE_BYTES[plane]
            qh = q.cpu().numpy()
            for _ in range(3):                 # pinned buffers, the capture of the call's CUDA graph
                e.nearest(qh, k)
            nrep = 20 if ms < 5 else 5
            t0 = time.perf_counter()
            for _ in range(nrep):
                e.nearest(qh, k)
            e2e_ms = (time.perf_counter() - t0) / nrep * 1e3
            out.append({"config": name, "rows": n, "dim": D, "kd_dim": K, "k": k, "queries_per_call": nq,
                        "ms_per_call": ms, "queries_per_s": nq / ms * 1e3, "e2e_queries_per_s": nq / e2e_ms * 1e3,
                        "scan_launches_per_call": passes, "scan_ms_per_launch": scan_ms,
                        "scan_reads": {0: "fp64 rows (or a tree / tensor-core path)", 1: "hi + lo bf16 planes", 2: "bf16 hi plane", 3: "one-byte plane"}[plane],
                        "scan_GBps_algorithmic": algo / scan_ms / 1e6 if scan_ms > 0 else None,
                        "frac_of_measured_peak": algo / scan_ms / 1e6 / PEAK if scan_ms > 0 else None,
                        "fp64_pass_ms_at_peak": n * K * 8 / PEAK / 1e6, "ingest_s": ingest_s,
                        "hbm_gib_mapped": st["hbm_bytes_mapped"] / 2**30, "tree_rounds": st["tree_rounds"],
                        "mtree_builds": st["mtree_builds"]})
            print(json.dumps(out[-1]), flush=True)
        if parity:
            torch.cuda.empty_cache()
            v = nearest_parity(e, n, D, K, seed, k)
            v["config"] = name + "_parity_check"
            st = e.stats()
            v["exact_reruns"], v["fp64_reruns"] = st["exact_reruns"], st["fp64_reruns"]
            out.append(v)
            print(json.dumps(v), flush=True)
    return out


def compare_parity(e: B.Engine, n: int, D: int, seed: int, i1, i2, got, npick: int = 2000):
    """fp32 BIT patterns of the three metrics for `npick` of the benchmarked pairs == the reference's own functions
    (oracle/_ref when present, else the port) on rows regenerated from their seeds."""
    from oracle import binding as OB
    ref = OB.load_ref() if OB.have_ref() else None
    port = OB.load_port()
    pick = np.random.default_rng(11).choice(len(i1), npick, replace=False)
    a_idx, b_idx = i1[pick], i2[pick]
    need = np.unique(np.concatenate([a_idx, b_idx]))
    rows = {}
    for first, t in chunks(n, D, seed):
        m = need[(need >= first) & (need < first + t.shape[0])]
        if len(m):
            sub = t[torch.from_numpy((m - first).astype(np.int64)).to(DEV)].cpu().numpy()
            rows.update({int(r): sub[j] for j, r in enumerate(m)})
        del t
    bad = 0
    for j, (a, b) in enumerate(zip(a_idx, b_idx)):
        for metric in range(3):
            want = (ref or port).metric(metric, rows[int(a)], rows[int(b)])
            bad += int(np.float32(want).view(np.uint32) != got[pick[j], metric].view(np.uint32))
    return {"pairs_checked": int(npick), "metrics": 3, "mismatching_bit_patterns": int(bad), "ok": bad == 0,
            "oracle": ("the reference's own cosine_similarity / euclidean_distance / dot_product (oracle/_ref)" if ref
                       else "oracle/svdb_oracle.c port") + " on rows regenerated from their seeds; fp32 bit patterns compared with =="}


def compare_case(name, n, D, npairs, iters=5, parity=False):
    out = []
    with B.Engine(D, 1, reserve_rows=n, flags=B.FLAG_NO_LOG) as e:
        e.set_stream(torch.cuda.current_stream().cuda_stream)
        fill(e, n, D, seed=7)
        e.flush()
        g = torch.Generator(device=DEV).manual_seed(3)
        i1 = torch.randint(0, n, (npairs,), dtype=torch.int64, device=DEV, generator=g)
        i2 = torch.randint(0, n, (npairs,), dtype=torch.int64, device=DEV, generator=g)
        res = torch.zeros((npairs, 3), dtype=torch.float32, device=DEV)
        for metric, label in ((0, "cosine"), (1, "euclidean"), (2, "dot"), (3, "all3")):
            ms = timed(lambda: e.compare_device(metric, i1.data_ptr(), i2.data_ptr(), npairs, res.data_ptr()), iters)
            algo = npairs * 2 * D * 8
            out.append({"config": name, "rows": n, "dim": D, "pairs": npairs, "metric": label, "ms": ms,
                        "pairs_per_s": npairs / ms * 1e3, "GBps_algorithmic": algo / ms / 1e6,
                        "frac_of_measured_peak": algo / ms / 1e6 / PEAK})
            print(json.dumps(out[-1]), flush=True)
        if parity:
            e.compare_device(3, i1.data_ptr(), i2.data_ptr(), npairs, res.data_ptr())
            torch.cuda.synchronize()
            v = compare_parity(e, n, D, 7, i1.cpu().numpy(), i2.cpu().numpy(), res.cpu().numpy())
            v["config"] = name + "_parity_check"
            v["store_gb"] = n * D * 8 / 1e9
            out.append(v)
            print(json.dumps(v), flush=True)
        # single pair latency through the host call
        a = np.random.rand(D)
        b = np.random.rand(D)
        B.compare_vectors(0, a, b)
        t0 = time.perf_counter()
        for _ in range(20):
            B.compare_vectors(0, a, b)
        out.append({"config": name, "single_pair_host_call_us": (time.perf_counter() - t0) / 20 * 1e6, "dim": D})
        print(json.dumps(out[-1]), flush=True)
    return out


def cpu_case(name, n, D, K, nq, npairs=0, script=False):
    """The reference's own CPU path (oracle/_ref, else the port) on this box's cores, same shapes."""
    from oracle import binding as OB
    from svdb import synth
    drv = OB.load_cpu_driver()
    cores = os.cpu_count() or 1
    rows = synth.script_values(1, (n, D)) if script else synth.uniform_rows(1, n, D)
    Q = synth.script_values(2, (nq, D)) if script else synth.uniform_rows(2, nq, D)
    t0 = time.perf_counter()
    h = drv.build(rows, K)
    build_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    drv.nearest_batch(h, Q[: max(1, nq // cores)], 1)
    one_core = (time.perf_counter() - t0) / max(1, nq // cores)
    t0 = time.perf_counter()
    drv.nearest_batch(h, Q, cores)
    all_cores = time.perf_counter() - t0
    out = {"config": name, "impl": drv.kind, "rows": n, "dim": D, "kd_dim": K, "queries": nq, "cores": cores,
           "build_s": build_s, "one_core_s_per_query": one_core, "one_core_queries_per_s": 1.0 / one_core,
           "all_cores_queries_per_s": nq / all_cores}
    if npairs:
        i1 = np.random.default_rng(3).integers(0, n, npairs).astype(np.uint64)
        i2 = np.random.default_rng(4).integers(0, n, npairs).astype(np.uint64)
        for m, label in ((0, "cosine"), (1, "euclidean"), (2, "dot")):
            t0 = time.perf_counter()
            drv.compare_batch(h, m, i1, i2, cores)
            out[f"compare_{label}_pairs_per_s_all_cores"] = npairs / (time.perf_counter() - t0)
    drv.free(h)
    print(json.dumps(out), flush=True)
    return [out]


def main():
    which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["c1", "c2", "c4"]
    torch.cuda.set_device(0)
    torch.cuda.set_stream(torch.cuda.Stream(device=DEV))
    res = []
    if "cpu" in which:  # CPU reference on the same shapes (configs 1, 2, 4); bounded so it ends in ~2 min
        res += cpu_case("cpu_c1_10k_x128_k3", 10_000, 128, 3, 10_000, script=True)
        res += cpu_case("cpu_c1b_100k_x128_k3", 100_000, 128, 3, 10_000, script=True)
        res += cpu_case("cpu_c2_1M_x128", 1_000_000, 128, 128, 64, npairs=100_000)
        res += cpu_case("cpu_c4_100k_x1536_compare", 100_000, 1536, 3, 16, npairs=100_000)
    if "c1" in which:   # the reference's own CPU-runnable case: kd_dim 3 prefix of 128-dim rows
        res += nearest_case("c1_10k_x128_k3", 10_000, 128, 3, 1, (1, 1024, 65536), iters=50)
        res += nearest_case("c1b_100k_x128_k3", 100_000, 128, 3, 1, (1, 1024, 65536), iters=50)
    if "lat" in which:  # fixed per-query cost: a store so small that the scan itself is ~free
        res += nearest_case("latency_4k_x768", 4096, 768, 768, 1, (1,), iters=200)
        res += nearest_case("latency_4k_x768_top10", 4096, 768, 768, 10, (1,), iters=200)
        res += nearest_case("latency_4k_x128", 4096, 128, 128, 1, (1,), iters=200)
    if "thin" in which:  # the exact thread-per-entry scan (K1'), tree traversal switched off
        res += nearest_case("thin_20M_x16_k3_scan", 20_000_000, 16, 3, 10, (1, 8), iters=10,
                            extra_opts=(("nearest.tree_max_k", 0),))
        res += nearest_case("thin_20M_x16_k16_scan", 20_000_000, 16, 16, 1, (1,), iters=10,
                            extra_opts=(("nearest.tree_max_k", 0),))
    if "ksweep" in which:  # one query over a ~1.5 GB kd log at every kd_dim regime (rows are wider: the kd log is compact)
        for K in (9, 12, 16, 17, 24, 32, 40, 48, 56, 64, 80, 96, 100, 128, 200, 256, 384, 500):
            n = int(1.5e9 / (K * 8))
            res += nearest_case(f"ksweep_k{K}", n, K + 16, K, 1, (1, 8), iters=10)
    if "k6" in which:      # the tree traversal (K6) on a big thin store, one large call
        res += nearest_case("k6_10M_x16_k3", 10_000_000, 16, 3, 1, (65536, 1048576), iters=5)
    if "ktop" in which:    # result size: top-1 / 10 / 24 of one query and of a 64-query call
        for k in (1, 10, 24):
            res += nearest_case(f"ktop{k}_2M_x768", 2_000_000, 768, 768, k, (1, 64), iters=10)
            res += nearest_case(f"ktop{k}_4M_x32", 4_000_000, 48, 32, k, (1,), iters=10)
    if "nsweep" in which:  # store size: where does the fixed cost of a query stop mattering
        for n in (1_000, 10_000, 100_000, 1_000_000):
            res += nearest_case(f"nsweep_{n}_x768", n, 768, 768, 1, (1,), iters=50)
            res += nearest_case(f"nsweep_{n}_x128", n, 128, 128, 1, (1,), iters=50)
    if "dsweep" in which:  # /compare over a ~3 GB store at every row length; 2M random pairs (or 1M for long rows)
        for D in (3, 16, 48, 128, 200, 256, 768, 1536, 4096):
            n = int(3e9 / (max(D, 16) * 8))
            res += compare_case(f"dsweep_d{D}", n, D, 2_000_000 if D <= 768 else 1_000_000, iters=5)
    if "c2" in which:
        res += nearest_case("c2_1M_x128", 1_000_000, 128, 128, 1, (1, 8, 64), iters=50, parity=True)
        res += compare_case("c2_compare_1M_x128", 1_000_000, 128, 100_000)
    if "c4" in which:
        res += compare_case("c4_compare_1M_x1536", 1_000_000, 1536, 1_000_000, parity=True)
    if "c5" in which:
        res += nearest_case("c5_100M_x128_k128", 100_000_000, 128, 128, 1, (1, 8), iters=5, parity=True)
    if "c5k3" in which:
        res += nearest_case("c5_100M_x128_k3", 100_000_000, 128, 3, 1, (1, 1024), iters=5)
    out = "gpurun_out/extra.jsonl"
    for a in sys.argv[1:]:
        if a.startswith("--out="):
            out = a.split("=", 1)[1]
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    with open(out, "a") as f:
        for r in res:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
