#!/usr/bin/env python
"""bench.py -- nearest queries/s + achieved HBM GB/s on 10M x 768 fp64 (BASELINE.json).

    python bench.py [--gpus N --steps K --warmup W]          ours (N=1), or under torchrun for N>1
    python bench.py --impl reference [...]                   the reference's CPU path on the host cores

A step = ONE single-query nearest (top-1, what /nearest answers) over the whole store:
one pass of the distance scan over every shard, the candidate exchange and the merge.
  value  queries/s with the query already in HBM (device-timed, CUDA events, max over ranks)
  e2e    the same through the public call with the query in pinned HOST memory and the
         result read back to the host every step
The store is row-sharded over the N GPUs (strong scaling: the 10M rows are fixed).
Also reported: a 1024-query top-10 batch (config 3) and, at N=1, the reference's CPU path
timed on this box's cores on a bounded prefix of the same rows.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "simple-vector-db_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "nearest_queries_per_s"
UNIT = "queries/s"
CHUNK_ROWS = 125_000          # generation granule; seeds are per chunk so any sharding sees the same rows
SEED = 2
L2_BYTES = 126 << 20


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU every 100 ms while active (NVML)."""

    REASONS = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
               0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
               0x100: "display_clock_setting"}

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thread = None
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as ex:  # pragma: no cover
            self.err = str(ex)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        if self.ok:
            self._stop.clear()
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread:
            self._stop.set()
            self._thread.join()
            self._thread = None

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_peaks() -> dict:
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def profile_traffic():
    """dram bytes per scan launch from the committed ncu captures ({kernel: {dram_bytes_per_launch, rows, source}})."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "scan_traffic.json")))
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------
# CPU arm: the reference's own kdtree_nearest on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_rows(n: int, D: int):
    """The first n rows of the synthetic store.  On a GPU box they are generated exactly like the GPU arm's store
    (torch.rand on the device with the per-chunk seeds, copied to the host), so both arms see the SAME rows; without a
    device (this arm also runs on CPU-only hosts) numpy PCG64 draws from the same distribution."""
    try:
        import torch
        if torch.cuda.is_available():
            dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
            parts = [part.cpu() for _, part in store_chunks(dev, 0, n, n, D)]
            return torch.cat(parts)[:n].numpy(), "identical to the GPU arm's rows (torch.rand on the device, per-chunk seeds)"
    except Exception:  # noqa: BLE001
        pass
    rng = np.random.Generator(np.random.PCG64(SEED))
    return rng.random((n, D), dtype=np.float64), "same distribution (numpy PCG64; no CUDA device to generate the GPU arm's rows)"


def run_cpu_reference(rows: np.ndarray, K: int, n_total: int, steps: int, warmup: int, budget_s: float):
    """Each step = `cores` independent queries, one per host thread, on the read-only tree built
    from `rows` (a prefix-sized sample of the workload).  Returns q/s on the sample and its linear
    extrapolation to n_total rows (at K=768 the tree search visits ~every node: SURVEY.md s6)."""
    from oracle import binding as OB
    drv = OB.load_cpu_driver()
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    h = drv.build(rows, K)
    build_s = time.perf_counter() - t0
    rng = np.random.Generator(np.random.PCG64(SEED + 1))
    Q = rng.random((cores, rows.shape[1]))
    t0 = time.perf_counter()
    drv.nearest_batch(h, Q[:1], 1)
    one = time.perf_counter() - t0
    # keep the whole arm inside the budget
    max_steps = max(1, int(budget_s / max(one * 1.3, 1e-6)))
    warmup = min(warmup, max(0, max_steps // 4))
    steps = max(1, min(steps, max_steps - warmup))
    for _ in range(warmup):
        drv.nearest_batch(h, Q, cores)
    t0 = time.perf_counter()
    for i in range(steps):
        drv.nearest_batch(h, np.roll(Q, i, axis=0), cores)
    dt = time.perf_counter() - t0
    ids = drv.nearest_batch(h, Q, cores)          # the reference's answers on the sample (compared with the GPU engine's)
    drv.free(h)
    qps_sample = steps * cores / dt
    scale = rows.shape[0] / n_total
    return {"kind": drv.kind, "cores": cores, "qps_sample": qps_sample, "qps_full": qps_sample * scale,
            "steps": steps, "warmup": warmup, "ms_per_step": dt / steps * 1e3, "build_s": build_s,
            "one_query_one_core_s": one, "Q": Q, "ids": ids}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = min(args.rows, args.cpu_sample_rows)
    rows, rows_how = cpu_rows(n_sample, args.dim)
    r = run_cpu_reference(rows, args.kd_dim, args.rows, args.steps, args.warmup, budget_s=90.0)
    sample = (f"first {n_sample} of {args.rows} rows ({rows_how}), {r['cores']} concurrent "
              f"queries per step on {r['cores']} threads; value = q/s on the sample x {n_sample}/{args.rows} "
              f"(linear in rows: at K={args.kd_dim} kdtree_nearest visits ~every node)")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["qps_full"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": r["steps"], "warmup": r["warmup"], "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": r["qps_full"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": sample,
                         "value_on_sample": r["qps_sample"], "one_query_one_core_s": r["one_query_one_core_s"]},
        "e2e": {"value": r["qps_full"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def store_chunks(dev, lo: int, hi: int, N: int, D: int):
    """The synthetic store, U[0,1) fp64, generated on the device chunk by chunk; seeds are per chunk, so any sharding
    (and the parity check, which regenerates the rows instead of reading them back) sees the same rows."""
    import torch
    c = lo // CHUNK_ROWS
    while c * CHUNK_ROWS < hi:
        g = torch.Generator(device=dev).manual_seed(SEED * 1_000_003 + c)
        chunk = torch.rand((CHUNK_ROWS, D), dtype=torch.float64, device=dev, generator=g)
        a, b = max(lo, c * CHUNK_ROWS), min(hi, (c + 1) * CHUNK_ROWS, N)
        yield a, chunk[a - c * CHUNK_ROWS: b - c * CHUNK_ROWS]
        del chunk
        c += 1


PLANE_KERNEL = {0: "scan_wide_kernel<TR,1> (K1, fp64 rows)", 1: "scan_shadow_kernel<1> (K11, hi + lo bf16 planes of the shadow)",
                2: "scan_plane_kernel<1,TRIPS,32,TR> (K12, bf16 hi plane of the shadow)",
                3: "scan_plane8_kernel<NQ,TRIPS,TR,LPR> (K13, one-byte plane, exact integer keys)"}


def plane_bytes(plane: int, rows: int, K: int) -> int:
    """Algorithmic bytes one scan launch reads: the copy of the log the kernel streams, once."""
    kp = -(-K // 64) * 64
    return rows * K * 8 if plane == 0 else rows * kp * {1: 4, 2: 2, 3: 1}[plane]


def workload_config(args):
    cfg = {(10_000_000, 768): "config3", (100_000_000, 128): "config5", (1_000_000, 128): "config2"}.get((args.rows, args.dim), "custom")
    return {"workload": f"{cfg}: {args.rows}x{args.dim} fp64 store, kd_dim={args.kd_dim}, single-query nearest top-1 "
                        f"per step, row-sharded over n_gpus",
            "rows": args.rows, "dim": args.dim, "kd_dim": args.kd_dim, "k": 1, "queries_per_step": 1,
            "parallelism": f"row-shards x{args.gpus}", "exchange": os.environ.get("SVDB_EXCHANGE", "p2p") if args.gpus > 1 else None,
            "l2": "bytes streamed per GPU and step >> 126 MB L2 (no flush needed)"
                  if plane_bytes(3, args.rows // max(1, args.gpus), args.kd_dim) > 4 * L2_BYTES else "store fits L2: flushed between steps"}


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def ours(args):
    import torch
    import torch.distributed as dist
    from svdb import binding as B
    from svdb.sharded import ShardedIndex

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
        args.gpus = world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # one explicit stream carries everything: our kernels, the NCCL exchange, the timing events
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))

    D, K, N = args.dim, args.kd_dim, args.rows
    idx = ShardedIndex(D, K, N, rank, world, local, exchange=os.environ.get("SVDB_EXCHANGE", "p2p"))
    idx.bind_current_stream()
    e = idx.engine
    for name, val in (kv.split("=") for kv in args.opt):
        e.set_option(name, int(val))

    # ---- synthetic store: U[0,1) fp64, generated on the device chunk by chunk -------------
    t0 = time.perf_counter()
    lo, hi = idx.lo, idx.hi
    for _, part in store_chunks(dev, lo, hi, N, D):
        idx.ingest_device(part)
        torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    torch.cuda.empty_cache()
    if rank == 0:
        log(f"[bench] store built: {hi - lo} rows/rank in {build_s:.1f}s, {e.stats()['hbm_bytes_mapped'] / 2**30:.1f} GiB mapped")

    flush_buf = None
    if plane_bytes(3, hi - lo, K) <= 4 * L2_BYTES:      # the smallest copy of the log a scan may stream
        flush_buf = torch.empty(2 * L2_BYTES, dtype=torch.uint8, device=dev)

    # ---- queries: a pool of distinct ones, in pinned host memory and in HBM --------------
    gq = torch.Generator().manual_seed(SEED + 7)
    pool = 64
    q_host = torch.rand((pool, 1, D), dtype=torch.float64, generator=gq).pin_memory()
    q_dev = q_host.to(dev)
    k = 1
    sampler = ClockSampler(local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    q_np = q_host.numpy()        # the same pinned host buffers, seen as plain host memory by the C-ABI

    def step_e2e(i):
        """The call a user makes.  1 GPU: the C-ABI host entry point svdb_nearest_batch (host query in,
        host result out; H2D, kernels, D2H and the sync all inside).  N GPUs: ShardedIndex.nearest
        (H2D of the query, scan, peer-memory exchange + merge, D2H of the merged result)."""
        if flush_buf is not None:
            flush_buf.zero_()
        if world == 1:
            ix, ds, sq = e.nearest(q_np[i % pool], k)
            return {"seq": sq, "dist": ds, "index": ix}
        return idx.nearest(q_host[i % pool], k)

    def step_dev(i):
        if flush_buf is not None:
            flush_buf.zero_()
        return idx.nearest_device(q_dev[i % pool], k)

    # ---- e2e: host query in, host result out, every step ----------------------------------
    for i in range(args.warmup):
        step_e2e(i)
    barrier()
    st0 = e.stats()
    sampler.start()
    t0 = time.perf_counter()
    for i in range(args.steps):
        last = step_e2e(args.warmup + i)
    barrier()
    e2e_s = time.perf_counter() - t0
    st1 = e.stats()
    e2e_launches = st1["kernels_launched"] - st0["kernels_launched"]

    # ---- device-resident: queries already in HBM, K steps back to back, CUDA events -------
    for i in range(args.warmup):
        step_dev(i)
    barrier()
    m0 = idx.merge_launches
    st0 = e.stats()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()           # `ncu --profile-from-start off` lists the timed region only
    ev0.record()
    for i in range(args.steps):
        step_dev(args.warmup + i)
    ev1.record()
    barrier()
    torch.cuda.profiler.stop()
    sampler.stop()
    dev_ms = ev0.elapsed_time(ev1)
    st1 = e.stats()
    launches = st1["kernels_launched"] - st0["kernels_launched"] + (idx.merge_launches - m0)
    # The scanning kernel alone, one launch at a time: K more steps, outside the timed region, with a CUDA event on either
    # side of every launch.  (Inside the timed region nothing sits between consecutive launches, so that the scan of one
    # step can start while the tail of the step before it still runs: engine option scan.overlap_steps.)
    e.set_option("profile.scan_events", 1)
    e.take_scan_time()
    for i in range(args.steps):
        step_dev(args.warmup + i)
    torch.cuda.synchronize()
    scan_ms, scan_launches = e.take_scan_time()
    e.set_option("profile.scan_events", 0)

    plane_used = int(e.stats()["scan_plane_last"])

    times = torch.tensor([dev_ms, e2e_s * 1e3, scan_ms / max(1, scan_launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, scan_ms_iso = (float(x) for x in times.cpu())
    # roofline time of the dominant kernel = the timed region / its launches when a step is exactly one launch (the scan with
    # its fused tail) -- conservative: whatever a step spends outside the kernel counts against it; else the isolated figure
    one_launch_per_step = int(launches) == args.steps
    scan_ms_avg = dev_ms / args.steps if one_launch_per_step else scan_ms_iso

    # ---- the same steps on the fp64 rows (K1, option scan.plane = 0): BASELINE.json's "fp64 scan at >= 80 % of HBM peak"
    # reading keeps its number next to the headline, whatever copy of the log the default path streams ----
    fp64_scan = None
    if plane_used != 0 and not args.no_fp64_scan:
        try:
            e.set_option("scan.plane", 0)
            for i in range(3):
                step_dev(i)
            barrier()
            e.set_option("profile.scan_events", 1)
            e.take_scan_time()
            fsampler = ClockSampler(local)
            fsampler.start()
            ev0.record()
            for i in range(args.steps):
                step_dev(args.warmup + i)
            ev1.record()
            barrier()
            fsampler.stop()
            f_ms = ev0.elapsed_time(ev1)
            f_scan_ms, f_launches = e.take_scan_time()
            e.set_option("profile.scan_events", 0)
            ft = torch.tensor([f_ms, f_scan_ms / max(1, f_launches)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ft, op=dist.ReduceOp.MAX)
            f_ms, f_scan_avg = (float(x) for x in ft.cpu())
            fb = plane_bytes(0, hi - lo, K)
            pk, pk_src = measured_peak_gbs()
            fp64_scan = {"workload": "the headline's steps with option scan.plane = 0: K1 streams the fp64 rows (8 bytes per coordinate)",
                         "dtype": "f64", "value": args.steps / (f_ms / 1e3), "unit": UNIT, "ms_per_step": f_ms / args.steps,
                         "roofline": {"bound": "hbm", "kernel": PLANE_KERNEL[0], "algorithmic_bytes_per_launch": fb,
                                      "avg_launch_ms": f_scan_avg, "achieved": fb / (f_scan_avg / 1e3) / 1e9 if f_scan_avg > 0 else 0.0,
                                      "peak": pk, "unit": "GB/s", "frac": fb / (f_scan_avg / 1e3) / 1e9 / pk if f_scan_avg > 0 else 0.0,
                                      "peak_source": pk_src, "launches_timed": int(f_launches)},
                         "clocks": fsampler.summary()}
        except Exception as ex:  # noqa: BLE001
            fp64_scan = {"error": f"{type(ex).__name__}: {ex}"}
        finally:
            e.set_option("scan.plane", 3 if not any(kv.startswith("scan.plane=") or kv.startswith("scan.shadow=") for kv in args.opt)
                         else plane_used)
            e.set_option("profile.scan_events", 0)

    # ---- config 3 batch: 1024 queries, top-10 -- K10 (tcgen05, split-bf16 keys) and K2 (FP64 DMMA) side by side ----
    batch = None
    batch_dmma = None
    batch_results = {}
    if args.batch_queries > 0:
        nb, kb = args.batch_queries, 10
        qb_host = torch.rand((nb, D), dtype=torch.float64, generator=gq).pin_memory()
        qb_dev = qb_host.to(dev)
        peaks = load_peaks()

        umma_default = 3 if K >= 256 else 32        # the engine's own default (engine.h)

        def run_batch(umma: bool):
            e.set_option("nearest.umma_min_queries", umma_default if umma else 0)
            # warm-up: K10 builds its bf16 shadow of the log and sizes its scratch on the first full call; the DMMA
            # path (0.6 s per full call) warms up on one query group
            idx.nearest_device(qb_dev if umma else qb_dev[:64], kb)
            barrier()
            l0 = e.stats()["kernels_launched"]
            bsampler = ClockSampler(local)
            bsampler.start()
            reps = 5 if umma else 1
            ev0.record()
            for _ in range(reps):
                idx.nearest_device(qb_dev, kb)
            ev1.record()
            barrier()
            bsampler.stop()
            b_launches = (e.stats()["kernels_launched"] - l0) // reps
            b_ms = torch.tensor([ev0.elapsed_time(ev1) / reps], dtype=torch.float64, device=dev)
            t0 = time.perf_counter()
            res_b = idx.nearest(qb_host, kb)
            barrier()
            b_e2e = torch.tensor([(time.perf_counter() - t0) * 1e3], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(b_ms, op=dist.ReduceOp.MAX)
                dist.all_reduce(b_e2e, op=dist.ReduceOp.MAX)
            sec = float(b_ms) / 1e3
            if umma:
                kp = -(-K // 64) * 64
                flops = 3 * 2.0 * nb * (hi - lo) * kp          # three bf16 products (hi*hi, hi*lo, lo*hi) per coordinate
                peak = peaks.get("bf16_tflops", 2250.0)
                roof = {"bound": "tensor", "kernel": "umma_filter_kernel (tcgen05.mma kind::f16, M=128, N=256)",
                        "achieved": flops / sec / 1e12, "peak": peak, "unit": "TFLOP/s", "frac": flops / sec / 1e12 / peak,
                        "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops, cuBLAS burst)" if "bf16_tflops" in peaks
                                       else "fallback: nominal dense bf16",
                        "algorithmic_tflops": 2.0 * nb * (hi - lo) * K / sec / 1e12,
                        "note": "whole step (query prep, filter, exact re-rank) averaged over 5 calls; executed flops = 3 split-bf16 "
                                "products x 2*queries*rows_per_rank*Kpad; algorithmic_tflops counts one fp64 product per coordinate"}
                what = "K10: split-bf16 keys on tcgen05 tensor cores (256 queries per CTA group) + exact fp64 re-rank"
            else:
                flops = 2.0 * nb * (hi - lo) * K
                roof = {"bound": "fp64 tensor (DMMA)", "achieved": flops / sec / 1e12, "peak": 37.1,
                        "unit": "TFLOP/s", "frac": flops / sec / 1e12 / 37.1,
                        "peak_source": "profiles/r01_fp64_peak_dfma_vs_dmma.txt (DMMA.8x8x4 microbenchmark on this pool)",
                        "note": "whole step incl. re-rank; per-GPU flops = 2*queries*rows_per_rank*K"}
                what = "K2: GEMM-form keys on FP64 DMMA (64 queries per CTA group) + exact re-rank"
            batch_results["K10" if umma else "K2"] = res_b
            return {"workload": f"{nb}-query batch, top-{kb}, {what}",
                    "dtype": "filter keys: split bf16 x 3 products, fp32 accumulate (tcgen05); answers: f64, reference operation order, "
                             "bit-identical to the f64 paths" if umma else "f64",
                    "queries": nb, "k": kb, "value": nb / sec, "unit": UNIT, "ms": float(b_ms),
                    "e2e": {"value": nb / (float(b_e2e) / 1e3), "unit": UNIT, "h2d_bytes": nb * D * 8, "d2h_bytes": nb * kb * 32},
                    "gpu_launches": int(b_launches), "roofline": roof, "clocks": bsampler.summary(),
                    "unsafe_flags": int(np.count_nonzero(res_b["flags"] & B.CAND_UNSAFE)),
                    "result_checksum": int(np.bitwise_xor.reduce(res_b["seq"].astype(np.uint64).ravel()))}

        # the batch figures are extras next to the headline: a failure here must not take the headline line with it
        try:
            batch = run_batch(True)
        except Exception as ex:  # noqa: BLE001
            batch = {"error": f"{type(ex).__name__}: {ex}"}
        try:
            batch_dmma = run_batch(False)
        except Exception as ex:  # noqa: BLE001
            batch_dmma = {"error": f"{type(ex).__name__}: {ex}"}
        if "result_checksum" in batch and "result_checksum" in batch_dmma:
            batch["identical_to_dmma_path"] = batch["result_checksum"] == batch_dmma["result_checksum"]
        e.set_option("nearest.umma_min_queries", umma_default)

    # ---- parity at THIS size: outside every timed region, at every N.  For 8 pool queries the product's
    # answers -- top-1 through the timed path's entry point and top-10 through the public host call -- and, for 8 queries of
    # the 1024-batch, K10's top-10, are checked against an independent brute force over the whole store + the CPU oracle
    # (oracle/bigcheck.py): ids and fp64 distance bits ==. ----
    parity = None
    if not args.no_parity_check:
        try:
            from oracle import bigcheck
            npq, kp10 = 8, 10
            def public_call(i, kk):
                """the call a user makes (step_e2e's entry point), with k = kk"""
                if world == 1:
                    ix, ds, sq = e.nearest(q_np[i], kk)
                    return {"index": ix, "dist": ds, "seq": sq}
                return idx.nearest(q_host[i], kk)

            got1 = [public_call(i, 1) for i in range(npq)]
            got10 = [public_call(i, kp10) for i in range(npq)]
            Qs = [q_dev[i, 0, :K] for i in range(npq)]
            have_batch = args.batch_queries >= npq and len(batch_results) > 0
            if have_batch:
                Qs += [qb_dev[i, :K] for i in range(npq)]
            Qd = torch.stack(Qs).contiguous()
            cand = bigcheck.brute_candidates(store_chunks(dev, lo, hi, N, D), Qd, 64)
            torch.cuda.empty_cache()
            if world > 1:
                allc = [None] * world
                dist.all_gather_object(allc, cand)
            else:
                allc = [cand]
            if rank == 0:
                Qn = Qd.cpu().numpy()
                ids10 = np.stack([g["index"][0] for g in got10])
                d10 = np.stack([g["dist"][0] for g in got10])
                parity = bigcheck.verdict(allc, Qn[:npq], ids10, d10, kp10, N)
                ids1 = np.stack([g["index"][0] for g in got1])
                d1 = np.stack([g["dist"][0] for g in got1])
                p1 = bigcheck.verdict(allc, Qn[:npq], ids1, d1, 1, N)
                parity["single_query_top1_ok"] = p1["ok"]
                parity["ok"] = parity["ok"] and p1["ok"]
                for name, rb in batch_results.items():             # the 1024-query batch: K10 (tcgen05) and K2 (DMMA) answers
                    pb = bigcheck.verdict([tuple(a[npq:] for a in c) for c in allc], Qn[npq:], rb["index"][:npq], rb["dist"][:npq], kp10, N)
                    parity[f"batch_{name}_top10_ok"] = pb["ok"]
                    parity["ok"] = parity["ok"] and pb["ok"]
                parity["exact_reruns"] = e.stats()["exact_reruns"]
                parity["fp64_reruns"] = e.stats()["fp64_reruns"]
        except Exception as ex:  # noqa: BLE001
            parity = {"ok": False, "error": f"{type(ex).__name__}: {ex}"}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- CPU baseline beside it (N=1 only) -------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        n_sample = min(N, args.cpu_sample_rows)
        rows = torch.cat([part.cpu() for _, part in store_chunks(dev, 0, n_sample, n_sample, D)])[:n_sample].numpy()
        r = run_cpu_reference(rows, K, N, steps=3, warmup=1, budget_s=25.0)
        cpu = {"value": r["qps_full"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
               "sample": f"first {n_sample} of {N} rows of the same store, {r['cores']} concurrent queries/step x {r['steps']} steps "
                         f"on {r['cores']} threads; value = measured {r['qps_sample']:.3f} q/s x {n_sample}/{N} (linear in rows)",
               "value_on_sample": r["qps_sample"], "one_query_one_core_s": r["one_query_one_core_s"]}
        # and the two arms agree on that sample: the reference's kdtree_nearest ids == the GPU engine's on the same rows
        try:
            with B.Engine(D, K, device=local) as es:
                rows_dev = torch.from_numpy(rows).to(dev)
                es.insert_device(rows_dev.data_ptr(), n_sample, D)
                torch.cuda.synchronize()
                ix, _, _ = es.nearest(r["Q"], 1)
                del rows_dev
            cpu["ids_equal_gpu_engine_on_sample"] = bool(np.array_equal(ix[:, 0].astype(np.uint64), r["ids"].astype(np.uint64)))
        except Exception as ex:  # noqa: BLE001
            cpu["ids_equal_gpu_engine_on_sample"] = f"{type(ex).__name__}: {ex}"

    rows_per_rank = idx.hi - idx.lo
    # the roofline is stated on the bytes of the copy of the log the timed launches actually streamed (engine stat
    # scan_plane_last: 3 = one-byte plane, K13, the default; 2 = bf16 hi plane, K12; 1 = hi + lo planes, K11; 0 = the fp64 rows, K1)
    algo_bytes = plane_bytes(plane_used, rows_per_rank, K)
    peak, peak_src = measured_peak_gbs()
    achieved = algo_bytes / (scan_ms_avg / 1e3) / 1e9 if scan_ms_avg > 0 else 0.0
    traffic = profile_traffic()
    tr = (traffic or {}).get({0: "scan_wide", 1: "scan_shadow", 2: "scan_plane", 3: "scan_plane8"}[plane_used])
    qps = args.steps / (dev_ms / 1e3)
    dtype = {0: "f64",
             1: "answers f64 (reference operation order, bit-identical to the reference); scan keys fp32 from the hi + lo bf16 planes",
             2: "answers f64 (reference operation order, bit-identical to the reference); scan keys fp32 from the bf16 hi plane of "
                "the log's split-bf16 shadow, completeness proven per query, unproven queries re-answered from the fp64 rows",
             3: "answers f64 (reference operation order, bit-identical to the reference); scan keys exact integers (u8 x u16 dot "
                "products) from a one-byte plane of the log, completeness proven per query, unproven queries re-answered from the fp64 rows"}[plane_used]
    line = {
        "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": dtype, "data": "synthetic", "config": workload_config(args),
        "e2e": {"value": args.steps / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": D * 8,
                "d2h_bytes_per_step": k * 32, "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": PLANE_KERNEL[plane_used] + (" (variant %d)" % e_variant(args) if plane_used == 0 else ""),
                     "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak,
                     # ncu dram__bytes_read+write of ONE launch of this kernel, captured once (profiles/scan_traffic.json) on
                     # tr["rows"] rows and scaled to this launch's rows (pure streaming: bytes scale with rows) -- a STATIC
                     # figure, not measured in this run
                     "traffic": tr["dram_bytes_per_launch"] * rows_per_rank / tr["rows"] if tr else None,
                     "traffic_kind": "static: one ncu capture scaled by rows" if tr else None,
                     "traffic_source": tr["source"] if tr else None,
                     "algorithmic_bytes_per_launch": algo_bytes, "avg_launch_ms": scan_ms_avg,
                     "avg_launch_ms_source": "CUDA events around the timed region / launches in it (one launch per step)"
                                             if one_launch_per_step else "CUDA events around every launch of K extra steps",
                     "launches_timed": int(launches) if one_launch_per_step else int(scan_launches), "peak_source": peak_src,
                     "isolated_launch_ms": scan_ms_iso,
                     "isolated_launch_note": "same kernel, K extra steps after the timed region with an event on either side of every "
                                             "launch (no overlap between consecutive launches)",
                     "scan_share_of_step": 1.0 if one_launch_per_step else scan_ms_avg / (dev_ms / args.steps),
                     "launch_includes": "scan + (fused tail: merge of the CTA lists, reference-order re-rank, proof"
                                        + (", peer-memory exchange + cross-shard merge)" if world > 1 else ")"),
                     "fp64_rows_equivalent_gbs": rows_per_rank * K * 8 / (scan_ms_avg / 1e3) / 1e9 if scan_ms_avg > 0 else 0.0},
        "cpu_baseline": cpu,
        "clocks": sampler.summary(),
        "parity_check": parity,
        "fp64_scan": fp64_scan,
        "batch": batch,
        "batch_dmma": batch_dmma,
        "store": {"rows_per_rank": rows_per_rank, "build_s": build_s, "hbm_gib_mapped": e.stats()["hbm_bytes_mapped"] / 2**30,
                  "exact_reruns": e.stats()["exact_reruns"], "fp64_reruns": e.stats()["fp64_reruns"],
                  "last_result_seq": int(last["seq"][0, 0]),
                  "e2e_entry_point": "svdb_nearest_batch (C-ABI, host buffers)" if world == 1 else "svdb.sharded.ShardedIndex.nearest"},
    }
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def e_variant(args):
    for kv in args.opt:
        if kv.startswith("scan.variant="):
            return int(kv.split("=")[1])
    return int(os.environ.get("SVDB_SCAN_VARIANT", "0"))


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the real stdout; everything else any library prints to fd 1
    (e.g. NCCL's version banner) has been redirected to stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--kd-dim", type=int, default=0)
    ap.add_argument("--batch-queries", type=int, default=1024)
    ap.add_argument("--cpu-sample-rows", type=int, default=100_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fp64-scan", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="engine option name=value (e.g. scan.variant=1)")
    args = ap.parse_args()
    if args.kd_dim <= 0:
        args.kd_dim = args.dim
    args.warmup = max(3, args.warmup) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
