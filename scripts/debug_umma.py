"""K10 bring-up on a GPU box: one case per process (a device trap in one case must not hide the others).

    python scripts/debug_umma.py <case>      # prints one JSON line

Checks (a) the approximate keys of the first tile against fp64 distances and the error bound finalize uses,
(b) that the answers are bit-identical to the K2 (FP64 DMMA) path and to K1, (c) timing against K2."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "simple-vector-db_b200")]
from svdb import binding as B  # noqa: E402
from svdb import synth  # noqa: E402

CASES = {
    # name: (n, D, K, nq, k, time_it)
    "tiny": (1000, 128, 128, 100, 10, False),
    "one_tile": (128, 64, 64, 70, 1, False),
    "k768": (20000, 768, 768, 300, 10, False),
    "ragged": (30011, 100, 100, 257, 5, False),
    "kd_prefix": (50000, 200, 50, 1024, 24, False),
    "time_1M": (1000000, 768, 768, 1024, 10, True),
    "time_1M_128": (1000000, 128, 128, 1024, 10, True),
}


def main():
    name = sys.argv[1]
    n, D, K, nq, k, time_it = CASES[name]
    rows = synth.uniform_rows(11, n, D)
    Q = synth.uniform_rows(12, nq, D)
    out = {"case": name, "n": n, "D": D, "K": K, "nq": nq, "k": k}
    with B.Engine(D, K) as e:
        e.insert(rows)
        e.flush()
        e.set_option("nearest.umma_min_queries", 0)
        t0 = time.perf_counter()
        ref = e.nearest(Q, k)                       # K2 (DMMA) path
        out["k2_first_call_s"] = time.perf_counter() - t0
        r0 = e.stats()["exact_reruns"]
        e.set_option("nearest.umma_min_queries", 1)
        e.set_option("nearest.umma_min_kd_dim", 1)
        e.set_option("umma.debug_keys", 1)
        l0 = e.stats()["kernels_launched"]
        got = e.nearest(Q, k)                       # K10
        out["k10_launches"] = e.stats()["kernels_launched"] - l0
        out["k10_exact_reruns"] = e.stats()["exact_reruns"] - r0
        bn = 64 if nq <= 64 else (128 if nq <= 128 else 256)
        keys = e.debug_filter_keys(bn)
        nr, ncol = min(128, n), min(bn, nq)
        X, QQ = rows[:nr, :K], Q[:ncol, :K]
        d_true = ((X[:, None, :] - QQ[None, :, :]) ** 2).sum(-1)
        err = np.abs(keys[:nr, :ncol].astype(np.float64) - d_true)
        scale = (rows[:, :K] ** 2).sum(1).max() + (QQ ** 2).sum(1)[None, :]
        coef = 3.2 * 2.0 ** -16 + (3.0 * K / 16.0) * 2.0 ** -21 + 8.0 * 2.0 ** -20
        out["key_err_max"] = float(err.max())
        out["key_err_over_scale_max"] = float((err / scale).max())
        out["bound_coef"] = coef
        out["bound_margin"] = float(coef / max((err / scale).max(), 1e-300))
        out["keys_sample"] = [float(x) for x in keys[0, :3]]
        out["true_sample"] = [float(x) for x in d_true[0, :3]]
        same = all(np.array_equal(a.view(np.uint64) if a.dtype == np.float64 else a, b.view(np.uint64) if b.dtype == np.float64 else b)
                   for a, b in zip(got, ref))
        out["identical_to_k2"] = bool(same)
        if not same:
            bad = np.nonzero((got[2] != ref[2]).any(1))[0]
            out["queries_differing"] = int(len(bad))
            out["first_bad"] = {"q": int(bad[0]), "got_seq": got[2][bad[0]].tolist(), "want_seq": ref[2][bad[0]].tolist()} if len(bad) else None
        if time_it:
            import torch
            for label, mq in (("k10", 1), ("k2", 0)):
                e.set_option("nearest.umma_min_queries", mq)
                e.nearest(Q, k)
                torch.cuda.synchronize()
                reps = 3 if label == "k10" else 1
                t0 = time.perf_counter()
                for _ in range(reps):
                    e.nearest(Q, k)
                torch.cuda.synchronize()
                out[f"{label}_ms_per_call_host"] = (time.perf_counter() - t0) / reps * 1e3
            out["speedup_vs_k2"] = out["k2_ms_per_call_host"] / out["k10_ms_per_call_host"]
            # the filter kernel alone (CUDA events around its launch on the engine's stream)
            e.set_option("nearest.umma_min_queries", 1)
            e.set_option("profile.scan_events", 1)
            e.take_scan_time()
            for _ in range(3):
                e.nearest(Q, k)
            ms, launches = e.take_scan_time()
            e.set_option("profile.scan_events", 0)
            out["k10_filter_kernel_ms"] = ms / max(1, launches)
            kp = -(-K // 64) * 64
            out["k10_filter_executed_tflops"] = 3 * 2.0 * nq * n * kp / (ms / max(1, launches) / 1e3) / 1e12
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
