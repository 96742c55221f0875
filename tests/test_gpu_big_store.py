"""Parity on a store of more than 2^32 elements (6M x 768 fp64 = 4.6e9 elements, 37 GB), through every scan family:
K1 (fp64 rows), K11 (hi + lo bf16 planes), K12 (bf16 hi plane), K13 (one-byte plane; single and paired queries), K2 (FP64
DMMA) and K10 (tcgen05).  32-bit row or byte offsets, tile counters or tensor-map extents would show up here and nowhere in
the small-size tests.  The CPU oracle cannot scan 37 GB, so the check is oracle/bigcheck.py's: an independent chunked torch
fp64 brute force over rows REGENERATED from their seeds keeps 64 candidates per query, the oracle (orc_sqdist, the
reference's operation order) re-ranks them, ids and fp64 distance bits must be equal.  Needs ~75 GB of free HBM."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

sys.path.insert(0, ROOT)
from oracle import bigcheck  # noqa: E402
from svdb import binding as B  # noqa: E402

N, D, CHUNK, SEED = 6_000_000, 768, 250_000, 4242


def chunks(dev):
    for c in range(N // CHUNK):
        g = torch.Generator(device=dev).manual_seed(SEED * 1_000_003 + c)
        yield c * CHUNK, torch.rand((CHUNK, D), dtype=torch.float64, device=dev, generator=g)


def test_every_scan_family_on_a_store_beyond_2_to_32_elements():
    dev = torch.device("cuda:0")
    free, _ = torch.cuda.mem_get_info(dev)
    if free < 80 * 2 ** 30:
        pytest.skip(f"needs ~75 GB of free HBM, {free / 2 ** 30:.0f} GiB available")
    assert N * D > 2 ** 32
    nq, k = 8, 10
    with B.Engine(D, D, device=0) as e:
        for first, rows in chunks(dev):
            torch.cuda.synchronize()
            e.insert_device(rows.data_ptr(), rows.shape[0], D)
            del rows
        Q = torch.rand((nq, D), dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(SEED + 1))
        cand = bigcheck.brute_candidates(chunks(dev), Q, 64)
        Qn = Q.cpu().numpy()
        torch.cuda.empty_cache()

        def check(what, k_call, calls):
            """calls: list of query index arrays, one host call each"""
            ids = np.zeros((nq, k_call), dtype=np.uint64)
            d = np.zeros((nq, k_call))
            for sel in calls:
                i, dist, _ = e.nearest(Qn[sel], k_call)
                ids[sel], d[sel] = i, dist
            asked = np.unique(np.concatenate(calls))
            v = bigcheck.verdict([tuple(a[asked] for a in cand)], Qn[asked], ids[asked], d[asked], k_call, N)
            assert v["ok"], (what, v.get("mismatches"))
            assert v["brute_force_margin_rel"] > 1e-6
            return v

        single = [np.array([i]) for i in range(nq)]
        pairs = [np.array([i, i + 1]) for i in range(0, nq, 2)]
        e.set_option("nearest.umma_min_queries", 0)
        e.set_option("nearest.mma_min_queries", 0)
        for plane, name in ((0, "K1 fp64 rows"), (1, "K11 hi + lo planes"), (2, "K12 hi plane"), (3, "K13 byte plane")):
            e.set_option("scan.plane", plane)
            before = e.stats()["fp64_reruns"]
            check(name + ", top-1", 1, single[:4])
            check(name + ", top-10", k, single[4:] + single[:4])
            assert e.stats()["scan_plane_last"] == plane
            assert e.stats()["fp64_reruns"] - before <= 2
        check("K13, two queries per pass", 3, pairs)
        e.set_option("scan.plane", 0)
        e.set_option("nearest.mma_min_queries", 4)
        check("K2 (FP64 DMMA), 8 queries", k, [np.arange(nq)])
        e.set_option("nearest.mma_min_queries", 0)
        e.set_option("nearest.umma_min_queries", 3)
        v = check("K10 (tcgen05), 8 queries", k, [np.arange(nq)])
        assert e.stats()["exact_reruns"] == 0
        if "reference_kdtree_nearest_top1_agrees" in v:           # the compiled reference's own kdtree_nearest on the 64 candidates
            assert v["reference_kdtree_nearest_top1_agrees"] == nq
