"""GPU parity of K11 (scan_shadow_kernel, csrc/scan_kernels.cu: hi + lo bf16 planes), K12 (scan_plane_kernel,
csrc/plane_scan.cu: the hi plane alone) and K13 (scan_plane8_kernel: the one-byte plane, exact integer keys): the single-query / small-batch scans over the split-bf16 shadow of
the log instead of the fp64 rows (option scan.plane).  Approximate fp32 keys, answers after the reference-order re-rank
(kdtree.c:134-137) bit-identical to the oracle's; with and without the fused tail (option scan.fuse_tail)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

from svdb import binding as B  # noqa: E402
from svdb import synth  # noqa: E402
from test_gpu_parity import assert_topk_equal, oracle_topk  # noqa: E402


@pytest.mark.parametrize("plane,fuse", [(1, 1), (2, 1), (3, 1), (1, 0), (2, 0), (3, 0)])
@pytest.mark.parametrize("n,D,K,nq,k,seed", [
    (20000, 128, 128, 1, 1, 1),        # the headline shape in small: one query, top-1
    (9000, 768, 768, 3, 10, 2),        # config-3 rows; passes of 2 + 1 queries
    (9000, 768, 768, 2, 4, 12),        # config-3 rows, what K13 serves: one or two queries, k <= 4
    (6000, 100, 100, 2, 5, 3),         # K not a multiple of 64 (zero-filled tail), half-empty last trip
    (5000, 200, 50, 1, 24, 4),         # compact kd array (K < D), k = SVDB_MAX_K
    (37, 40, 40, 1, 3, 5),             # fewer rows than one tile
    (7000, 320, 320, 3, 10, 6),        # two trips, the second one partial
    (8000, 64, 64, 1, 10, 7),          # K12: four rows packed to a warp step
    (8003, 64, 64, 2, 3, 17),          # K13: four rows packed to a warp step, ragged last tile of 128 rows
    (70001, 128, 128, 1, 4, 18),       # K13: two rows packed, several 64-row tiles per warp
    (9001, 500, 500, 2, 3, 19),        # K13: a pair of queries per pass, two trips
    (5003, 1024, 1024, 2, 1, 20),      # K13: a pair, four trips, 4-row tiles
    (9002, 256, 256, 2, 16, 21),       # K13: a pair, one trip, the largest k it serves
    (8001, 192, 192, 2, 2, 8),         # K12: one trip, 24 of 32 lanes
    (3000, 1000, 1000, 1, 10, 9),      # K12: four trips
    (2000, 1100, 1100, 1, 5, 10),      # beyond K12's register-resident query: K11 serves plane 2 as well
])
def test_shadow_scan_vs_oracle(port, n, D, K, nq, k, seed, plane, fuse):
    rows = synth.uniform_rows(seed, n, D)
    Q = synth.uniform_rows(seed + 70, nq, D)
    want = oracle_topk(port, rows, K, Q, k)
    with B.Engine(D, K) as e:
        e.insert(rows)
        e.flush()
        e.set_option("scan.plane", 0)
        e.set_option("scan.fuse_tail", fuse)
        e.set_option("nearest.umma_min_queries", 0)             # 3 queries at kd_dim >= 256 would go to K10 (tests/test_gpu_umma.py)
        assert_topk_equal(e.nearest(Q, k), want, k)             # K1 first: the next call of this shape is the one that gets captured
        e.set_option("scan.plane", plane)
        e.set_option("nearest.umma_min_kd_dim", 1)
        for _ in range(3):                                      # the shadow is built before the capture, never inside it
            assert_topk_equal(e.nearest(Q, k), want, k)
        st = e.stats()
        assert st["exact_reruns"] == 0 and (plane >= 2 or st["fp64_reruns"] == 0)
        if plane == 3 and K <= 1024 and nq <= 2 and k <= 16:
            assert st["scan_plane_last"] == 3            # K13 really ran (kd_dim it supports, one or two queries)
        e.set_option("scan.plane", 0)
        assert_topk_equal(e.nearest(Q, k), want, k)


@pytest.mark.parametrize("kind", ["uniform", "normal", "offset", "heavy_tail", "constant_columns"])
@pytest.mark.parametrize("K", [64, 128, 192, 256, 768, 1000])       # 64 / 128: four / two rows packed to a warp step
def test_byte_plane_scan_on_distributions(port, kind, K):
    """K13 on data a store-wide uniform grid resolves well and badly: answers identical to the oracle either way (what the
    plane cannot prove is re-answered from the fp64 rows); on heavy tails the engine stops using the plane."""
    rng = np.random.Generator(np.random.PCG64(K + len(kind)))
    n = 6000
    if kind == "uniform":
        rows, Q = rng.random((n, K)), rng.random((12, K))
    elif kind == "normal":
        rows, Q = rng.standard_normal((n, K)), rng.standard_normal((12, K))
    elif kind == "offset":
        rows, Q = 1000.0 + rng.random((n, K)), 1000.0 + rng.random((12, K))
    elif kind == "heavy_tail":
        rows, Q = rng.standard_cauchy((n, K)).clip(-1e6, 1e6), rng.standard_cauchy((12, K)).clip(-1e6, 1e6)
    else:
        rows, Q = rng.random((n, K)), rng.random((12, K))
        rows[:, ::3] = 0.5
        Q[:, ::3] = 0.5
    want = oracle_topk(port, rows, K, Q, 3)
    with B.Engine(K, K) as e:
        e.insert(rows)
        e.set_option("scan.plane", 3)
        e.set_option("nearest.umma_min_queries", 0)
        for i in range(12):
            assert_topk_equal(e.nearest(Q[i:i + 1], 3), want[i:i + 1], 3)
        st = e.stats()
        assert st["exact_reruns"] == 0
        if kind in ("uniform", "offset", "constant_columns"):
            assert st["scan_plane_last"] == 3 and st["fp64_reruns"] <= 2
        # queries far outside the grid: the query's own quantisation error makes the proof fail, K1 answers
        far = Q[:2] * 50.0 + 7.0
        assert_topk_equal(e.nearest(far[:1], 3), oracle_topk(port, rows, K, far[:1], 3), 3)
        # rows appended after the grid was fixed, outside its range: clamped, measured, still exact answers
        more = rows[:40] * 3.0 + 2.0
        e.insert(more)
        allrows = np.vstack([rows, more])
        for q in (Q[0:1], more[7:8]):
            assert_topk_equal(e.nearest(q, 3), oracle_topk(port, allrows, K, q, 3), 3)


@pytest.mark.parametrize("plane", [1, 2, 3])
def test_shadow_scan_follows_inserts_and_extremes(port, plane):
    D = 64
    rng = np.random.Generator(np.random.PCG64(7))
    rows = synth.uniform_rows(31, 3000, D)
    more = synth.uniform_rows(32, 300, D)
    with B.Engine(D, D) as e:
        e.insert(rows)
        e.set_option("scan.plane", plane)
        e.set_option("nearest.umma_min_queries", 0)
        Q = synth.uniform_rows(33, 2, D)
        assert_topk_equal(e.nearest(Q, 5), oracle_topk(port, rows, D, Q, 5), 5)
        e.insert(more)
        allrows = np.vstack([rows, more])
        assert_topk_equal(e.nearest(Q, 5), oracle_topk(port, allrows, D, Q, 5), 5)
        idx, dist, _ = e.nearest(more[7:8], 1)                  # a freshly inserted row finds itself
        assert idx[0, 0] == 3007 and dist[0, 0] == 0.0
    # far from the origin the fp32 keys cancel: the proof fails and the exact scan answers
    rows = 1.0e6 + rng.random((4000, D))
    Q = 1.0e6 + rng.random((2, D))
    with B.Engine(D, D) as e:
        e.insert(rows)
        e.set_option("scan.plane", plane)
        assert_topk_equal(e.nearest(Q, 5), oracle_topk(port, rows, D, Q, 5), 5)
        if plane == 3:      # the byte plane's grid starts at the data's minimum: an offset costs it nothing
            assert e.stats()["scan_plane_last"] == 3 and e.stats()["exact_reruns"] + e.stats()["fp64_reruns"] == 0
        else:               # low-precision float keys -> K1 (fp64 rows) -> exact
            assert e.stats()["exact_reruns"] + e.stats()["fp64_reruns"] > 0


@pytest.mark.parametrize("plane", [2, 3])
@pytest.mark.parametrize("K,cluster,k,exact_dups", [
    (256, 60, 10, 0),        # a cluster of near neighbours the plane cannot tell apart: 33..128 candidates, one thread each
    (128, 100, 16, 0),       # packed rows, almost all slots
    (768, 45, 4, 0),
    (256, 70, 5, 40),        # 40 exact duplicates at the minimum among 70 candidates: more ties than warp 0 has lanes
    (64, 300, 3, 0),         # more candidates than slots: the window overflows, the fp64 rows answer
])
def test_plane_tail_with_many_candidates(port, plane, K, cluster, k, exact_dups):
    """The selection path of the tail (tail.cuh) with more candidates than one warp holds: every candidate is re-ranked
    by a thread of its own, the best 32 by exact distance go to warp 0.  Rows of a tight cluster differ by less than the
    plane resolves, so all of them land inside the window; the answers must be the oracle's whatever path that takes."""
    rng = np.random.Generator(np.random.PCG64(K + cluster))
    n = 5000
    rows = rng.random((n, K))
    q = rng.random(K)
    centre = q + 0.02 * rng.standard_normal(K)
    where = rng.choice(n, size=cluster, replace=False)
    rows[where] = centre + 1e-7 * rng.standard_normal((cluster, K))      # far below a step of either plane
    if exact_dups:
        rows[np.sort(where)[:exact_dups]] = centre                        # identical kd-points: the earliest insert wins
    want = oracle_topk(port, rows, K, [q], k)
    with B.Engine(K, K) as e:
        e.insert(rows)
        e.set_option("scan.plane", plane)
        e.set_option("nearest.umma_min_kd_dim", 1)
        for _ in range(2):
            assert_topk_equal(e.nearest(q, k), want, k)
        st = e.stats()
        if cluster <= 100 and not exact_dups:
            assert st["scan_plane_last"] == plane and st["fp64_reruns"] == 0 and st["exact_reruns"] == 0
        if cluster > 128:
            assert st["fp64_reruns"] >= 1


def test_few_queries_per_call_take_byte_plane_passes(port):
    """Up to scan.plane8_max_queries (4) queries per call are answered by K13 passes, two queries to a pass -- an eighth of
    the bytes K2 would stream, a quarter of K10's -- unless the caller set the batch thresholds itself.  (Stores below 4e7
    elements: up to 2; there a pass is all launch latency and one K2 / K10 call is cheaper than three.)"""
    rows = synth.uniform_rows(41, 170000, 256)
    Q = synth.uniform_rows(42, 7, 256)
    want = oracle_topk(port, rows, 256, Q, 3)

    def launches(e, q, w):
        """scan launches of one call: what it launched minus one per query that had to be re-answered from the fp64 rows"""
        s0 = e.stats()
        assert_topk_equal(e.nearest(q, 3), w, 3)
        s1 = e.stats()
        return (s1["kernels_launched"] - s0["kernels_launched"]) - (s1["fp64_reruns"] - s0["fp64_reruns"]), s1

    with B.Engine(256, 256) as e:
        e.insert(rows)
        for nq in (3, 4):
            assert_topk_equal(e.nearest(Q[:nq], 3), want[:nq], 3)
            assert e.stats()["scan_plane_last"] == 3
        n, st = launches(e, Q[:4], want[:4])
        assert n == 2, st                                             # two queries share a pass: two fused launches, nothing else
        n, st = launches(e, Q[:3], want[:3])
        assert n == 2, st                                             # a pair and a single
        e.set_option("scan.plane8_pair", 0)
        n, st = launches(e, Q[:4], want[:4])
        assert n == 4, st                                             # A/B switch: one launch per query
        e.set_option("scan.plane8_pair", 1)
        assert_topk_equal(e.nearest(Q, 3), want, 3)                   # 7 queries: the tensor-core path
    small = synth.uniform_rows(43, 20000, 256)
    want_small = oracle_topk(port, small, 256, Q, 3)
    with B.Engine(256, 256) as e:
        e.insert(small)
        assert_topk_equal(e.nearest(Q[:2], 3), want_small[:2], 3)
        assert e.stats()["scan_plane_last"] == 3
        n, st = launches(e, Q[:3], want_small[:3])                    # three queries on a small store: one K10 call
        assert n > 3, st                                              # query prep + planes + filter + finalize
        e.set_option("nearest.umma_min_queries", 0)
        e.set_option("nearest.mma_min_queries", 3)                    # explicit threshold: taken literally
        n, st = launches(e, Q[:4], want_small[:4])
        assert n == 3, st                                             # K2: prep + DMMA scan + finalize, not K13 passes


def test_back_to_back_device_calls_overlap_safely(port):
    """Device-resident single-query calls back to back: from the second one on the K13 scan is launched with programmatic
    stream serialization and may start while the tail of the call before it still runs (option scan.overlap_steps).
    Answers must be what the oracle says for the query the caller passed -- also when the caller rewrites ONE query buffer
    between the calls with a kernel of its own on the same stream, the case the scan's re-check of the query exists for:
    an answer is then either right or flagged SVDB_CAND_UNSAFE, never silently the previous query's."""
    import torch
    n, K, k, calls = 60000, 256, 3, 40
    rows = synth.uniform_rows(51, n, K)
    Q = synth.uniform_rows(52, calls, K)
    want = oracle_topk(port, rows, K, Q, k)
    dev = torch.device("cuda:0")
    with B.Engine(K, K) as e:
        e.insert(rows)
        e.set_stream(torch.cuda.current_stream().cuda_stream)
        q_all = torch.from_numpy(Q).to(dev)
        outs = torch.zeros((calls, k, 4), dtype=torch.int64, device=dev)
        for overlap in (1, 0):
            e.set_option("scan.overlap_steps", overlap)
            # (a) every call has its own query and answer buffers; nothing else is enqueued in between
            outs.zero_()
            for rep in range(2):
                for i in range(calls):
                    e.nearest_device(q_all[i].data_ptr(), 1, K, k, outs[i].data_ptr())
            torch.cuda.synchronize()
            got = outs.cpu().numpy().view(np.uint64)
            unproven = 0
            for i in range(calls):
                wseq, widx, wd = want[i]
                if (got[i, :, 3] & B.CAND_UNSAFE).any():       # the device entry point hands an unproven answer to its caller
                    unproven += 1
                    continue
                np.testing.assert_array_equal(got[i, :, 1].astype(np.int64), wseq)        # svdb_candidate: dist, seq, index, flags
                np.testing.assert_array_equal(got[i, :, 0], wd.view(np.uint64))
            assert unproven <= 2, unproven                      # nothing changes the queries here: no scan may call its query stale
            # (b) one query buffer, rewritten by a torch kernel right in front of every call
            qbuf = torch.zeros(K, dtype=torch.float64, device=dev)
            outs.zero_()
            for i in range(calls):
                qbuf.copy_(q_all[i])
                e.nearest_device(qbuf.data_ptr(), 1, K, k, outs[i].data_ptr())
            torch.cuda.synchronize()
            got = outs.cpu().numpy().view(np.uint64)
            flagged = 0
            for i in range(calls):
                wseq, widx, wd = want[i]
                if (got[i, :, 3] & B.CAND_UNSAFE).any():
                    flagged += 1
                    continue
                np.testing.assert_array_equal(got[i, :, 1].astype(np.int64), wseq)
                np.testing.assert_array_equal(got[i, :, 0], wd.view(np.uint64))
            assert flagged <= calls // 4
        assert e.stats()["scan_plane_last"] == 3
