"""GPU parity of K10 (csrc/umma_filter.cu): batched /nearest on the tcgen05 tensor cores with split-bf16 keys.
The keys are approximate by design; the answers, after finalize's reference-order re-rank (kdtree.c:134-137),
must be bit-identical to the oracle's -- and the measured key error must stay inside the bound the proof uses."""
import os

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

from svdb import binding as B  # noqa: E402
from svdb import synth  # noqa: E402
from test_gpu_parity import assert_topk_equal, oracle_topk  # noqa: E402


def umma_on(e):
    e.set_option("nearest.umma_min_queries", 1)
    e.set_option("nearest.umma_min_kd_dim", 1)


@pytest.mark.parametrize("n,D,K,nq,k,seed", [
    (20000, 128, 128, 100, 10, 1),     # one group of 128 queries, 28 of them padding
    (9000, 768, 768, 64, 10, 2),       # config-3 rows, group of 64
    (6000, 100, 100, 300, 5, 3),       # K not a multiple of the 64-coordinate stage (zero-filled tail); two groups of 256
    (5000, 200, 50, 33, 24, 4),        # compact kd array (K < D), k = SVDB_MAX_K
    (100, 40, 40, 16, 3, 5),           # fewer rows than one 128-row tile
    (40000, 96, 96, 1024, 10, 6),      # config-3 batch shape: four groups of 256
    (129, 64, 64, 5, 1, 7),            # one full tile + one row
])
def test_umma_path_vs_oracle(port, n, D, K, nq, k, seed):
    rows = synth.uniform_rows(seed, n, D)
    Q = synth.uniform_rows(seed + 70, nq, D)
    want = oracle_topk(port, rows, K, Q, k)
    with B.Engine(D, K) as e:
        e.insert(rows)
        e.flush()
        umma_on(e)
        l0 = e.stats()["kernels_launched"]
        assert_topk_equal(e.nearest(Q, k), want, k)
        assert e.stats()["exact_reruns"] == 0
        assert e.stats()["kernels_launched"] - l0 == 6          # shadow split (hi plane, lo plane) + (prep, query split, filter, finalize)
        l0 = e.stats()["kernels_launched"]
        assert_topk_equal(e.nearest(Q, k), want, k)             # shadow is kept
        assert e.stats()["kernels_launched"] - l0 == 4


def test_umma_key_error_is_inside_the_bound():
    n, K, nq = 4096, 768, 256
    rows = synth.uniform_rows(21, n, K)
    Q = synth.uniform_rows(22, nq, K)
    with B.Engine(K, K) as e:
        e.insert(rows)
        umma_on(e)
        e.set_option("umma.debug_keys", 1)
        e.nearest(Q, 1)
        keys = e.debug_filter_keys(256).astype(np.float64)
    d = ((rows[:128, None, :] - Q[None, :, :]) ** 2).sum(-1)
    scale = (rows ** 2).sum(1).max() + (Q ** 2).sum(1)[None, :]
    coef = 3.2 * 2.0 ** -16 + (3.0 * K / 16.0) * 2.0 ** -21 + 8.0 * 2.0 ** -20
    rel = np.abs(keys - d) / scale
    assert rel.max() < coef / 4, (rel.max(), coef)              # the proof's bound with a margin of at least 4


@pytest.mark.parametrize("K", [64, 256, 768, 1024])
def test_tcgen05_accumulator_loss_is_inside_the_budget(K):
    """What the fp32 accumulation inside and between the chained tcgen05.mma instructions loses, MEASURED on
    truncation-adversarial tiles of +-powers of two (scripts/umma_accumulator_probe.py: no lo plane, exact products, the
    kernel's own keys dumped through umma.debug_keys and compared with exact arithmetic): at most 1 ulp of the running sum
    (2 * 2^-24 sum|x_i q_i|) per instruction -- umma_eabs_coef budgets 2^-21 = four times that -- and the whole key error
    at most a quarter of coef * (max|x|^2 + |q|^2)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from umma_accumulator_probe import probe
    r = probe(K)
    assert r["in_units_of_2^-24_per_mma"] <= 2.0, r
    coef = 3.2 * 2.0 ** -16 + (3.0 * K / 16.0) * 2.0 ** -21 + 8.0 * 2.0 ** -20
    assert r["key_err_over_scale_max"] <= coef / 4, (r["key_err_over_scale_max"], coef)


def test_umma_shadow_follows_inserts_and_updates(port):
    D = 64
    rows = synth.uniform_rows(31, 3000, D)
    more = synth.uniform_rows(32, 500, D)
    Q = synth.uniform_rows(33, 80, D)
    with B.Engine(D, D) as e:
        e.insert(rows)
        umma_on(e)
        assert_topk_equal(e.nearest(Q, 5), oracle_topk(port, rows, D, Q, 5), 5)
        e.insert(more)
        allrows = np.vstack([rows, more])
        assert_topk_equal(e.nearest(Q, 5), oracle_topk(port, allrows, D, Q, 5), 5)
        # a query that IS a freshly inserted row must find it at distance 0
        idx, dist, _ = e.nearest(np.vstack([more[:70], rows[:10]]), 1)
        assert np.array_equal(idx[:70, 0], 3000 + np.arange(70)) and np.all(dist[:, 0] == 0.0)


def test_umma_cancellation_and_huge_values_fall_back(port):
    """Rows far from the origin (the GEMM form cancels) and rows beyond fp32 range: the proof fails or the scale check
    trips, the exact scan answers, results still identical to the oracle."""
    rng = np.random.Generator(np.random.PCG64(9))
    rows = 1.0e6 + rng.random((4000, 64))
    Q = 1.0e6 + rng.random((70, 64))
    with B.Engine(64, 64) as e:
        e.insert(rows)
        umma_on(e)
        assert_topk_equal(e.nearest(Q, 5), oracle_topk(port, rows, 64, Q, 5), 5)
        assert e.stats()["exact_reruns"] + e.stats()["fp64_reruns"] > 0      # low-precision keys -> K1 (fp64 rows) -> exact
    rows = rng.random((2000, 64)) * 1.0e25
    Q = rng.random((70, 64)) * 1.0e25
    with B.Engine(64, 64) as e:
        e.insert(rows)
        umma_on(e)
        assert_topk_equal(e.nearest(Q, 3), oracle_topk(port, rows, 64, Q, 3), 3)
        assert e.stats()["exact_reruns"] + e.stats()["fp64_reruns"] > 0      # low-precision keys -> K1 (fp64 rows) -> exact
    rows = rng.random((2000, 64)) * 1.0e-20
    Q = rng.random((70, 64)) * 1.0e-20
    with B.Engine(64, 64) as e:
        e.insert(rows)
        umma_on(e)
        assert_topk_equal(e.nearest(Q, 3), oracle_topk(port, rows, 64, Q, 3), 3)


def test_umma_mass_duplicates(port):
    """More identical rows than a candidate list holds: completeness cannot be proven from approximate keys; the
    escalation chain must still return the lowest sequence numbers."""
    rng = np.random.Generator(np.random.PCG64(5))
    base = rng.random((1, 48))
    rows = np.vstack([np.repeat(base, 200, axis=0), rng.random((3000, 48))])
    Q = np.vstack([base + 1e-9, rng.random((69, 48))])
    with B.Engine(48, 48) as e:
        e.insert(rows)
        umma_on(e)
        assert_topk_equal(e.nearest(Q, 10), oracle_topk(port, rows, 48, Q, 10), 10)


def test_umma_on_row_shards_and_merge(port):
    """Two row-range SHARD engines on one GPU take a 200-query batch through K10 each; K7's merge of their candidates
    equals one engine over all rows (what bench.py's batch does per rank at N > 1)."""
    n, D, k, nq = 8000, 96, 10, 200
    rows = synth.uniform_rows(41, n, D)
    Q = synth.uniform_rows(42, nq, D)
    want = oracle_topk(port, rows, D, Q, k)
    half = n // 2
    dev_rows = torch.from_numpy(rows).cuda()
    dq = torch.from_numpy(Q).cuda()
    shards = [B.Engine(D, D, seq_base=0, flags=B.FLAG_SHARD), B.Engine(D, D, seq_base=half, flags=B.FLAG_SHARD)]
    stream = torch.cuda.current_stream().cuda_stream
    gathered = torch.zeros((2, nq, k, 4), dtype=torch.int64, device="cuda")
    for s, e in enumerate(shards):
        e.set_stream(stream)
        umma_on(e)
        e.insert_device(dev_rows[s * half:(s + 1) * half].data_ptr(), half, D)
        e.nearest_device(dq.data_ptr(), nq, D, k, gathered[s].data_ptr())
    merged = torch.zeros((nq, k, 4), dtype=torch.int64, device="cuda")
    B.merge_candidates_device(0, stream, gathered.data_ptr(), 2, nq, k, merged.data_ptr())
    torch.cuda.synchronize()
    res = merged.cpu().numpy().view(B.candidate_dtype).reshape(nq, k)
    assert not np.any(res["flags"] & B.CAND_UNSAFE)
    assert_topk_equal((res["seq"], res["dist"], res["seq"]), want, k)
    for e in shards:
        e.close()


@pytest.mark.parametrize("sparse", [1, 0])
def test_umma_rows_in_order_of_decreasing_distance(port, sparse):
    """The order of rows that defeats every running threshold: each tile is closer to the queries than all tiles before it, so
    every key passes.  With the bookkeeping between the CTA barriers running only every fourth tile (umma.sparse_checks) the
    append buffers overflow between two checks; the filter must then say so -- the query comes back unprovable and is
    re-answered by the next rung -- and never lose a neighbour silently.  With a check after every tile nothing overflows."""
    rng = np.random.Generator(np.random.PCG64(77))
    n, K, nq, k = 148 * 128 * 13 + 5, 64, 48, 5
    c = rng.random(K)
    rows = rng.random((n, K))
    rows = rows[np.argsort(-((rows - c) ** 2).sum(1), kind="stable")]
    Q = c + 1e-3 * rng.standard_normal((nq, K))
    want = oracle_topk(port, rows, K, Q, k)
    with B.Engine(K, K) as e:
        e.insert(rows)
        umma_on(e)
        e.set_option("umma.sparse_checks", sparse)
        assert_topk_equal(e.nearest(Q, k), want, k)
        st = e.stats()
        if sparse:
            assert st["fp64_reruns"] + st["exact_reruns"] > 0       # the overflow was noticed, not papered over
        else:
            assert st["fp64_reruns"] + st["exact_reruns"] == 0
