"""GPU parity: the CUDA path, called through the C-ABI, against the oracle and the golden
vectors the reference produced.  Bit-exact for ids, sequence numbers, distances (fp64 bits)
and metrics (fp32 bits)."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

from oracle import binding as OB  # noqa: E402
from oracle.binding import PortDB  # noqa: E402
from svdb import binding as B  # noqa: E402
from svdb import synth  # noqa: E402

NEAREST_CASES = ["nearest_script_k3", "nearest_coarse_k3", "nearest_uniform_k32",
                 "nearest_normal_k128", "nearest_uniform_k5_d12"]


def oracle_topk(port, rows, K, Q, k):
    db = PortDB(port, rows.shape[1], K)
    for r in rows:
        db.insert(r)
    out = [db.topk(q, k) for q in Q]
    db.close()
    return out


def assert_topk_equal(got, want, k):
    idx, dist, seq = got
    for i, (wseq, widx, wd) in enumerate(want):
        m = len(wseq)
        np.testing.assert_array_equal(seq[i, :m].astype(np.int64), wseq, err_msg=f"query {i}: seq")
        np.testing.assert_array_equal(idx[i, :m], widx, err_msg=f"query {i}: index")
        np.testing.assert_array_equal(dist[i, :m].view(np.uint64), wd.view(np.uint64), err_msg=f"query {i}: dist bits")
        assert np.all(idx[i, m:] == B.NONE) and np.all(np.isinf(dist[i, m:]))


# ---- golden vectors from the reference ---------------------------------------------------

def oracle_tree_ids(port, rows, K, Q):
    h = port.build(rows, K)
    ids = port.nearest_batch(h, Q)
    port.free(h)
    return ids


@pytest.mark.parametrize("name", NEAREST_CASES)
def test_nearest_matches_reference_golden(port, name):
    """Identical ids, distinct-point ties included (nearest_coarse_k3 is full of them)."""
    g = load_golden(name)
    rows, K, Q, want = g["rows"], int(g["K"]), g["queries"], g["ids"]
    flat = oracle_topk(port, rows, K, Q, 2)
    with B.Engine(rows.shape[1], K) as e:
        e.insert(rows)
        idx1, dist1, _ = e.nearest(Q, 1)            # thin K: tree traversal (K6); wide K: scan + re-rank
        idx2, dist2, _ = e.nearest(Q, 2)            # scan + finalize, ties resolved against the tree
        e.set_option("nearest.tree_max_k", 0)       # force the scan path for k = 1 as well
        idx1s, dist1s, _ = e.nearest(Q, 1)
    np.testing.assert_array_equal(idx1[:, 0], want)
    np.testing.assert_array_equal(idx1s[:, 0], want)
    np.testing.assert_array_equal(idx2[:, 0], want)
    for i, (wseq, widx, wd) in enumerate(flat):
        assert dist1[i, 0].view(np.uint64) == wd[0].view(np.uint64) == dist1s[i, 0].view(np.uint64)
        np.testing.assert_array_equal(dist2[i].view(np.uint64), wd.view(np.uint64))


def test_metrics_match_reference_golden_bits():
    g = load_golden("metrics")
    off = 0
    for n, want in zip(g["lens"], g["res"]):
        a, b = g["a"][off:off + n], g["b"][off:off + n]
        off += n
        got = np.array([B.compare_vectors(m, a, b) for m in range(3)], dtype=np.float32)
        for m in range(3):
            if np.isnan(want[m]):
                assert np.isnan(got[m])
            else:
                assert got[m].view(np.uint32) == want[m].view(np.uint32), (n, m, got, want)


def test_delta_stream_matches_reference_golden(port):
    """insert / update / delete stream on coarse-grid values: stale entries, index drift and
    plenty of exact ties; the id after every op is the reference's."""
    g = load_golden("delta_ops")
    D, K = int(g["D"]), int(g["K"])
    model = PortDB(port, D, K)
    with B.Engine(D, K) as e, B.Engine(D, K) as e_scan:
        e_scan.set_option("nearest.tree_max_k", 0)
        for i, ((code, j), v, q, want, size) in enumerate(zip(g["ops"], g["vals"], g["queries"], g["ids"], g["sizes"])):
            for eng in (e, e_scan):
                if code == 0:
                    assert eng.insert(v) == j
                elif code == 1:
                    eng.update(int(j), v)
                else:
                    eng.delete(int(j))
                assert eng.size == size
            (model.insert(v) if code == 0 else model.update(int(j), v) if code == 1 else model.delete(int(j)))
            assert e.nearest(q, 1)[0][0, 0] == want, f"op {i} (tree traversal)"
            assert e_scan.nearest(q, 1)[0][0, 0] == want, f"op {i} (scan + tie resolution)"
            k = min(3, e.log_size)
            idx, dist, seq = e.nearest(q, k)
            assert idx[0, 0] == want, f"op {i} (top-{k})"
            np.testing.assert_array_equal(dist[0].view(np.uint64), model.topk(q, k)[2].view(np.uint64))
    model.close()


def test_ties_golden(port):
    g = load_golden("ties")
    ro = qo = 0
    for ci, ((n, D, K), want) in enumerate(zip(g["meta"], g["ids"])):
        rows = g["rows"][ro:ro + n * D].reshape(n, D)
        q = g["queries"][qo:qo + D]
        ro += n * D
        qo += D
        with B.Engine(int(D), int(K)) as e:
            e.insert(rows)
            assert e.nearest(q, 1)[0][0, 0] == want, f"tie case {ci} (tree)"
            assert e.nearest(q, min(2, n))[0][0, 0] == want, f"tie case {ci} (scan)"


def _lattice_ties(K, n_far, seed):
    """Rows with MANY distinct kd-points at exactly the same distance from the origin
    (signed permutations of (3,4,0..) and (5,0,0..): d = 25 in exact arithmetic), mixed with
    farther rows, in a seeded random insertion order."""
    rng = np.random.Generator(np.random.PCG64(seed))
    tied = []
    for a in range(K):
        for b in range(K):
            if a == b:
                continue
            for sa in (3.0, -3.0):
                for sb in (4.0, -4.0):
                    v = np.zeros(K)
                    v[a], v[b] = sa, sb
                    tied.append(v)
        for s5 in (5.0, -5.0):
            v = np.zeros(K)
            v[a] = s5
            tied.append(v)
    tied = np.array(tied)
    far = rng.integers(-9, 10, size=(n_far, K)).astype(np.float64)
    far = far[(far ** 2).sum(axis=1) > 25]
    rows = np.concatenate([tied, far])
    rng.shuffle(rows)
    return rows


@pytest.mark.parametrize("K,n_far", [(2, 300), (3, 2000), (6, 3000), (24, 4000)])
def test_many_distinct_points_at_equal_distance(port, K, n_far):
    """12 .. 2256 distinct entries tie for the minimum: more than any candidate list holds.
    The answer must still be the one the reference's traversal reaches first, whichever path
    produces it (traversal, scan + resolver, or the escalation EXACT -> TREE)."""
    rows = _lattice_ties(K, n_far, seed=K)
    Q = np.zeros((1, K))
    Q2 = np.concatenate([Q, 1e-3 * synth.normal_rows(K, 5, K)])       # and a few tie-free queries
    want = oracle_tree_ids(port, rows, K, Q2)
    with B.Engine(K, K) as e:
        e.insert(rows)
        np.testing.assert_array_equal(e.nearest(Q2, 1)[0][:, 0], want)
        if K <= 8:                                                     # k-smallest traversal (K6 for k > 1)
            idx, dist, _ = e.nearest(Q2, 4)
            np.testing.assert_array_equal(idx[:, 0], want)
            assert np.all(dist[0] == 25.0)
        e.set_option("nearest.tree_max_k", 0)                          # k = 1 through the scan
        np.testing.assert_array_equal(e.nearest(Q2, 1)[0][:, 0], want)
        st = e.stats()
        assert st["tree_reruns"] >= 1 or K <= 3                        # 12 / 30 ties still fit the lists
        idx, dist, _ = e.nearest(Q2, 4)
        np.testing.assert_array_equal(idx[:, 0], want)
        assert np.all(dist[0] == 25.0)


def test_tree_traversal_equals_scan_on_random_data(port):
    n, D, K = 200_000, 8, 3
    rows = synth.uniform_rows(17, n, D)
    Q = synth.uniform_rows(18, 500, D)
    with B.Engine(D, K) as e:
        e.insert(rows)
        a = e.nearest(Q, 1)
        a10 = e.nearest(Q, 10)                       # k-smallest traversal
        a24 = e.nearest(Q[:50], 24)
        e.set_option("nearest.tree_max_k", 0)
        b = e.nearest(Q, 1)
        b10 = e.nearest(Q, 10)                       # exact scan + finalize
        b24 = e.nearest(Q[:50], 24)
        assert e.stats()["tree_rounds"] > 0
    for x, y in list(zip(a, b)) + list(zip(a10, b10)) + list(zip(a24, b24)):
        np.testing.assert_array_equal(x, y)
    np.testing.assert_array_equal(a[0][:, 0], oracle_tree_ids(port, rows, K, Q))


# ---- the reference's own API served by the CUDA library (drop-in) ---------------------------

@pytest.mark.parametrize("name", ["nearest_script_k3", "nearest_uniform_k32", "nearest_normal_k128"])
def test_dropin_api_nearest_golden(name):
    g = load_golden(name)
    api = OB.RefApi(B.LIB_PATH)
    L = api.lib
    db = L.vector_db_init(0, int(g["K"]))
    for i, r in enumerate(g["rows"]):
        assert L.vector_db_insert(db, api.make_vector(r, uuid=f"row-{i}")) == i
    assert db.contents.size == len(g["rows"])
    got = np.array([api.nearest(db, q) for q in g["queries"][:80]], dtype=np.uint64)
    np.testing.assert_array_equal(got, g["ids"][:80])
    v = L.vector_db_read(db, 7).contents
    assert v.uuid == b"row-7"
    np.testing.assert_array_equal(np.ctypeslib.as_array(v.data, shape=(v.dimension,)), g["rows"][7])
    assert L.vector_db_read_by_uuid(db, b"row-11").contents.dimension == g["rows"].shape[1]
    L.vector_db_free(db)


def test_dropin_api_delta_stream_golden():
    g = load_golden("delta_ops")
    api = OB.RefApi(B.LIB_PATH)
    L = api.lib
    db = L.vector_db_init(0, int(g["K"]))
    for i, ((code, j), v, q, want, size) in enumerate(zip(g["ops"], g["vals"], g["queries"], g["ids"], g["sizes"])):
        if code == 0:
            assert L.vector_db_insert(db, api.make_vector(v, uuid=f"u{i}")) == j
        elif code == 1:
            vec = api.make_vector(v, uuid=f"u{i}")
            L.vector_db_update(db, int(j), vec)
            if j >= size:
                api._libc.free(vec.data)   # the store ignored it: still ours
        else:
            L.vector_db_delete(db, int(j))
        assert db.contents.size == size
        assert api.nearest(db, q) == want, f"op {i}"
    L.vector_db_free(db)


def test_dropin_metrics_and_batch_compare(cpu):
    rows = synth.normal_rows(21, 300, 100)
    api = OB.RefApi(B.LIB_PATH)
    L = api.lib
    db = L.vector_db_init(0, 3)
    for r in rows:
        L.vector_db_insert(db, api.make_vector(r))
    h = cpu.build(rows, 3)
    i1, i2 = synth.index_pairs(5, 64, 300)
    for m in range(3):
        want = cpu.compare_batch(h, m, i1, i2)
        f = (L.cosine_similarity, L.euclidean_distance, L.dot_product)[m]
        for a, b, w in list(zip(i1, i2, want))[:8]:
            got = np.float32(f(L.vector_db_read(db, int(a)).contents, L.vector_db_read(db, int(b)).contents))
            assert got.view(np.uint32) == w.view(np.uint32)
        out = np.empty(64, dtype=np.float32)
        L.vector_db_compare_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        assert L.vector_db_compare_batch(db, m, i1.ctypes.data, i2.ctypes.data, 64, out.ctypes.data) == 0
        np.testing.assert_array_equal(out.view(np.uint32), want.view(np.uint32))
    cpu.free(h)
    L.vector_db_free(db)


def test_dropin_save_load_roundtrip(tmp_path):
    api = OB.RefApi(B.LIB_PATH)
    L = api.lib
    rows = synth.uniform_rows(9, 40, 6)
    db = L.vector_db_init(0, 2)
    for i, r in enumerate(rows):
        L.vector_db_insert(db, api.make_vector(r, uuid=f"id-{i}"))
    path = str(tmp_path / "db.bin").encode()
    L.vector_db_save(db, path)
    assert len(open(path, "rb").read()) == 8 + 40 * (37 + 8 + 6 * 8)
    db2 = L.vector_db_load(path, 2)
    assert db2.contents.size == 40
    for i in (0, 17, 39):
        assert api.nearest(db2, rows[i]) == i
        assert L.vector_db_read(db2, i).contents.uuid == f"id-{i}".encode()
    L.vector_db_free(db)
    L.vector_db_free(db2)


# ---- seeded inputs against the oracle -------------------------------------------------------

@pytest.mark.parametrize("n,D,K,k,seed", [
    (20000, 8, 3, 5, 1),        # thin path, prefix subspace
    (3000, 12, 12, 10, 2),      # thin path at its upper K
    (4099, 20, 13, 10, 21),     # packed rows: 8 lanes per row (stride 14 / 16) ...
    (3000, 16, 16, 10, 22),
    (5000, 17, 17, 3, 3),       # ... 16 lanes per row, odd K (zero padded stride 18) ...
    (2777, 40, 24, 7, 23),
    (6001, 32, 32, 10, 24),     # ... up to 32 doubles
    (3500, 33, 33, 5, 25),      # 8 / 16-row tiles of a whole warp per row
    (3100, 63, 63, 5, 26),
    (2900, 96, 96, 5, 27),
    (12000, 128, 128, 10, 4),   # config-2 shape (K = D = 128), rows alias the kd log
    (2500, 768, 768, 10, 5),    # config-3 row shape
    (3000, 200, 50, 24, 6),     # K < D wide: compact kd array, k = SVDB_MAX_K
    (40, 64, 64, 10, 7),        # fewer rows than CTAs
    (7, 40, 40, 10, 8),         # fewer rows than k
])
def test_topk_vs_oracle(port, n, D, K, k, seed):
    rows = synth.uniform_rows(seed, n, D)
    Q = synth.uniform_rows(seed + 50, 13, D)    # 13: ragged against the 8/4/1 query passes
    want = oracle_topk(port, rows, K, Q, k)
    with B.Engine(D, K) as e:
        e.insert(rows)
        assert_topk_equal(e.nearest(Q, k), want, k)
        st = e.stats()
        assert st["kernels_launched"] > 0 and st["exact_reruns"] == 0
        if K > 12:
            e.set_option("nearest.mma_min_queries", 0)      # 13 queries: K1 in passes of 8 + 4 + 1 from here on
            assert_topk_equal(e.nearest(Q, k), want, k)
            for opts in ({"scan.variant": 1}, {"scan.variant": 0, "scan.nq_per_pass": 1},
                         {"scan.nq_per_pass": 8, "scan.warps": 4, "scan.stages": 2},
                         {"scan.tile_rows": 1, "scan.ctas_per_sm": 2}, {"scan.force_exact": 1}):
                for name, v in opts.items():
                    e.set_option(name, v)
                assert_topk_equal(e.nearest(Q, k), want, k)


@pytest.mark.parametrize("n,D,K,nq,k,seed", [
    (20000, 128, 128, 100, 10, 1),     # 100 queries: two groups of 64, the second ragged
    (9000, 768, 768, 64, 10, 2),       # config-3 rows, exactly one group
    (6000, 100, 100, 40, 5, 3),        # K not a multiple of the 32-coordinate chunk (zero-filled tail)
    (5000, 200, 50, 33, 24, 4),        # compact kd array (K < D), k = SVDB_MAX_K
    (300, 40, 40, 16, 3, 5),           # fewer rows than one 128-row tile per stream; group of 16
    (7000, 96, 96, 24, 10, 6),         # group of 32
    (7000, 64, 64, 5, 1, 7),           # smallest batch that takes the tensor-core path (group of 16, 11 padded)
])
def test_batched_dmma_path_vs_oracle(port, n, D, K, nq, k, seed):
    """K2: >= 4 queries per call go through the FP64 tensor-core GEMM-form scan; after the
    reference-order re-rank the answers are bit-identical to the oracle's."""
    rows = synth.uniform_rows(seed, n, D)
    Q = synth.uniform_rows(seed + 70, nq, D)
    want = oracle_topk(port, rows, K, Q, k)
    with B.Engine(D, K) as e:
        e.insert(rows)
        e.flush()
        e.set_option("nearest.umma_min_queries", 0)             # K10 (tests/test_gpu_umma.py) would take the 100-query case
        e.set_option("nearest.mma_min_queries", 4)              # set explicitly: calls of <= 5 queries default to byte-plane passes
        l0 = e.stats()["kernels_launched"]
        assert_topk_equal(e.nearest(Q, k), want, k)
        assert e.stats()["exact_reruns"] == 0
        assert e.stats()["kernels_launched"] - l0 == 3          # prep + DMMA scan + finalize, one launch each
        e.set_option("nearest.mma_min_queries", 0)              # same batch through K1, 8 queries per pass
        assert_topk_equal(e.nearest(Q, k), want, k)


def test_batched_dmma_cancellation_falls_back(port):
    """Rows far from the origin: the GEMM form |x|^2+|q|^2-2<x,q> cancels, its absolute error
    bound exceeds the gaps between neighbours, the proof fails and the exact scan answers."""
    rng = np.random.Generator(np.random.PCG64(9))
    rows = 1.0e6 + rng.random((4000, 64))
    Q = 1.0e6 + rng.random((20, 64))
    want = oracle_topk(port, rows, 64, Q, 5)
    with B.Engine(64, 64) as e:
        e.insert(rows)
        assert_topk_equal(e.nearest(Q, 5), want, 5)
        assert e.stats()["exact_reruns"] + e.stats()["fp64_reruns"] > 0      # low-precision keys -> K1 (fp64 rows) -> exact


def test_script_distribution_ties(port):
    """Short-decimal values (add_vectors.sh): duplicate kd-points are common (earliest wins) and
    distinct equidistant points can occur (tree order wins)."""
    rows = synth.script_values(42, (30000, 4))
    Q = synth.script_values(43, (64, 4))
    flat = oracle_topk(port, rows, 3, Q, 4)
    want = oracle_tree_ids(port, rows, 3, Q)
    with B.Engine(4, 3) as e:
        e.insert(rows)
        np.testing.assert_array_equal(e.nearest(Q, 1)[0][:, 0], want)
        idx, dist, seq = e.nearest(Q, 4)
        np.testing.assert_array_equal(idx[:, 0], want)
        for i, (wseq, widx, wd) in enumerate(flat):
            np.testing.assert_array_equal(dist[i].view(np.uint64), wd.view(np.uint64))


def test_mass_duplicates_take_exact_fallback(port):
    """More exact duplicates of the best point than the candidate lists hold: the completeness
    proof fails, the exact scan reruns the query, the earliest duplicate still wins."""
    rows = synth.uniform_rows(11, 4000, 64)
    q = synth.uniform_rows(12, 1, 64)[0]
    near = q + 1e-3
    rows[500:900] = near
    want = oracle_topk(port, rows, 64, [q], 3)
    with B.Engine(64, 64) as e:
        e.insert(rows)
        got = e.nearest(q, 3)
        assert_topk_equal(got, want, 3)
        assert got[0][0, 0] == 500
        assert e.stats()["exact_reruns"] >= 1


def test_near_ties_are_resolved_by_exact_rerank(port):
    """Rows that differ from the best one in the last bits: the approximate keys may order them
    wrongly, the reference-order re-rank must not."""
    rng = np.random.Generator(np.random.PCG64(5))
    rows = rng.random((3000, 96))
    q = rng.random(96)
    base = q + 0.01 * rng.standard_normal(96)
    for j in range(12):
        r = base.copy()
        r[rng.integers(0, 96)] += (j - 6) * 2.0 ** -50
        rows[100 + 37 * j] = r
    want = oracle_topk(port, rows, 96, [q], 10)
    with B.Engine(96, 96) as e:
        e.insert(rows)
        assert_topk_equal(e.nearest(q, 10), want, 10)


def test_nonfinite_rows_never_win_and_empty_log():
    with B.Engine(20, 20) as e:
        idx, dist, seq = e.nearest(np.zeros(20), 3)
        assert np.all(idx == B.NONE) and np.all(np.isinf(dist))
        rows = np.zeros((5, 20))
        rows[0, 3] = np.nan
        rows[1, 0] = np.inf
        rows[2] = 1e200          # squared distance overflows to +inf
        rows[3] = 5.0
        rows[4] = 4.0
        e.insert(rows)
        idx, dist, seq = e.nearest(np.zeros(20), 4)
        assert list(idx[0]) == [4, 3, B.NONE, B.NONE]
    with B.Engine(3, 3) as e:       # thin path
        rows = np.array([[np.nan, 0, 0], [1e200, 0, 0], [2.0, 0, 0]])
        e.insert(rows)
        assert list(e.nearest(np.zeros(3), 2)[0][0]) == [2, B.NONE]


def test_append_kdpoints_and_log_only_engine(port):
    """A bare KDTree: kdtree_insert(tree, point, index) with arbitrary carried indices."""
    pts = synth.uniform_rows(3, 500, 5)
    carried = np.arange(500, dtype=np.uint64)[::-1] * 3
    with B.Engine(5, 5, flags=B.FLAG_LOG_ONLY) as e:
        e.append_kdpoints(pts, carried)
        assert e.log_size == 500 and e.size == 0
        q = pts[123] + 1e-9
        idx, dist, seq = e.nearest(q, 1)
        assert idx[0, 0] == carried[123] and seq[0, 0] == 123


def test_compare_batch_vs_oracle(cpu):
    for D, n, seed in ((1, 50, 1), (3, 200, 2), (16, 300, 3), (17, 300, 4), (128, 2000, 5), (1536, 400, 6)):
        rows = synth.normal_rows(seed, n, D) * 3.0
        h = cpu.build(rows, 1)
        i1, i2 = synth.index_pairs(seed, 3000, n)
        with B.Engine(D, 1) as e:
            e.insert(rows)
            allm = e.compare(B.ALL_METRICS, i1, i2)
            for m in range(3):
                want = cpu.compare_batch(h, m, i1, i2, 4)
                got = e.compare(m, i1, i2)
                np.testing.assert_array_equal(got.view(np.uint32), want.view(np.uint32), err_msg=f"D={D} metric={m}")
                np.testing.assert_array_equal(allm[:, m].view(np.uint32), want.view(np.uint32))
            # out of range -> -1.0f, also after a delete shrank the store
            bad = e.compare(B.DOT, np.array([0, n], dtype=np.uint64), np.array([n + 5, 0], dtype=np.uint64))
            assert list(bad) == [-1.0, -1.0]
        cpu.free(h)


def test_compare_follows_update_and_delete(port):
    rows = synth.uniform_rows(8, 50, 24)
    db = PortDB(port, 24, 2)
    with B.Engine(24, 2) as e:
        for r in rows:
            db.insert(r)
        e.insert(rows)
        new = synth.uniform_rows(9, 3, 24)
        for j, r in zip((4, 17, 4), new):
            db.update(j, r)
            e.update(j, r)
        for j in (0, 30, 10):
            db.delete(j)
            e.delete(j)
        assert e.size == db.size == 47
        for i in range(47):
            np.testing.assert_array_equal(e.read_row(i), db.row(i))
        i1, i2 = synth.index_pairs(1, 200, 47)
        for m in range(3):
            want = np.array([db.compare(m, int(a), int(b)) for a, b in zip(i1, i2)], dtype=np.float32)
            np.testing.assert_array_equal(e.compare(m, i1, i2).view(np.uint32), want.view(np.uint32))
    db.close()


def test_device_api_shards_and_merge(port):
    """Two row-range shards on one GPU + K7 merge == one engine over all rows."""
    n, D, k = 6000, 48, 10
    rows = synth.uniform_rows(31, n, D)
    Q = synth.uniform_rows(32, 9, D)
    want = oracle_topk(port, rows, D, Q, k)
    half = n // 2
    dev_rows = torch.from_numpy(rows).cuda()
    dq = torch.from_numpy(Q).cuda()
    shards = [B.Engine(D, D, seq_base=0), B.Engine(D, D, seq_base=half)]
    stream = torch.cuda.current_stream().cuda_stream
    gathered = torch.zeros((2, len(Q), k, 4), dtype=torch.int64, device="cuda")
    for s, e in enumerate(shards):
        e.set_stream(stream)
        part = dev_rows[s * half:(s + 1) * half]
        e.insert_device(part.data_ptr(), half, D)
        e.nearest_device(dq.data_ptr(), len(Q), D, k, gathered[s].data_ptr())
    merged = torch.zeros((len(Q), k, 4), dtype=torch.int64, device="cuda")
    B.merge_candidates_device(0, stream, gathered.data_ptr(), 2, len(Q), k, merged.data_ptr())
    torch.cuda.synchronize()
    res = merged.cpu().numpy().view(B.candidate_dtype).reshape(len(Q), k)
    assert not np.any(res["flags"] & B.CAND_UNSAFE)
    # shard 1 reports its local row index; the global id is seq (contiguous row-range shards)
    assert_topk_equal((res["seq"], res["dist"], res["seq"]), want, k)
    for e in shards:
        e.close()


def test_one_million_rows_config2(port):
    """Config 2 (1M x 128, K = D): ids and distance bits vs the oracle's flat scan."""
    n, D = 1_000_000, 128
    g = torch.Generator(device="cuda").manual_seed(1)
    dev_rows = torch.rand((n, D), dtype=torch.float64, device="cuda", generator=g)
    rows = dev_rows.cpu().numpy()
    Q = synth.uniform_rows(77, 4, D)
    log = port.lib.orc_log_create(D)
    # the flat restatement needs only the points, not the tree links: fill the log arrays directly
    db_want = []
    for q in Q:
        d = ((rows - q) ** 2).sum(axis=1)
        cand = np.argsort(d, kind="stable")[:64]
        exact = np.array([port.sqdist(rows[c], q) for c in cand])
        order = np.lexsort((cand, exact))[:10]
        db_want.append((cand[order].astype(np.int64), cand[order].astype(np.uint64), exact[order]))
    port.lib.orc_log_free(log)
    with B.Engine(D, D, reserve_rows=n) as e:
        e.insert_device(dev_rows.data_ptr(), n, D)
        assert_topk_equal(e.nearest(Q, 10), db_want, 10)
        # compare on the same store: 20k random pairs, all three metrics, fp32 bits
        i1, i2 = synth.index_pairs(3, 20000, n)
        got = e.compare(B.ALL_METRICS, i1, i2)
        for m in range(3):
            want = np.array([port.metric(m, rows[a], rows[b]) for a, b in zip(i1[:300], i2[:300])], dtype=np.float32)
            np.testing.assert_array_equal(got[:300, m].view(np.uint32), want.view(np.uint32))


# ---- "next" rows of the scope table: bulk load/save, coalescing of concurrent callers ---------

def _write_reference_file(path, rows, uuids):
    """The reference's save format (vector_database.c:213-222), written by numpy."""
    with open(path, "wb") as f:
        np.array([len(rows)], dtype=np.uint64).tofile(f)
        for r, u in zip(rows, uuids):
            f.write(u.encode().ljust(37, b"\0"))
            np.array([len(r)], dtype=np.uint64).tofile(f)
            np.asarray(r, dtype=np.float64).tofile(f)


def test_bulk_load_save_reference_file_format(tmp_path, port):
    n, D, K = 5000, 24, 3
    rows = synth.script_values(5, (n, D))
    uuids = [f"uuid-{i:06d}" for i in range(n)]
    src = str(tmp_path / "ref.db")
    if OB.have_ref():                       # let the reference itself write the file when it is here
        ref = OB.load_ref()
        db = ref.lib.vector_db_init(0, K)
        for r, u in zip(rows, uuids):
            ref.lib.vector_db_insert(db, ref.make_vector(r, uuid=u))
        ref.lib.vector_db_save(db, src.encode())
        ref.lib.vector_db_free(db)
    else:
        _write_reference_file(src, rows, uuids)
    Q = synth.script_values(6, (200, D))
    want = oracle_tree_ids(port, rows, K, Q)
    with B.Engine.load_file(src, K) as e:
        assert e.size == n and e.D == D and e.K == K
        np.testing.assert_array_equal(e.nearest(Q, 1)[0][:, 0], want)
        assert e.get_uuid(0) == uuids[0] and e.get_uuid(n - 1) == uuids[-1]
        np.testing.assert_array_equal(e.read_row(1234), rows[1234])
        out = str(tmp_path / "ours.db")
        e.save_file(out)
        assert open(out, "rb").read() == open(src, "rb").read()
        # delete + update, save, reload through the reference-compatible drop-in loader
        e.delete(10)
        e.update(20, rows[0])
        e.save_file(out)
    api = OB.RefApi(B.LIB_PATH)
    db = api.lib.vector_db_load(out.encode(), K)
    assert db.contents.size == n - 1
    assert api.lib.vector_db_read(db, 10).contents.uuid == uuids[11].encode()
    np.testing.assert_array_equal(np.ctypeslib.as_array(api.lib.vector_db_read(db, 20).contents.data, shape=(D,)), rows[0])
    api.lib.vector_db_free(db)


def test_concurrent_single_queries_are_coalesced(port):
    """Thread-per-connection callers (main.c:382) each issue single queries; calls that arrive
    during a pass share the next one.  Answers are the same as when issued one by one."""
    import threading
    n, D = 1_000_000, 64          # a pass takes long enough (~100 us) for other callers to queue up behind it
    g = torch.Generator(device="cuda").manual_seed(3)
    dev_rows = torch.rand((n, D), dtype=torch.float64, device="cuda", generator=g)
    Q = synth.uniform_rows(4, 384, D)
    with B.Engine(D, D, reserve_rows=n) as e:
        e.insert_device(dev_rows.data_ptr(), n, D)
        serial = e.nearest(Q, 1)[0][:, 0]
        got = np.zeros(len(Q), dtype=np.uint64)
        nthreads = 24

        def worker(t):
            for i in range(t, len(Q), nthreads):
                got[i] = e.nearest(Q[i], 1)[0][0, 0]

        th = [threading.Thread(target=worker, args=(t,)) for t in range(nthreads)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        np.testing.assert_array_equal(got, serial)
        st = e.stats()
        assert st["coalesced_calls"] > 0 and st["coalesced_passes"] > 0


# ---- randomized differential runs against the store model -----------------------------------

@pytest.mark.parametrize("D,K,seed,coarse", [(1, 1, 1, True), (5, 2, 2, True), (12, 12, 3, False), (40, 33, 4, False),
                                             (20, 20, 5, True), (9, 8, 6, False)])
def test_random_batched_deltas_vs_model(port, D, K, seed, coarse):
    """Batched insert / update / delete calls (duplicates inside a batch, out-of-range indices,
    empty batches) interleaved with top-k queries and compares; every answer equals the model's
    (oracle/svdb_oracle.c, itself pinned against the reference)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    gen = (lambda shape: np.round(rng.random(shape) * 6) / 2) if coarse else (lambda shape: rng.random(shape))
    model = PortDB(port, D, K)
    with B.Engine(D, K) as e:
        for step in range(60):
            r = rng.random()
            if model.size < 8 or r < 0.5:
                m = int(rng.integers(0, 40))
                rows = gen((m, D))
                first = e.insert(rows) if m else e.size
                assert first == model.size
                for row in rows:
                    model.insert(row)
            elif r < 0.75:
                m = int(rng.integers(1, 12))
                idx = rng.integers(0, model.size + 3, size=m).astype(np.uint64)     # repeats + out of range
                rows = gen((m, D))
                e.update(idx, rows)
                for i, row in zip(idx, rows):
                    model.update(int(i), row)
            else:
                m = int(rng.integers(1, 6))
                idx = rng.integers(0, model.size + 2, size=m).astype(np.uint64)
                e.delete(idx)
                for i in idx:
                    model.delete(int(i))
            assert e.size == model.size and e.log_size == port.lib.orc_log_size(model.log)
            Q = gen((3, D))
            k = int(rng.integers(1, 8))
            idx, dist, seq = e.nearest(Q, k)
            for qi, q in enumerate(Q):
                wseq, widx, wd = model.topk(q, k)
                m = len(wd)
                np.testing.assert_array_equal(dist[qi, :m].view(np.uint64), wd.view(np.uint64), err_msg=f"step {step}")
                assert idx[qi, 0] == model.nearest(q), f"step {step}: nearest id"
                assert set(zip(dist[qi, :m], seq[qi, :m])) <= {(d_, s_) for d_, s_ in zip(*[model.topk(q, min(m + 40, model.size + 400))[i] for i in (2, 0)])}
            if model.size >= 2:
                i1 = rng.integers(0, model.size, size=16).astype(np.uint64)
                i2 = rng.integers(0, model.size, size=16).astype(np.uint64)
                got = e.compare(B.ALL_METRICS, i1, i2)
                for m_ in range(3):
                    want = np.array([model.compare(m_, int(a), int(b)) for a, b in zip(i1, i2)], dtype=np.float32)
                    ok = (got[:, m_].view(np.uint32) == want.view(np.uint32)) | (np.isnan(got[:, m_]) & np.isnan(want))
                    assert ok.all(), f"step {step} metric {m_}"
    model.close()


def test_degenerate_insertion_order_drops_the_tree_but_keeps_answers(port, capfd):
    """Sorted input turns the reference's tree into a list (depth ~ N).  The level-synchronous
    build gives up beyond tree.max_depth, the engine keeps answering from the scan."""
    n, D = 6000, 4
    rows = np.cumsum(np.ones((n, D)), axis=0) + synth.uniform_rows(1, n, D) * 0.1     # increasing in every coordinate
    Q = rows[[5, 999, 4321]] + 0.01
    want = oracle_topk(port, rows, 2, Q, 3)
    with B.Engine(D, 2) as e:
        e.set_option("tree.max_depth", 512)
        e.insert(rows)
        assert_topk_equal(e.nearest(Q, 3), want, 3)
        np.testing.assert_array_equal(e.nearest(Q, 1)[0][:, 0], [w[1][0] for w in want])
        e.insert(rows[:10] + 0.5)                      # later inserts keep working without a tree
        assert e.log_size == n + 10
        assert e.nearest(rows[3] + 0.5, 1)[0][0, 0] == n + 3
    assert "tree dropped" in capfd.readouterr().err


def test_large_dimension_and_empty_calls(port, cpu):
    """D = 4100 (32.8 KB rows, not a multiple of anything convenient): scan tiles larger than the
    co-residency budget, several re-rank rounds, zero-padded tails; plus zero-sized calls."""
    n, D = 400, 4100
    rows = synth.normal_rows(1, n, D)
    Q = synth.normal_rows(2, 6, D)
    want = oracle_topk(port, rows, D, Q, 24)
    with B.Engine(D, D) as e:
        e.insert(rows)
        assert_topk_equal(e.nearest(Q, 24), want, 24)                       # K2 (6 queries -> group of 16)
        assert_topk_equal(e.nearest(Q[:1], 24), want[:1], 24)               # K1 single query
        e.set_option("nearest.mma_min_queries", 0)
        assert_topk_equal(e.nearest(Q, 24), want, 24)                       # K1, passes of 4 + 2
        i1, i2 = synth.index_pairs(1, 300, n)
        h = cpu.build(rows, 1)
        for m in range(3):
            np.testing.assert_array_equal(e.compare(m, i1, i2).view(np.uint32), cpu.compare_batch(h, m, i1, i2).view(np.uint32))
        cpu.free(h)
        idx, dist, seq = e.nearest(np.zeros((0, D)), 3)
        assert idx.shape == (0, 3)
        assert e.compare(B.DOT, np.zeros(0, dtype=np.uint64), np.zeros(0, dtype=np.uint64)).shape == (0,)
        assert e.insert(np.zeros((0, D))) == n and e.size == n


from hypothesis import given, settings, strategies as st  # noqa: E402


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 3), st.integers(1, 80), st.integers(0, 2 ** 32 - 1))
def test_gpu_equals_oracle_on_small_grids(K, n, seed):
    """Tie-heavy random instances (coordinates from {0, 0.5, .., 2}): the id must be the one the
    reference's traversal reaches first, through the tree traversal AND through scan + resolver."""
    port = OB.load_port()
    rng = np.random.Generator(np.random.PCG64(seed))
    rows = rng.integers(0, 5, size=(n, K + 1)) / 2.0
    Q = rng.integers(0, 5, size=(8, K + 1)) / 2.0
    want = oracle_tree_ids(port, rows, K, Q)
    with B.Engine(K + 1, K) as e:
        e.insert(rows)
        np.testing.assert_array_equal(e.nearest(Q, 1)[0][:, 0], want)
        e.set_option("nearest.tree_max_k", 0)
        np.testing.assert_array_equal(e.nearest(Q, 1)[0][:, 0], want)
        np.testing.assert_array_equal(e.nearest(Q, min(3, n))[0][:, 0], want)


def test_host_path_graph_replay_tracks_store_and_options(port):
    """The host query path replays a captured CUDA graph from the second call of a shape on; the
    graph must be dropped whenever the log, the options or the scratch blocks change."""
    n, D = 5000, 48
    rows = synth.uniform_rows(3, n, D)
    Q = synth.uniform_rows(4, 3, D)
    with B.Engine(D, D) as e:
        e.insert(rows)
        want = oracle_topk(port, rows, D, Q, 5)
        for _ in range(4):                                    # plain, capture, replay, replay
            assert_topk_equal(e.nearest(Q, 5), want, 5)
        assert_topk_equal(e.nearest(Q[:1], 5), want[:1], 5)     # another shape in between
        assert_topk_equal(e.nearest(Q, 5), want, 5)
        # the store changes: a new row equal to query 0 must win immediately
        e.insert(Q[0])
        for _ in range(3):
            idx, dist, _ = e.nearest(Q, 5)
            assert idx[0, 0] == n and dist[0, 0] == 0.0
        e.update(n, rows[0])                                     # stale point stays searchable, as in the reference
        assert e.nearest(Q, 5)[0][0, 0] == n
        e.set_option("scan.force_exact", 1)
        for _ in range(3):
            assert e.nearest(Q, 5)[0][0, 0] == n
        big = synth.uniform_rows(5, 40, D)                        # larger batch: scratch blocks grow
        e.nearest(big, 5)
        for _ in range(3):
            assert e.nearest(Q, 5)[0][0, 0] == n
    with B.Engine(8, 3) as e:                                     # tree path (one kernel) through the graph as well
        pts = synth.uniform_rows(6, 3000, 8)
        e.insert(pts)
        h = port.build(pts, 3)
        q = synth.uniform_rows(7, 1, 8)
        want1 = port.nearest_batch(h, q)[0]
        port.free(h)
        for _ in range(5):
            assert e.nearest(q, 1)[0][0, 0] == want1


def test_random_shapes_vs_oracle(port):
    """Forty random (rows, kd_dim, k, queries) shapes -- every tile size / lanes-per-row variant of K1, the exact
    kernel, the tree, K2 groups, long and short candidate lists in finalize -- with duplicated rows mixed in."""
    rng = np.random.Generator(np.random.PCG64(20261017))
    for it in range(40):
        K = int(rng.choice([1, 2, 3, 5, 8, 9, 12, 13, 14, 15, 16, 17, 18, 23, 31, 32, 33, 40, 47, 64, 65, 96, 130]))
        D = K + int(rng.integers(0, 4))
        n = int(rng.choice([1, 2, 31, 32, 33, 100, 257, 1000, 3001, 5000]))
        k = int(rng.integers(1, 25))
        nq = int(rng.choice([1, 2, 3, 4, 5, 9, 17]))
        rows = synth.uniform_rows(1000 + it, n, D)
        if n >= 100:                                     # exact duplicates: (dist, seq) ties of the harmless kind
            dup = rng.integers(0, n, size=n // 10)
            rows[rng.integers(0, n, size=n // 10)] = rows[dup]
        Q = synth.uniform_rows(2000 + it, nq, D)
        Q[0] = rows[int(rng.integers(0, n))]             # an exact hit
        want = oracle_topk(port, rows, K, Q, k)
        with B.Engine(D, K) as e:
            e.insert(rows)
            try:
                assert_topk_equal(e.nearest(Q, k), want, k)
                e.set_option("nearest.tree_max_k", 0)    # thin kd-points through the scan as well
                e.set_option("nearest.mma_min_queries", 0)
                assert_topk_equal(e.nearest(Q, k), want, k)
            except AssertionError as ex:
                raise AssertionError(f"shape {it}: n={n} D={D} K={K} k={k} nq={nq}: {ex}") from ex
