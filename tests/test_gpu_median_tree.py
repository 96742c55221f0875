"""GPU parity of the balanced median KD tree (K8 build + K9 traversal, csrc/median_tree.cu) for thin kd-points:
the id it returns -- with the queries it flags re-answered by the reference's own traversal (K6) -- must be the id
kdtree_nearest returns (src/kdtree.c:171-178), distances bit-identical, on random, tie-heavy, duplicated, sorted and
non-finite data, at every lanes-per-query variant, while the log grows under it."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

from oracle.binding import PortDB  # noqa: E402
from svdb import binding as B  # noqa: E402
from svdb import synth  # noqa: E402


def ref_ids(port, rows, K, Q):
    h = port.build(rows, K)
    ids = port.nearest_batch(h, Q)
    port.free(h)
    return ids


def exact_dist(rows, K, q, i):
    d = 0.0
    for c in range(K):                         # kdtree.c:134-137, sequential
        t = rows[i, c] - q[c]
        d = d + t * t
    return d


def _data(kind, n, D, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    if kind == "uniform":
        return synth.uniform_rows(seed, n, D)
    if kind == "script":                       # add_vectors.sh short decimals: duplicates, the occasional tie
        return synth.script_values(seed, (n, D))
    if kind == "grid":                         # coarse lattice: distinct points tie all the time
        return rng.integers(0, 9, size=(n, D)) / 2.0
    if kind == "dups":                         # every point stored three times
        base = synth.uniform_rows(seed, (n + 2) // 3, D)
        return np.concatenate([base, base, base])[:n][rng.permutation(n)]
    raise ValueError(kind)


CASES = [
    ("uniform", 200_000, 3, 400),      # seven radix-select levels, then single-CTA segments
    ("uniform", 70_001, 8, 200),       # K = 8
    ("uniform", 5_000, 1, 100),        # K = 1
    ("script", 100_000, 3, 400),
    ("grid", 30_000, 3, 300),
    ("grid", 3_000, 2, 300),
    ("dups", 50_000, 3, 300),
    ("uniform", 257, 3, 50),           # just above the build threshold
    ("uniform", 200, 3, 50),           # below it: tail scan only
    ("uniform", 2049, 2, 50),
]


@pytest.mark.parametrize("kind,n,K,nq", CASES)
def test_median_tree_returns_the_reference_id(port, kind, n, K, nq):
    D = K + 2
    rows = _data(kind, n, D, seed=n + K)
    Q = _data(kind, nq, D, seed=n + K + 1)
    if kind == "grid":
        Q[1::2] += 0.25                        # between lattice points: 2^K distinct points at the minimum
    if kind in ("script", "dups"):
        Q[::3] = rows[:: max(1, n // len(Q[::3]))][: len(Q[::3])]     # exact hits on (duplicated) points
    want = ref_ids(port, rows, K, Q)
    with B.Engine(D, K) as e:
        e.insert(rows)
        e.set_option("nearest.mtree", 0)
        i6, d6, s6 = e.nearest(Q, 1)                                   # K6: the reference's traversal
        np.testing.assert_array_equal(i6[:, 0], want)
        assert e.stats()["mtree_builds"] == 0
        e.set_option("nearest.mtree", 1)
        for blk in (1, 3):                                             # split values in heap order / 64-byte blocks of 3 levels
            e.set_option("mtree.block_levels", blk)
            for lanes in (32, 16, 8, 0):
                e.set_option("mtree.lanes", lanes)
                i9, d9, s9 = e.nearest(Q, 1)
                np.testing.assert_array_equal(i9[:, 0], want, err_msg=f"block {blk} lanes {lanes}")
                np.testing.assert_array_equal(s9, s6)
                np.testing.assert_array_equal(d9.view(np.uint64), d6.view(np.uint64))
        st = e.stats()
        assert st["mtree_builds"] == (2 if n > 256 else 0) and st["mtree_rows"] == (n if n > 256 else 0)
        for i in (0, nq // 2, nq - 1):
            assert d9[i, 0] == exact_dist(rows, K, Q[i], int(i9[i, 0]))
        one = e.nearest(Q[:1], 1)                                      # single-query host call (coalescing + graph path)
        assert one[0][0, 0] == want[0]
        # k > 1: a warp per query keeps the k smallest (distance, seq); position 0 stays the reference's answer
        for k in (2, 10, 24):
            e.set_option("nearest.mtree", 0)
            a = e.nearest(Q[:120], k)                                  # K6's k-smallest traversal
            e.set_option("nearest.mtree", 1)
            b = e.nearest(Q[:120], k)
            np.testing.assert_array_equal(b[0][:, 0], want[:120], err_msg=f"k {k}")
            for x, y in zip(a, b):
                np.testing.assert_array_equal(x.view(np.uint64), y.view(np.uint64), err_msg=f"k {k}")


def test_median_tree_follows_the_growing_log(port):
    """Entries appended after a build are scanned as a tail; the tree is rebuilt once the tail outgrows its limit.
    Updates append a kd-point that reports the updated index; deletes leave the log alone (kdtree.c has no remove)."""
    D, K = 5, 3
    rng = np.random.Generator(np.random.PCG64(5))
    rows = synth.uniform_rows(1, 6000, D)
    Q = synth.uniform_rows(2, 64, D)
    db = PortDB(port, D, K)
    with B.Engine(D, K) as e:
        builds = 0
        done = 0
        for step, m in enumerate([100, 200, 300, 1, 1, 600, 40, 3000, 300, 1400]):
            chunk = rows[done:done + m]
            for r in chunk:
                db.insert(r)
            e.insert(chunk)
            done += m
            if step == 6:
                up = rng.integers(0, done, size=20)
                for i in up:
                    v = synth.uniform_rows(100 + int(i), 1, D)[0]
                    db.update(int(i), v)
                    e.update(int(i), v)
            Qs = np.concatenate([Q, chunk[:8]])                        # fresh entries are found at distance 0
            got = e.nearest(Qs, 1)[0][:, 0]
            np.testing.assert_array_equal(got, [db.nearest(q) for q in Qs], err_msg=f"step {step}")
            st = e.stats()
            assert st["mtree_builds"] >= builds
            builds = st["mtree_builds"]
            assert e.log_size - st["mtree_rows"] <= max(256, min(4096, st["mtree_rows"] // 8))
        assert builds >= 3
        e.set_option("mtree.tail_max", 100000)                          # a long tail is legal, just slower
        e.set_option("mtree.tail_min", 100000)
        more = synth.uniform_rows(9, 5000, D)
        for r in more:
            db.insert(r)
        e.insert(more)
        np.testing.assert_array_equal(e.nearest(Q, 1)[0][:, 0], [db.nearest(q) for q in Q])
        assert e.stats()["mtree_builds"] == builds
    db.close()


def test_median_tree_survives_degenerate_order_and_nonfinite_rows(port, capfd):
    """Sorted input makes the reference's tree a list (K5 drops it); the median tree does not care and keeps
    answering in O(log n).  Rows with NaN / inf coordinates never win (kdtree.c:139 against INFINITY)."""
    n, D, K = 20_000, 4, 2
    rows = np.cumsum(np.ones((n, D)), axis=0) + synth.uniform_rows(1, n, D) * 0.1
    Q = rows[[5, 999, 4321, 19_999]] + 0.01
    db = PortDB(port, D, K)
    for r in rows:
        db.insert(r)
    with B.Engine(D, K) as e:
        e.set_option("tree.max_depth", 512)
        e.insert(rows)
        np.testing.assert_array_equal(e.nearest(Q, 1)[0][:, 0], [db.topk(q, 1)[1][0] for q in Q])
        assert e.stats()["mtree_builds"] == 1
    assert "tree dropped" in capfd.readouterr().err
    db.close()
    rows = synth.normal_rows(3, 5000, 3)
    rows[::5, 0] = np.nan
    rows[1::5, 1] = np.inf
    rows[2::5, 2] = -np.inf
    Q = synth.normal_rows(4, 40, 3)
    Q[0, 1] = np.nan                                                   # no finite distance at all -> nothing found
    want = ref_ids(port, rows, 3, Q)
    with B.Engine(3, 3) as e:
        e.insert(rows)
        got = e.nearest(Q, 1)[0][:, 0]
        np.testing.assert_array_equal(got[1:], want[1:])
        assert got[0] == B.NONE and want[0] == B.NONE


def test_median_tree_device_api_and_errors(port):
    n, D, K = 40_000, 3, 3
    rows = synth.uniform_rows(11, n, D)
    Q = synth.uniform_rows(12, 3000, D)
    want = ref_ids(port, rows, K, Q)
    dq = torch.from_numpy(Q).cuda()
    out = torch.zeros((len(Q), 4), dtype=torch.int64, device="cuda")
    with B.Engine(D, K, flags=B.FLAG_LOG_ONLY) as e:                   # a bare KDTree (kdtree_create / kdtree_insert)
        e.set_stream(torch.cuda.current_stream().cuda_stream)
        e.append_kdpoints(rows, np.arange(n) + 7)
        e.set_option("nearest.mtree", 0)
        e.nearest_device(dq.data_ptr(), len(Q), D, 1, out.data_ptr(), B.MODE_MTREE)   # explicit mode overrides AUTO's choice
        torch.cuda.synchronize()
        res = out.cpu().numpy().view(B.candidate_dtype).reshape(-1)
        np.testing.assert_array_equal(res["index"], want + 7)
        assert not np.any(res["flags"])
        assert e.stats()["mtree_builds"] == 1
        out5 = torch.zeros((len(Q), 5, 4), dtype=torch.int64, device="cuda")
        e.nearest_device(dq.data_ptr(), len(Q), D, 5, out5.data_ptr(), B.MODE_MTREE)
        torch.cuda.synchronize()
        res5 = out5.cpu().numpy().view(B.candidate_dtype).reshape(len(Q), 5)
        np.testing.assert_array_equal(res5["index"][:, 0], want + 7)
        assert np.all(np.diff(res5["dist"], axis=1) >= 0)
    with B.Engine(16, 16) as e:                                        # wide kd-points: no median tree
        e.insert(synth.uniform_rows(1, 100, 16))
        q16 = torch.zeros((1, 16), dtype=torch.float64, device="cuda")
        with pytest.raises(B.SvdbError):
            e.nearest_device(q16.data_ptr(), 1, 16, 1, out.data_ptr(), B.MODE_MTREE)
