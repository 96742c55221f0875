"""The cross-shard tie walk (csrc/tie_protocol.cu: svdb_tie_resolve) on CPU ranks.

Shards are numpy (tests/tie_backend_np.py), ranks are threads (an in-process all-gather) or gloo
processes (test_sharding_gloo.py); the walk is the product's C++.  Expected answers come from the
reference-shaped tree over the WHOLE log (oracle port, pinned to the reference in test_oracle.py;
and oracle/_ref itself where it is built): the sharded answer must be the same id, tie or no tie."""
import threading

import numpy as np
import pytest

from svdb import binding as B
from svdb.sharded import merge_candidates_host, shard_range
from tie_backend_np import NumpyTieShard, local_topk


class ThreadRanks:
    """world threads + an all-gather over a barrier."""

    def __init__(self, world):
        self.world = world
        self.bar = threading.Barrier(world)
        self.slots = [None] * world

    def allgather_for(self, rank):
        def ag(send, recv):
            self.slots[rank] = bytes(send)
            self.bar.wait()
            recv[:] = np.frombuffer(b"".join(self.slots), dtype=np.uint8)
            self.bar.wait()
        return ag


def run_sharded(rows, Q, K, k, world):
    """Every rank: local top-k -> (the all-gather + merge, done once here) -> the walk. Returns rank 0's answer."""
    n = len(rows)
    spans = [shard_range(n, world, r) for r in range(world)]
    local = np.stack([local_topk(rows[lo:hi], lo, Q, K, k) for lo, hi in spans])
    merged0 = merge_candidates_host(local, k)
    ranks = ThreadRanks(world)
    results, errors, shards = [None] * world, [], []

    def work(r):
        try:
            lo, hi = spans[r]
            shard = NumpyTieShard(rows[lo:hi], lo, K, r, world, ranks.allgather_for(r))
            shards.append(shard)
            m = merged0.copy()
            B.tie_resolve(shard.backend, Q, m)
            results[r] = m
        except Exception as ex:       # pragma: no cover
            errors.append(ex)
            ranks.bar.abort()

    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(120)
    assert not errors, errors
    for r in range(1, world):
        np.testing.assert_array_equal(results[r], results[0])
    return results[0], merged0, shards


def tree_ids(port, rows, Q, K):
    h = port.build(rows, K)
    ids = port.nearest_batch(h, Q)
    port.free(h)
    return ids


@pytest.mark.parametrize("world", [2, 3, 4])
@pytest.mark.parametrize("K,levels_hi", [(1, 3), (2, 4), (3, 5)])
def test_lattice_ties_resolve_like_the_global_tree(port, world, K, levels_hi):
    """Coordinates from a few half-integers: distinct equidistant points and duplicates everywhere."""
    rng = np.random.Generator(np.random.PCG64(100 * world + K))
    rows = rng.integers(0, levels_hi, size=(400, K + 1)) / 2.0
    Q = rng.integers(0, levels_hi, size=(64, K + 1)) / 2.0 + 0.25 * rng.integers(0, 2, size=(64, K + 1))
    want = tree_ids(port, rows, Q, K)
    got, merged0, _ = run_sharded(rows, Q, K, 1, world)
    assert np.count_nonzero(merged0["flags"][:, 0] & B.CAND_TIE) > 5          # the case is not vacuous
    if K >= 2:
        assert np.count_nonzero(merged0["index"][:, 0] != want) > 0           # and (dist, seq) alone is wrong
    np.testing.assert_array_equal(got["index"][:, 0], want)
    assert not np.any(got["flags"] & B.CAND_TIE)


def test_topk_keeps_dist_seq_order_behind_the_winner(port):
    rng = np.random.Generator(np.random.PCG64(7))
    K, k, world = 2, 5, 3
    rows = rng.integers(0, 4, size=(300, K)) / 2.0
    Q = rng.integers(0, 4, size=(40, K)) / 2.0 + 0.25
    want = tree_ids(port, rows, Q, K)
    got, merged0, _ = run_sharded(rows, Q, K, k, world)
    np.testing.assert_array_equal(got["index"][:, 0], want)
    for g, m in zip(got, merged0):
        # winner first, the others exactly the (dist, seq) list without it, cut to k
        rest = [int(s) for s in m["seq"] if s != g["seq"][0]][:k - 1]
        assert [int(s) for s in g["seq"][1:1 + len(rest)]] == rest
        assert np.all(np.diff(g["dist"][1:]) >= 0)


def test_duplicates_only_need_no_walk(port):
    """Identical kd-points in different shards: one collect + one all-gather, no tree level is walked."""
    rng = np.random.Generator(np.random.PCG64(9))
    base = rng.random((50, 3))
    rows = np.concatenate([base, base, base])           # every point three times, one copy per shard
    Q = base[:20] + 1e-3
    want = tree_ids(port, rows, Q, 3)
    got, merged0, shards = run_sharded(rows, Q, 3, 1, 3)
    assert np.all(merged0["flags"][:, 0] & B.CAND_TIE)
    np.testing.assert_array_equal(got["index"][:, 0], want)
    np.testing.assert_array_equal(want, np.arange(20))  # the earliest copy
    assert all(s.calls["first"] == 0 and s.calls["split"] == 0 for s in shards)


def test_no_flag_no_communication(port):
    rng = np.random.Generator(np.random.PCG64(11))
    rows, Q = rng.random((200, 4)), rng.random((10, 4))
    got, merged0, shards = run_sharded(rows, Q, 4, 2, 2)
    np.testing.assert_array_equal(got, merged0)
    assert all(s.calls["collect"] == 0 for s in shards)
    np.testing.assert_array_equal(got["index"][:, 0], tree_ids(port, rows, Q, 4))


def test_hidden_tie_inside_one_shard(port):
    """Both tied entries live in ONE shard (its flag is the only hint), the node that separates them in another."""
    rows = np.array([[5.0, 5.0],      # shard 0: the root, far away
                     [9.0, 9.0],
                     [1.0, 0.0],      # shard 1: two points at distance 1 from the query (0, 0) ...
                     [0.0, 1.0]])     # ... (1,0) is reached first?  the tree decides, not the seq
    Q = np.array([[0.0, 0.0]])
    want = tree_ids(port, rows, Q, 2)
    got, merged0, _ = run_sharded(rows, Q, 2, 1, 2)
    assert merged0["flags"][0, 0] & B.CAND_TIE
    np.testing.assert_array_equal(got["index"][:, 0], want)


def test_mass_ties_on_a_binary_lattice(port):
    """0/1 coordinates in 12 dimensions (Hamming distances): dozens of entries tie at the minimum."""
    rng = np.random.Generator(np.random.PCG64(13))
    rows = rng.integers(0, 2, size=(1500, 12)).astype(np.float64)
    Q = rng.integers(0, 2, size=(30, 12)).astype(np.float64)
    Q[:, 0] = 0.5                                       # never an exact hit: the minimum is shared widely
    want = tree_ids(port, rows, Q, 12)
    got, merged0, _ = run_sharded(rows, Q, 12, 1, 4)
    np.testing.assert_array_equal(got["index"][:, 0], want)


def test_against_the_compiled_reference(ref):
    rng = np.random.Generator(np.random.PCG64(17))
    rows = rng.integers(0, 5, size=(500, 3)) / 2.0
    Q = rng.integers(0, 5, size=(80, 3)) / 2.0
    h = ref.build(rows, 3)
    want = ref.nearest_batch(h, Q)
    ref.free(h)
    got, _, _ = run_sharded(rows, Q, 3, 1, 4)
    np.testing.assert_array_equal(got["index"][:, 0], want)


def test_inconsistent_ranks_fail_loudly(port):
    """A flagged minimum that no shard holds (merged candidates differ from the data): an error, not a hang."""
    rows = np.array([[0.0], [1.0], [2.0], [3.0]])
    Q = np.array([[0.5]])
    merged = np.zeros((1, 1), dtype=B.candidate_dtype)
    merged[0, 0] = (123.0, 0, 0, B.CAND_TIE)            # nobody is at distance 123
    ranks = ThreadRanks(1)
    shard = NumpyTieShard(rows, 0, 1, 0, 1, ranks.allgather_for(0))
    with pytest.raises(B.SvdbError):
        B.tie_resolve(shard.backend, Q, merged)


def test_many_small_random_instances(port):
    """Tie-heavy random instances of every small shape: the order of the walk, not just the distances."""
    rng = np.random.Generator(np.random.PCG64(2024))
    for it in range(120):
        K = int(rng.integers(1, 4))
        n = int(rng.integers(1, 70))
        world = int(rng.integers(2, 5))
        k = int(rng.integers(1, 4))
        rows = rng.integers(0, 5, size=(n, K + 1)) / 2.0
        Q = rng.integers(0, 5, size=(6, K + 1)) / 2.0
        want = tree_ids(port, rows, Q, K)
        got, _, _ = run_sharded(rows, Q, K, k, world)
        np.testing.assert_array_equal(got["index"][:, 0], want, err_msg=f"instance {it}: n={n} K={K} world={world} k={k}")
