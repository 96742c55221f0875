"""Exact ties on a SHARDED store, CUDA backend: several SVDB_FLAG_SHARD engines on one GPU stand in for
the ranks (one thread each, an in-process all-gather), everything goes through the C-ABI
(svdb_nearest_batch_device -> svdb_merge_candidates_device -> svdb_resolve_ties_sharded), and the answer
must be the id the reference's tree over the WHOLE log returns.  The multi-process / multi-GPU run of the
same path is scripts/check_sharded.py (profiles/r01_check_sharded_*.json)."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

from svdb import binding as B  # noqa: E402
from svdb.sharded import merge_candidates_host, shard_range  # noqa: E402
from test_gpu_parity import _lattice_ties, oracle_tree_ids  # noqa: E402
from test_tie_protocol import ThreadRanks  # noqa: E402


class GpuShards:
    def __init__(self, rows, K, world, global_index=False):
        self.rows, self.K, self.world = rows, K, world
        self.spans = [shard_range(len(rows), world, r) for r in range(world)]
        self.engines = []
        for lo, hi in self.spans:
            e = B.Engine(rows.shape[1], K, seq_base=lo, flags=B.FLAG_SHARD)
            if global_index:
                e.set_option("log.index_base", lo)       # log entries report global row numbers
            if hi > lo:
                e.insert(rows[lo:hi])
            self.engines.append(e)

    def close(self):
        for e in self.engines:
            e.close()

    def local(self, Q, k, mode=B.MODE_AUTO):
        dq = torch.from_numpy(np.ascontiguousarray(Q)).cuda()
        out = torch.zeros((self.world, len(Q), k, 4), dtype=torch.int64, device="cuda")
        for r, e in enumerate(self.engines):
            e.nearest_device(dq.data_ptr(), len(Q), dq.stride(0), k, out[r].data_ptr(), mode)
        torch.cuda.synchronize()
        return out

    def merged(self, Q, k):
        """local candidates of every shard -> K7 on the device (escalating UNSAFE to the exact scan)."""
        g = self.local(Q, k)
        m = torch.zeros((len(Q), k, 4), dtype=torch.int64, device="cuda")
        B.merge_candidates_device(0, None, g.data_ptr(), self.world, len(Q), k, m.data_ptr())
        torch.cuda.synchronize()
        res = m.cpu().numpy().view(B.candidate_dtype).reshape(len(Q), k)
        if np.any(res["flags"] & B.CAND_UNSAFE):
            g = self.local(Q, k, B.MODE_EXACT)
            B.merge_candidates_device(0, None, g.data_ptr(), self.world, len(Q), k, m.data_ptr())
            torch.cuda.synchronize()
            res = m.cpu().numpy().view(B.candidate_dtype).reshape(len(Q), k)
        host = merge_candidates_host(g.cpu().numpy().view(B.candidate_dtype).reshape(self.world, len(Q), k), k)
        np.testing.assert_array_equal(res, host)                 # the kernel and its host restatement agree, flags too
        return res.copy()

    def resolve(self, Q, merged):
        ranks = ThreadRanks(self.world)
        results, errors = [None] * self.world, []

        def work(r):
            try:
                m = merged.copy()
                self.engines[r].resolve_ties_sharded(r, self.world, Q, m, allgather=ranks.allgather_for(r))
                results[r] = m
            except Exception as ex:
                errors.append(ex)
                ranks.bar.abort()

        th = [threading.Thread(target=work, args=(r,)) for r in range(self.world)]
        for t in th:
            t.start()
        for t in th:
            t.join(300)
        assert not errors, errors
        for r in range(1, self.world):
            np.testing.assert_array_equal(results[r], results[0])
        return results[0]


@pytest.mark.parametrize("world,K,D,levels", [(2, 2, 2, 4), (3, 3, 5, 5), (4, 1, 3, 6), (4, 12, 12, 2), (3, 40, 48, 2)])
def test_sharded_lattice_ties_equal_the_global_tree(port, world, K, D, levels):
    """Thin and wide kd-points on coarse lattices (K = 40 with 0/1 coordinates: Hamming distances, mass ties)."""
    rng = np.random.Generator(np.random.PCG64(1000 + world * 50 + K))
    n = 3000
    rows = rng.integers(0, levels, size=(n, D)) / 2.0
    Q = rng.integers(0, levels, size=(40, D)) / 2.0
    Q[:, 0] += 0.25
    want = oracle_tree_ids(port, rows, K, Q)
    sh = GpuShards(rows, K, world)
    try:
        merged = sh.merged(Q, 1)
        n_flag = np.count_nonzero(merged["flags"][:, 0] & B.CAND_TIE)
        assert n_flag > 0
        got = sh.resolve(Q, merged)
        np.testing.assert_array_equal(got["seq"][:, 0], want)      # insert-only data: global row == seq
        assert not np.any(got["flags"] & B.CAND_TIE)
        st = sh.engines[0].stats()
        assert st["tie_events"] == n_flag
        # top-k: winner first, the rest in (dist, seq) order
        k = 4
        merged = sh.merged(Q, k)
        got = sh.resolve(Q, merged)
        np.testing.assert_array_equal(got["seq"][:, 0], want)      # insert-only data: global row == seq
        for g, m in zip(got, merged):
            rest = [int(s) for s in m["seq"] if s != g["seq"][0]][:k - 1]
            assert [int(s) for s in g["seq"][1:1 + len(rest)]] == rest
    finally:
        sh.close()


@pytest.mark.parametrize("K,n_far", [(3, 2000), (24, 4000)])
def test_sharded_mass_ties_beyond_any_candidate_list(port, K, n_far):
    """30 / 2256 distinct entries at exactly distance 25, spread over the shards."""
    rows = _lattice_ties(K, n_far, seed=K)
    Q = np.concatenate([np.zeros((1, K)), 1e-3 * np.random.Generator(np.random.PCG64(K)).standard_normal((4, K))])
    want = oracle_tree_ids(port, rows, K, Q)
    sh = GpuShards(rows, K, 4)
    try:
        merged = sh.merged(Q, 1)
        assert merged["flags"][0, 0] & B.CAND_TIE and merged["dist"][0, 0] == 25.0
        got = sh.resolve(Q, merged)
        np.testing.assert_array_equal(got["seq"][:, 0], want)      # insert-only data: global row == seq
    finally:
        sh.close()


def test_shard_engines_report_plain_order_and_flags(port):
    """A SVDB_FLAG_SHARD engine never reorders by its local tree: (dist, seq) order + SVDB_CAND_TIE."""
    rows = np.array([[5.0, 5.0], [1.0, 0.0], [0.0, 1.0], [1.0, 0.0], [7.0, 7.0], [7.0, 7.0]])
    Q = np.array([[0.0, 0.0], [7.0, 7.5], [5.0, 5.0]])
    sh = GpuShards(rows, 2, 1)
    try:
        res = sh.local(Q, 3)[0].cpu().numpy().view(B.candidate_dtype).reshape(3, 3)
        assert list(res["seq"][0]) == [1, 2, 3] and np.all(res["flags"][0] & B.CAND_TIE)      # distinct points tie
        assert list(res["seq"][1][:2]) == [4, 5] and not np.any(res["flags"][1] & B.CAND_TIE)  # duplicates only
        assert res["seq"][2][0] == 0 and not np.any(res["flags"][2] & B.CAND_TIE)
    finally:
        sh.close()


def test_duplicate_rows_across_shards_skip_the_walk(port):
    rng = np.random.Generator(np.random.PCG64(5))
    base = rng.random((500, 64))
    rows = np.concatenate([base, base])
    Q = base[:16] + 1e-4
    want = oracle_tree_ids(port, rows, 64, Q)
    sh = GpuShards(rows, 64, 2)
    try:
        merged = sh.merged(Q, 2)
        assert np.all(merged["flags"][:, 0] & B.CAND_TIE)
        got = sh.resolve(Q, merged)
        np.testing.assert_array_equal(got["seq"][:, 0], want)      # insert-only data: global row == seq
        assert sh.engines[0].stats()["tie_levels"] == 0
    finally:
        sh.close()


@pytest.mark.parametrize("K,D,levels", [(3, 6, 5), (32, 32, 2)])
def test_one_call_sharded_path_with_peer_exchange_world_1(port, K, D, levels):
    """svdb_nearest_batch_sharded end to end (scan, peer-memory exchange, merge, tie walk over the exchange's
    own all-gather) with a world of one rank: a SHARD engine has no tree, so every tie goes through the walk."""
    rng = np.random.Generator(np.random.PCG64(77 + K))
    rows = rng.integers(0, levels, size=(5000, D)) / 2.0
    Q = rng.integers(0, levels, size=(33, D)) / 2.0
    Q[:, 0] += 0.25                                  # half way between lattice planes: never an exact hit
    want = oracle_tree_ids(port, rows, K, Q)
    xch = B.Exchange(0, 0, 1, 4096)
    xch.connect(xch.handle)
    try:
        with B.Engine(D, K, flags=B.FLAG_SHARD) as e:
            e.insert(rows)
            for _ in range(3):                       # the third call replays the captured graph
                idx, dist, seq = e.nearest_sharded(xch, Q, 1)
                np.testing.assert_array_equal(idx[:, 0], want)
            idx, dist, seq = e.nearest_sharded(xch, Q, 5)
            np.testing.assert_array_equal(idx[:, 0], want)
            assert np.all(np.diff(dist[:, 1:], axis=1) >= 0)
            st = e.stats()
            assert st["tie_events"] > 0 and st["tie_levels"] > 0
    finally:
        xch.close()


# ---- inserts and updates after the bulk ingest: the log grows at its tail (the last shard) ------------------

def test_sharded_store_follows_inserts_and_updates(port):
    """A mutable sharded store: vector_db_insert / vector_db_update append kd-points to the END of the global log
    (the tail shard), old points stay searchable (vector_database.c:174), answers == the reference-shaped tree
    over the whole log after every step -- on lattice values, so stale and tied entries do turn up."""
    from oracle.binding import PortDB
    rng = np.random.Generator(np.random.PCG64(31))
    K, D, world, n0 = 3, 5, 3, 600
    rows = rng.integers(0, 6, size=(n0, D)) / 2.0
    sh = GpuShards(rows, K, world, global_index=True)
    db = PortDB(port, D, K)
    for r in rows:
        db.insert(r)
    n_rows = n0
    try:
        for step in range(40):
            v = rng.integers(0, 6, size=D) / 2.0
            tail = sh.engines[-1]
            if step % 3 == 0:                                    # insert: a new global row
                assert db.insert(v) == n_rows
                tail.insert(v[None, :])
                n_rows += 1
            else:                                                # update of a random existing row: re-append, same index
                j = int(rng.integers(0, n_rows))
                db.update(j, v)
                tail.append_kdpoints(v[None, :K], np.array([j], dtype=np.uint64))
            Q = rng.integers(0, 6, size=(6, D)) / 2.0
            Q[:, 0] += 0.25
            want = np.array([db.nearest(q) for q in Q], dtype=np.uint64)
            merged = sh.merged(Q, 1)
            got = sh.resolve(Q, merged)
            np.testing.assert_array_equal(got["index"][:, 0], want, err_msg=f"step {step}")
        assert sh.engines[0].stats()["tie_events"] > 0
    finally:
        sh.close()
        db.close()
