"""Host-side logic of the multi-GPU path on CPU: shard ranges, the candidate exchange over a
world_size-2 gloo group and the (dist, seq) merge.  Each rank's local top-k is produced by the
ORACLE here (there is no GPU); the exchange + merge code is the product's."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, PKG

from svdb import binding as B
from svdb import synth
from svdb.sharded import merge_candidates_host, query_span, shard_range


def test_shard_ranges_cover_exactly():
    for n in (0, 1, 7, 8, 9, 1000, 10_000_000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b
            per = -(-n // world) if n else 0
            assert all(b - a <= per for a, b in spans)


def test_query_spans_of_replicated_stores_cover_every_query_once():
    for world in (1, 2, 3, 8):
        for nq in (0, 1, 2 * world - 1, 2 * world, 17, 1000, 65536):
            spans = [query_span(nq, world, r) for r in range(world)]
            if nq < 2 * world:                       # too few to split: everybody answers everything, no exchange
                assert all(s == (0, nq, 0) for s in spans)
                continue
            per = spans[0][2]
            assert all(s[2] == per for s in spans) and per * world >= nq
            assert spans[0][0] == 0 and spans[-1][1] == nq
            for (a, b, _), (c, d, _) in zip(spans, spans[1:]):
                assert b == c and a <= b <= a + per
            # the all-gather lays rank r's answers at r*per: query i sits at slot i
            assert all(lo == min(nq, r * per) for r, (lo, _, _) in enumerate(spans))


def test_merge_host_orders_by_dist_then_seq():
    c = np.zeros((2, 1, 3), dtype=B.candidate_dtype)
    c["dist"][0, 0] = [1.0, 2.0, 2.0]
    c["seq"][0, 0] = [5, 1, 9]
    c["dist"][1, 0] = [1.0, 2.0, np.inf]
    c["seq"][1, 0] = [3, 0, B.NONE]
    c["index"] = c["seq"]
    c["flags"][1, 0, 2] = 1
    m = merge_candidates_host(c, 3)
    assert list(m["seq"][0]) == [3, 5, 0] and list(m["dist"][0]) == [1.0, 1.0, 2.0]
    # UNSAFE of any input survives; TIE because two inputs (seq 3 and 5) sit at the merged minimum
    assert np.all(m["flags"][0] == (B.CAND_UNSAFE | B.CAND_TIE))
    c["dist"][1, 0, 0] = 1.5
    c["flags"][0, 0, 1] = B.CAND_TIE          # a shard's own flag counts only if that shard holds the minimum
    m = merge_candidates_host(c, 3)
    assert np.all(m["flags"][0] == B.CAND_UNSAFE)
    c["flags"][0, 0, 0] = B.CAND_TIE
    assert np.all(merge_candidates_host(c, 3)["flags"][0] == (B.CAND_UNSAFE | B.CAND_TIE))


def _worker(rank, world, port_no, n, D, k, out_dir):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import binding as OB
    from oracle.binding import PortDB
    port = OB.load_port()
    rows = synth.uniform_rows(3, n, D)           # every rank can regenerate any row range
    Q = synth.uniform_rows(4, 6, D)
    lo, hi = shard_range(n, world, rank)
    db = PortDB(port, D, D)
    for r in rows[lo:hi]:
        db.insert(r)
    local = np.zeros((len(Q), k), dtype=B.candidate_dtype)
    local["dist"], local["seq"], local["index"] = np.inf, B.NONE, B.NONE
    for i, q in enumerate(Q):
        seq, idx, d = db.topk(q, k)
        m = len(seq)
        local["dist"][i, :m], local["seq"][i, :m], local["index"][i, :m] = d, seq + lo, idx
    t_local = torch.from_numpy(local.view(np.int64).reshape(len(Q), k, 4).copy())
    gathered = [torch.zeros_like(t_local) for _ in range(world)]
    dist.all_gather(gathered, t_local)
    g = torch.stack(gathered).numpy().view(B.candidate_dtype).reshape(world, len(Q), k)
    merged = merge_candidates_host(g, k)
    np.save(os.path.join(out_dir, f"merged_{rank}.npy"), merged)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_exchange_and_merge_equals_single_scan(tmp_path, port):
    from oracle.binding import PortDB
    n, D, k, world = 1500, 24, 5, 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port_no = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(world, port_no, n, D, k, str(tmp_path)), nprocs=world, join=True)
    rows = synth.uniform_rows(3, n, D)
    Q = synth.uniform_rows(4, 6, D)
    db = PortDB(port, D, D)
    for r in rows:
        db.insert(r)
    merged = [np.load(tmp_path / f"merged_{r}.npy") for r in range(world)]
    np.testing.assert_array_equal(merged[0], merged[1])      # every rank holds the same answer
    for i, q in enumerate(Q):
        seq, idx, d = db.topk(q, k)
        np.testing.assert_array_equal(merged[0]["seq"][i].astype(np.int64), seq)
        np.testing.assert_array_equal(merged[0]["dist"][i].view(np.uint64), d.view(np.uint64))
    db.close()


# ---- exact ties across shards: the product's walk (svdb_tie_resolve) over a real process group ----

def _tie_worker(rank, world, port_no, seed, n, K, k, out_dir):
    for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tie_backend_np import NumpyTieShard, local_topk
    rng = np.random.Generator(np.random.PCG64(seed))        # every rank can regenerate any row range
    rows = rng.integers(0, 4, size=(n, K)) / 2.0
    Q = rng.integers(0, 4, size=(48, K)) / 2.0 + 0.25
    lo, hi = shard_range(n, world, rank)
    local = local_topk(rows[lo:hi], lo, Q, K, k)
    t_local = torch.from_numpy(local.view(np.int64).reshape(len(Q), k, 4).copy())
    gathered = [torch.zeros_like(t_local) for _ in range(world)]
    dist.all_gather(gathered, t_local)
    merged = merge_candidates_host(torch.stack(gathered).numpy().view(B.candidate_dtype).reshape(world, len(Q), k), k)

    def allgather(send, recv):
        mine = torch.from_numpy(np.array(send, copy=True))
        parts = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)
        recv[:] = torch.cat(parts).numpy()

    shard = NumpyTieShard(rows[lo:hi], lo, K, rank, world, allgather)
    levels = B.tie_resolve(shard.backend, Q, merged)
    np.save(os.path.join(out_dir, f"tie_{rank}.npy"), merged)
    np.save(os.path.join(out_dir, f"tie_levels_{rank}.npy"), np.array([levels]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_tie_walk_matches_the_global_tree(tmp_path, port):
    seed, n, K, k, world = 21, 600, 2, 3, 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port_no = s.getsockname()[1]
    s.close()
    mp.spawn(_tie_worker, args=(world, port_no, seed, n, K, k, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.Generator(np.random.PCG64(seed))
    rows = rng.integers(0, 4, size=(n, K)) / 2.0
    Q = rng.integers(0, 4, size=(48, K)) / 2.0 + 0.25
    h = port.build(rows, K)
    want = port.nearest_batch(h, Q)
    port.free(h)
    got = [np.load(tmp_path / f"tie_{r}.npy") for r in range(world)]
    np.testing.assert_array_equal(got[0], got[1])
    np.testing.assert_array_equal(got[0]["index"][:, 0], want)
    assert int(np.load(tmp_path / "tie_levels_0.npy")[0]) > 0      # the walk really ran
    assert not np.any(got[0]["flags"] & B.CAND_TIE)
