"""oracle/bigcheck.py is what stands behind "parity at benchmark scale" (bench.py's parity_check, tests/test_gpu_big_store.py):
an independent torch brute force nominates 64 rows per query, the oracle re-ranks them.  Here the checker itself is checked, on
the CPU and at a size the oracle can scan completely: it must accept the oracle's own full-scan answers -- ids and distance
bits -- and reject answers that are off by one neighbour, by one row id, or by one bit of a distance."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import bigcheck
from oracle import binding as OB
from oracle.binding import PortDB


@pytest.fixture(scope="module")
def small_store():
    rng = np.random.Generator(np.random.PCG64(11))
    n, D, K = 6000, 40, 40
    rows = rng.random((n, D))
    rows[100] = rows[7]                       # an exact duplicate: equal distances, order by log sequence
    Q = rng.random((6, D))
    Q[5] = rows[7] + 1e-3
    db = PortDB(OB.load_port(), D, K)
    for r in rows:
        db.insert(r)
    want = [db.topk(q, 10) for q in Q]
    db.close()
    chunks = [(0, torch.from_numpy(rows[:2500])), (2500, torch.from_numpy(rows[2500:]))]
    cand = bigcheck.brute_candidates(chunks, torch.from_numpy(Q), 64)
    return rows, Q, want, cand


def answers(want, k):
    ids = np.array([w[1][:k] for w in want], dtype=np.uint64)
    d = np.array([w[2][:k] for w in want], dtype=np.float64)
    return ids, d


@pytest.mark.parametrize("k", [1, 10])
def test_checker_accepts_the_oracles_full_scan(small_store, k):
    rows, Q, want, cand = small_store
    ids, d = answers(want, k)
    v = bigcheck.verdict([cand], Q, ids, d, k, len(rows))
    assert v["ok"] and v["queries"] == len(Q) and v["elements"] == rows.size
    assert v["brute_force_margin_rel"] > 0
    if "reference_kdtree_nearest_top1_agrees" in v:            # oracle/_ref built: the reference's own kdtree_nearest was asked too
        assert v["reference_kdtree_nearest_top1_agrees"] == len(Q)


def test_checker_works_on_sharded_candidates(small_store):
    rows, Q, want, _ = small_store
    parts = [bigcheck.brute_candidates([(lo, torch.from_numpy(rows[lo:hi]))], torch.from_numpy(Q), 64)
             for lo, hi in ((0, 1500), (1500, 3000), (3000, 6000))]
    ids, d = answers(want, 10)
    assert bigcheck.verdict(parts, Q, ids, d, 10, len(rows))["ok"]


def test_checker_rejects_wrong_answers(small_store):
    rows, Q, want, cand = small_store
    ids, d = answers(want, 10)
    swapped = ids.copy()
    swapped[2, [3, 4]] = swapped[2, [4, 3]]                     # right set, wrong order
    assert not bigcheck.verdict([cand], Q, swapped, d, 10, len(rows))["ok"]
    other = ids.copy()
    other[0, 9] = (other[0, 9] + 1) % len(rows)                 # one wrong neighbour
    assert not bigcheck.verdict([cand], Q, other, d, 10, len(rows))["ok"]
    bit = d.copy()
    bit.view(np.uint64)[4, 0] ^= 1                              # one bit of one distance
    assert not bigcheck.verdict([cand], Q, ids, bit, 10, len(rows))["ok"]
    dup = ids.copy()                                            # the duplicate pair: the later copy must not come first
    q5 = list(dup[5])
    if 7 in q5 and 100 in q5:
        a, b = q5.index(7), q5.index(100)
        assert a < b
        dup[5, [a, b]] = dup[5, [b, a]]
        assert not bigcheck.verdict([cand], Q, dup, d, 10, len(rows))["ok"]
