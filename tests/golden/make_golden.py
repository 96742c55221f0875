"""Generate tests/golden/*.npz by RUNNING THE REFERENCE ITSELF.

Needs oracle/_ref/libsvdb_ref.so (built from /root/reference by oracle/Makefile),
so it runs only in the build container; the fixtures it writes are committed and
travel to the GPU box, where /root/reference does not exist.

    python tests/golden/make_golden.py

Every expected value in the fixtures is an output of the reference's own
kdtree_nearest / cosine_similarity / euclidean_distance / dot_product /
vector_db_insert / vector_db_update / vector_db_delete (called through ctypes),
never of our port.  Inputs are stored next to the outputs.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "simple-vector-db_b200"))

from oracle import binding  # noqa: E402
from svdb import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def nearest_case(ref, name, rows, K, queries):
    h = ref.build(rows, K)
    ids = ref.nearest_batch(h, queries)
    ref.free(h)
    np.savez_compressed(os.path.join(OUT, name), rows=rows, K=np.int64(K), queries=queries, ids=ids)
    print(name, rows.shape, "K", K, "queries", len(queries))


def metrics_case(ref, name):
    a_list, b_list, res = [], [], []
    dims = [1, 2, 3, 7, 16, 128, 129, 768, 1536]
    for D in dims:
        for kind in range(3):
            if kind == 0:
                ab = synth.uniform_rows(1000 + D, 8, D)
            elif kind == 1:
                ab = synth.normal_rows(2000 + D, 8, D) * 37.5
            else:
                ab = synth.script_values(3000 + D, (8, D))
            for j in range(0, 8, 2):
                a, b = ab[j], ab[j + 1]
                a_list.append(a)
                b_list.append(b)
                res.append([ref.metric(m, a, b) for m in range(3)])
    # self-pairs and a cancelling pair
    v = synth.uniform_rows(7, 1, 64)[0]
    for a, b in ((v, v), (v, -v), (np.zeros(4), np.ones(4))):
        a_list.append(a)
        b_list.append(b)
        with np.errstate(all="ignore"):
            res.append([ref.metric(m, a, b) for m in range(3)])
    lens = np.array([len(a) for a in a_list], dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, name), lens=lens, a=np.concatenate(a_list), b=np.concatenate(b_list),
                        res=np.array(res, dtype=np.float32))
    print(name, len(a_list), "pairs")


def delta_case(ref, name, D=6, K=2, n_ops=400, seed=11):
    """Random insert / update / delete stream with a nearest query after every op.

    ops[i] = (code, index): 0 insert, 1 update, 2 delete.  vals[i] = the row used
    (zeros for delete).  After each op: reference nearest id for queries[i], and
    the store size.  Values come from a coarse grid so stale entries, duplicates
    and index drift after deletes all occur.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    L = ref.lib
    db = L.vector_db_init(0, K)
    ops = np.zeros((n_ops, 2), dtype=np.int64)
    vals = np.zeros((n_ops, D))
    queries = np.round(rng.random((n_ops, D)) * 8) / 2
    ids = np.zeros(n_ops, dtype=np.uint64)
    sizes = np.zeros(n_ops, dtype=np.int64)
    for i in range(n_ops):
        size = db.contents.size
        r = rng.random()
        if size < 4 or r < 0.55:
            v = np.round(rng.random(D) * 8) / 2
            got = L.vector_db_insert(db, ref.make_vector(v, uuid=f"u{i}"))
            assert got == size
            ops[i] = (0, got)
            vals[i] = v
        elif r < 0.8:
            j = int(rng.integers(0, size + 2))          # sometimes out of range: silent no-op
            v = np.round(rng.random(D) * 8) / 2
            vec = ref.make_vector(v, uuid=f"u{i}")
            L.vector_db_update(db, j, vec)
            ops[i] = (1, j)
            vals[i] = v
        else:
            j = int(rng.integers(0, size + 2))
            L.vector_db_delete(db, j)
            ops[i] = (2, j)
        ids[i] = L.kdtree_nearest(db.contents.kdtree, queries[i].ctypes.data_as(binding._dp))
        sizes[i] = db.contents.size
    L.vector_db_free(db)
    np.savez_compressed(os.path.join(OUT, name), D=np.int64(D), K=np.int64(K), ops=ops, vals=vals,
                        queries=queries, ids=ids, sizes=sizes)
    print(name, n_ops, "ops")


def ties_case(ref, name):
    """Hand-made tie situations: duplicates and distinct equidistant points."""
    cases = []
    # (rows, K, query): duplicate kd-points -> earliest wins
    cases.append((np.array([[1.0, 1.0], [3.0, 3.0], [1.0, 1.0], [1.0, 1.0]]), 2, np.array([1.0, 1.25])))
    # the survey's counter-example: traversal order beats lowest id
    cases.append((np.array([[0.5, 100.0], [-2.0, 0.0], [4.0, 0.0]]), 2, np.array([1.0, 0.0])))
    # symmetric equidistant pair around the query, both orders of insertion
    cases.append((np.array([[0.0, 0.0], [2.0, 0.0]]), 2, np.array([1.0, 0.0])))
    cases.append((np.array([[2.0, 0.0], [0.0, 0.0]]), 2, np.array([1.0, 0.0])))
    # prefix subspace: only the first K coordinates count
    cases.append((np.array([[0.0, 0.0, 9.0], [1.0, 1.0, 0.0]]), 2, np.array([0.1, 0.1, 0.0])))
    # K = 1
    cases.append((np.array([[5.0], [1.0], [3.0], [3.0]]), 1, np.array([2.9])))
    rows_l, q_l, meta, ids = [], [], [], []
    for rows, K, q in cases:
        h = ref.build(rows, K)
        ids.append(ref.nearest_batch(h, q[None, :])[0])
        ref.free(h)
        meta.append((rows.shape[0], rows.shape[1], K))
        rows_l.append(rows.ravel())
        q_l.append(q)
    np.savez_compressed(os.path.join(OUT, name), meta=np.array(meta, dtype=np.int64),
                        rows=np.concatenate(rows_l), queries=np.concatenate(q_l),
                        ids=np.array(ids, dtype=np.uint64))
    print(name, len(cases), "cases ->", ids)


def main():
    binding.build()
    ref = binding.load_ref()
    assert ref.kind == "reference"
    # config 1 shape, shrunk: script-distributed values, K = 3 prefix of wider rows
    nearest_case(ref, "nearest_script_k3", synth.script_values(42, (2000, 8)), 3, synth.script_values(43, (400, 8)))
    # very coarse values: many exact duplicates and exact ties
    nearest_case(ref, "nearest_coarse_k3", np.round(synth.uniform_rows(5, 1500, 4) * 6) / 2, 3,
                 np.round(synth.uniform_rows(6, 300, 4) * 6) / 2)
    nearest_case(ref, "nearest_uniform_k32", synth.uniform_rows(1, 1200, 32), 32, synth.uniform_rows(2, 100, 32))
    nearest_case(ref, "nearest_normal_k128", synth.normal_rows(3, 500, 128), 128, synth.normal_rows(4, 60, 128))
    nearest_case(ref, "nearest_uniform_k5_d12", synth.uniform_rows(8, 1000, 12), 5, synth.uniform_rows(9, 100, 12))
    metrics_case(ref, "metrics")
    delta_case(ref, "delta_ops")
    ties_case(ref, "ties")


if __name__ == "__main__":
    main()
