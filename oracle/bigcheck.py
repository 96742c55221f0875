"""Parity at benchmark scale.  TEST INFRASTRUCTURE ONLY (used by bench.py, scripts/bench_extra.py and tests/).

The stores the headline numbers are measured on (10M x 768 = 7.7e9 elements, 61 GB; 100M x 128; 1M x 1536) are far too
large for the CPU oracle to scan, so the check has two independent halves:

  1. `brute_candidates`: a brute force over the WHOLE store that shares nothing with the product -- chunked torch fp64
     `((x - q)^2).sum(1)` on the GPU over rows REGENERATED from their seeds (not read back from the engine) -- keeps the
     `m` (64) rows with the smallest sums per query, together with the rows themselves.  torch's summation order differs
     from the reference's by a relative 1e-13 at most, the gaps between the 10th and the 64th neighbour are ~1e-2
     relative, so the true top-k of the reference is inside these m rows with an enormous margin (asserted: the margin
     is reported).
  2. `oracle_topk`: the CPU oracle (oracle/svdb_oracle.c: orc_sqdist, the restatement of kdtree.c:134-137, pinned
     against the compiled reference) recomputes those m distances in the reference's operation order; order by
     (d, seq); top-k.  When oracle/_ref is present the reference's OWN kdtree_nearest is also asked for its answer on a
     tree built from the m rows in seq order (top-1 id).

The product's answer must match ids and fp64 distance BITS.
"""
from __future__ import annotations

import numpy as np


def brute_candidates(chunks, Q, m: int = 64):
    """chunks: iterable of (first_global_row, torch fp64 CUDA tensor [rows, D]) covering this rank's store.
    Q: torch fp64 CUDA tensor [nq, K] (K <= D: the first K coordinates are measured).
    Returns (dist [nq, m] float64, row_id [nq, m] int64, rows [nq, m, K] float64) as numpy, ascending by torch's distance."""
    import torch
    nq, K = Q.shape
    best_d = torch.full((nq, m), float("inf"), dtype=torch.float64, device=Q.device)
    best_id = torch.full((nq, m), -1, dtype=torch.int64, device=Q.device)
    best_rows = torch.zeros((nq, m, K), dtype=torch.float64, device=Q.device)
    for first, rows in chunks:
        if rows.shape[0] == 0:
            continue
        x = rows[:, :K]
        for qi in range(nq):
            d = ((x - Q[qi]) ** 2).sum(1)
            kk = min(m, d.shape[0])
            dv, di = torch.topk(d, kk, largest=False)
            cat_d = torch.cat([best_d[qi], dv])
            cat_id = torch.cat([best_id[qi], di + first])
            cat_rows = torch.cat([best_rows[qi], x[di]])
            order = torch.argsort(cat_d, stable=True)[:m]
            best_d[qi], best_id[qi], best_rows[qi] = cat_d[order], cat_id[order], cat_rows[order]
    return best_d.cpu().numpy(), best_id.cpu().numpy(), best_rows.cpu().numpy()


def oracle_topk(port, cand_rows: np.ndarray, cand_ids: np.ndarray, q: np.ndarray, k: int):
    """Reference-order distances of the candidates (orc_sqdist), ordered by (d, seq).  Returns (ids [k], d [k])."""
    valid = cand_ids >= 0
    ids = cand_ids[valid]
    d = np.array([port.sqdist(r, q) for r in cand_rows[valid]], dtype=np.float64)
    order = np.lexsort((ids, d))[:k]
    return ids[order].astype(np.uint64), d[order]


def reference_nearest_among(ref, cand_rows: np.ndarray, cand_ids: np.ndarray, q: np.ndarray):
    """The compiled reference's own kdtree_nearest over the candidates inserted in seq order (top-1 id), or None."""
    if ref is None:
        return None
    valid = cand_ids >= 0
    ids, rows = cand_ids[valid], cand_rows[valid]
    order = np.argsort(ids, kind="stable")
    h = ref.build(np.ascontiguousarray(rows[order]), rows.shape[1])
    try:
        local = int(ref.nearest_batch(h, np.ascontiguousarray(q[None, :]), 1)[0])
    finally:
        ref.free(h)
    return int(ids[order][local])


def verdict(cands, Q: np.ndarray, got_ids: np.ndarray, got_d: np.ndarray, k: int, rows_total: int, m: int = 64):
    """cands: list over ranks of (dist, id, rows) from brute_candidates.  Q: [nq, K] numpy.  got_*: the product's [nq, k].
    Merges the per-rank candidate sets, re-ranks with the oracle, compares ids and distance bits."""
    from oracle import binding as OB
    port = OB.load_port()
    ref = OB.load_ref() if OB.have_ref() else None
    nq = Q.shape[0]
    bad = []
    margin = np.inf
    ref_top1_ok = 0
    for qi in range(nq):
        d = np.concatenate([c[0][qi] for c in cands])
        ids = np.concatenate([c[1][qi] for c in cands])
        rows = np.concatenate([c[2][qi] for c in cands])
        keep = np.argsort(d, kind="stable")[:m]
        d, ids, rows = d[keep], ids[keep], rows[keep]
        want_ids, want_d = oracle_topk(port, rows, ids, Q[qi], k)
        # how far the m-th torch distance is above the k-th: the room the brute force has for its summation-order noise
        fin = np.isfinite(d)
        if fin.sum() > k:
            margin = min(margin, float((d[fin][-1] - d[k - 1]) / max(d[k - 1], 1e-300)))
        same = np.array_equal(want_ids, got_ids[qi][:len(want_ids)].astype(np.uint64)) and \
            np.array_equal(want_d.view(np.uint64), got_d[qi][:len(want_d)].view(np.uint64))
        if not same:
            bad.append({"query": qi, "want_ids": want_ids.tolist()[:3], "got_ids": got_ids[qi].tolist()[:3]})
        r1 = reference_nearest_among(ref, rows, ids, Q[qi])
        if r1 is not None and r1 == int(got_ids[qi][0]):
            ref_top1_ok += 1
    out = {"queries": nq, "k": k, "rows": rows_total, "elements": rows_total * Q.shape[1], "candidates_per_query": m,
           "ok": not bad, "brute_force_margin_rel": None if not np.isfinite(margin) else margin,
           "oracle": "orc_sqdist (oracle/svdb_oracle.c, kdtree.c:134-137) on the 64 smallest of an independent torch-fp64 "
                     "brute force over rows regenerated from their seeds; ids and fp64 distance bits compared with =="}
    if ref is not None:
        out["reference_kdtree_nearest_top1_agrees"] = ref_top1_ok
    if bad:
        out["mismatches"] = bad[:4]
    return out
