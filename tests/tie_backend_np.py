"""A numpy shard for the cross-shard tie walk (svdb_tie_resolve): TEST INFRASTRUCTURE.

The walk itself -- which cell to enter, when to stop, how the winner is applied -- is the product's C++
(csrc/tie_protocol.cu); a backend only answers three shard-local questions.  On a GPU box the engine
answers them with CUDA kernels; here numpy does, so that the collective logic can run on CPU ranks
(threads or gloo processes) and be checked against the reference's own tree.
"""
import ctypes as C

import numpy as np

from svdb import binding as B


def seq_sqdist(rows: np.ndarray, q: np.ndarray, K: int) -> np.ndarray:
    """kdtree.c:134-137 for every row: rounded sub, mul, add, in index order."""
    d = np.zeros(len(rows))
    for i in range(K):
        t = rows[:, i] - q[i]
        d = d + t * t
    return d


def local_topk(rows: np.ndarray, lo: int, Q: np.ndarray, K: int, k: int) -> np.ndarray:
    """What a SVDB_FLAG_SHARD engine returns: (dist, seq) order, SVDB_CAND_TIE when distinct points tie at the minimum."""
    out = np.zeros((len(Q), k), dtype=B.candidate_dtype)
    out["dist"], out["seq"], out["index"] = np.inf, B.NONE, B.NONE
    for i, q in enumerate(Q):
        d = seq_sqdist(rows, q, K)
        d = np.where(np.isfinite(d), d, np.inf)
        order = np.lexsort((np.arange(len(rows)), d))[:k]
        order = order[np.isfinite(d[order])]
        m = len(order)
        out["dist"][i, :m], out["seq"][i, :m], out["index"][i, :m] = d[order], order + lo, order + lo
        if m:
            tied = np.flatnonzero(d == d[order[0]])
            if len(tied) >= 2 and np.any(rows[tied, :K] != rows[tied[0], :K]):
                out["flags"][i] = B.CAND_TIE
    return out


class NumpyTieShard:
    """rows[lo:hi] of the global log, as one rank of the walk."""

    def __init__(self, rows_local: np.ndarray, lo: int, K: int, rank: int, world: int, allgather):
        self.rows, self.lo, self.K = np.ascontiguousarray(rows_local[:, :K]), lo, K
        self.tied = []
        self.calls = {"collect": 0, "first": 0, "split": 0}
        self._ag = B.make_allgather_cb(allgather, world)
        self._collect = B.TIE_COLLECT_FN(self._collect_cb)
        self._first = B.TIE_FIRST_FN(self._first_cb)
        self._split = B.TIE_SPLIT_FN(self._split_cb)
        self.backend = B.TieBackend(None, world, rank, K, self._ag, None, self._collect, self._first, self._split)

    def _cells(self, na, depth, pv, ps, ld):
        depth = B._np_at(depth, np.uint32, na)
        pv = B._np_at(pv, np.float64, na * ld).reshape(na, ld)
        ps = B._np_at(ps, np.uint8, na * ld).reshape(na, ld)
        return depth, pv, ps

    def _in_cell(self, depth, pv, ps):
        ok = np.ones(len(self.rows), dtype=bool)
        for j in range(depth):
            side = np.where(self.rows[:, j % self.K] < pv[j], 0, 1)
            ok &= side == ps[j]
        return ok

    def _collect_cb(self, _ctx, ne, queries, dstar, n_local, same, first):
        try:
            K = self.K
            Q = B._np_at(queries, np.float64, ne * K).reshape(ne, K)
            ds = B._np_at(dstar, np.float64, ne)
            n_local = B._np_at(n_local, np.uint64, ne)
            same = B._np_at(same, np.uint8, ne)
            first = B._np_at(first, np.float64, ne * K).reshape(ne, K)
            self.tied = []
            for e in range(ne):
                t = np.flatnonzero(seq_sqdist(self.rows, Q[e], K) == ds[e]) if len(self.rows) else np.empty(0, dtype=np.int64)
                self.tied.append(t)
                n_local[e] = len(t)
                same[e] = 1
                if len(t):
                    first[e] = self.rows[t[0]]
                    same[e] = int(np.all(self.rows[t].view(np.uint64) == self.rows[t[0]].view(np.uint64)))
            self.calls["collect"] += 1
            return 0
        except Exception:
            import traceback
            traceback.print_exc()
            return -5

    def _first_cb(self, _ctx, na, ev, depth, pv, ps, ld, after, out):
        try:
            ev = B._np_at(ev, np.uint32, na)
            depth, pv, ps = self._cells(na, depth, pv, ps, ld)
            after = B._np_at(after, np.uint64, na)
            out = B._np_at(out, B.tie_first_dtype, na)
            seqs = np.arange(len(self.rows), dtype=np.uint64) + np.uint64(self.lo)
            for a in range(na):
                ok = self._in_cell(int(depth[a]), pv[a], ps[a])
                if after[a] != B.NONE:
                    ok &= seqs > after[a]
                hit = np.flatnonzero(ok)
                if len(hit) == 0:
                    out[a] = (B.NONE, B.NONE, 0.0, 0)
                else:
                    s = int(hit[0])
                    out[a] = (s + self.lo, s + self.lo, self.rows[s, int(depth[a]) % self.K], int(s in self.tied[ev[a]]))
            self.calls["first"] += 1
            return 0
        except Exception:
            import traceback
            traceback.print_exc()
            return -5

    def _split_cb(self, _ctx, na, ev, depth, pv, ps, ld, v, out):
        try:
            ev = B._np_at(ev, np.uint32, na)
            depth, pv, ps = self._cells(na, depth, pv, ps, ld)
            v = B._np_at(v, np.float64, na)
            out = B._np_at(out, B.tie_split_dtype, na)
            for a in range(na):
                t = self.tied[ev[a]]
                ok = self._in_cell(int(depth[a]), pv[a], ps[a])
                t = t[ok[t]]
                side = np.where(self.rows[t, int(depth[a]) % self.K] < v[a], 0, 1)
                for s in (0, 1):
                    m = t[side == s]
                    out[a]["n"][s] = len(m)
                    out[a]["min_seq"][s] = m[0] + self.lo if len(m) else B.NONE
                    out[a]["min_index"][s] = m[0] + self.lo if len(m) else B.NONE
            self.calls["split"] += 1
            return 0
        except Exception:
            import traceback
            traceback.print_exc()
            return -5
