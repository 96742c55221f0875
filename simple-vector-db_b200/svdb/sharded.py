"""Row-range sharding of one store over the GPUs of a box.

One process per GPU (torch.distributed, NCCL over NVLink).  Rank g owns the contiguous
slice [lo, hi) of the log's sequence numbers; a query is broadcast, every rank scans its
shard (K1) and re-ranks its survivors, the per-rank top-k candidate blocks (k x 32 bytes per
query) are exchanged with ONE all-gather, and every rank merges them (K7).  The exchange is
latency-bound (16..320 bytes per rank and query), so it rides the same stream directly behind
the scan epilogue.  The key (reference-order distance, global sequence number) is exact, so
the merged answer is identical to a single-GPU scan of all rows.  When distinct kd-points tie at
exactly the minimal distance, the shards additionally walk the path of the reference's GLOBAL tree
together (csrc/tie_protocol.cu) so that position 0 is the entry the reference itself returns.

Thin kd-points (kd_dim <= 8, the reference's default is 3) are different: the tree answers a query
in O(log N) visits, so scanning row shards would make N GPUs SLOWER than one.  There the 8*kd_dim
bytes per row that /nearest looks at are replicated on every GPU (all-gathered once at ingest: the
whole log + the reference-shaped tree over ALL rows), and the QUERIES are split over the ranks:
no collective in the query path except the all-gather of the answers, answers identical to one
GPU by construction, throughput scales with the ranks ("replicas").
"""
from __future__ import annotations

import numpy as np

from . import binding as B


def shard_range(n_rows: int, world: int, rank: int):
    """Contiguous ceil-split: shard g owns [g*ceil(n/G), min(n, (g+1)*ceil(n/G)))."""
    per = -(-n_rows // world)
    lo = min(n_rows, rank * per)
    return lo, min(n_rows, lo + per)


def query_span(nq: int, world: int, rank: int):
    """Replicated stores split the QUERIES: rank r answers [lo, hi) of a call, `per` = slots per rank in the
    all-gather of the answers.  Calls with fewer than 2 x world queries are answered by every replica itself
    (per == 0: nothing is exchanged)."""
    if nq < 2 * world:
        return 0, nq, 0
    per = -(-nq // world)
    return min(nq, rank * per), min(nq, (rank + 1) * per), per


def merge_candidates_host(gathered: np.ndarray, k: int) -> np.ndarray:
    """Host restatement of the K7 merge for a gathered [nshards, nq, k] candidate array.

    Used where the payload is already on the host (tests over gloo, the exact-rerun path);
    the product's data path uses the CUDA kernel (svdb_merge_candidates_device)."""
    nshards, nq, kk = gathered.shape
    out = np.empty((nq, k), dtype=B.candidate_dtype)
    for q in range(nq):
        c = gathered[:, q, :].reshape(-1)
        order = np.lexsort((c["seq"], c["dist"]))[:k]
        out[q] = c[order]
        flags = np.bitwise_or.reduce(c["flags"] & ~np.uint64(B.CAND_TIE))
        # SVDB_CAND_TIE is about the merged minimum: two entries there, or one whose shard flagged it
        if out[q]["seq"][0] != B.NONE:
            at_min = (c["seq"] != B.NONE) & (c["dist"] == out[q]["dist"][0])
            if at_min.sum() >= 2 or np.any(c["flags"][at_min] & np.uint64(B.CAND_TIE)):
                flags |= np.uint64(B.CAND_TIE)
        out[q]["flags"] = flags
    return out


class ShardedIndex:
    """One rank's shard plus the exchange.  world == 1 needs no process group."""

    REPLICATE_MAX_KD = 8        # the engine's tree regime (option nearest.tree_max_k)

    def __init__(self, dimension: int, kd_dim: int, n_rows_total: int, rank: int = 0, world: int = 1,
                 device: int = 0, group=None, exchange: str = "p2p", max_records: int = 32768,
                 replicate_thin: bool = True):
        import torch
        self.torch = torch
        self.rank, self.world, self.device, self.group = rank, world, device, group
        self.D, self.K = dimension, kd_dim
        self.n_total = n_rows_total
        self.lo, self.hi = shard_range(n_rows_total, world, rank)
        self.replicated = world > 1 and replicate_thin and kd_dim <= self.REPLICATE_MAX_KD
        self.merge_launches = 0
        self._bufs = {}
        self._pending, self._sealed = [], False
        self.xch = None
        self.max_records = max_records
        if self.replicated:
            # every rank: the WHOLE kd log (kd_dim doubles per row) and the tree over it; queries are split
            self.engine = B.Engine(kd_dim, kd_dim, device=device, reserve_rows=max(1, n_rows_total),
                                   flags=B.FLAG_LOG_ONLY)
            return
        self.engine = B.Engine(dimension, kd_dim, device=device, seq_base=self.lo,
                               reserve_rows=max(1, self.hi - self.lo),
                               flags=B.FLAG_SHARD if world > 1 else 0)
        self.engine.set_option("log.index_base", self.lo)      # merged answers carry global row numbers
        # "p2p": candidates are stored into the peers' HBM over NVLink and merged in the same launch
        # (svdb_exchange); "nccl": all_gather_into_tensor + svdb_merge_candidates_device
        if world > 1 and exchange == "p2p":
            self.xch = B.Exchange(device, rank, world, max_records)
            mine = torch.frombuffer(bytearray(self.xch.handle), dtype=torch.uint8).to(torch.device("cuda", device))
            allh = torch.empty(64 * world, dtype=torch.uint8, device=mine.device)
            torch.distributed.all_gather_into_tensor(allh, mine, group=group)
            self.xch.connect(bytes(allh.cpu().numpy().tobytes()))
            torch.distributed.barrier(group=group)

    def close(self):
        if self.xch is not None:
            self.torch.cuda.synchronize(self.device)
            if self.world > 1:
                self.torch.distributed.barrier(group=self.group)   # nobody unmaps while a peer may still store
            self.xch.close()
            self.xch = None
        self.engine.close()

    def bind_current_stream(self):
        self.engine.set_stream(self.torch.cuda.current_stream(self.device).cuda_stream)

    def ingest_device(self, rows) -> None:
        """rows: contiguous float64 CUDA tensor [n, D] belonging to this shard, in order."""
        assert rows.dtype == self.torch.float64 and rows.is_contiguous() and rows.shape[1] == self.D
        if self.replicated:
            self._pending.append(rows[:, :self.K].contiguous())     # exchanged when the shard is complete
            if sum(p.shape[0] for p in self._pending) >= self.hi - self.lo:
                self._seal()
            return
        self.engine.insert_device(rows.data_ptr(), rows.shape[0], rows.shape[1])

    # -- the log after the bulk ingest: inserts and updates append to its END ------------------
    def append_rows(self, rows) -> int:
        """vector_db_insert for a sharded store (collective: every rank passes the same rows, a float64 numpy
        array [m, D]).  New rows get the next global row numbers and go to the TAIL of the log, i.e. to the last
        shard (sequence numbers stay global and ordered); replicas append them everywhere.  Returns the first
        new row number."""
        rows = np.ascontiguousarray(rows, dtype=np.float64).reshape(-1, self.D)
        first = self.n_total
        index = np.arange(first, first + len(rows), dtype=np.uint64)
        self._append_log(rows, index, new_rows=True)
        self.n_total += len(rows)
        return first

    def append_update(self, index, rows) -> None:
        """vector_db_update: the reference re-inserts the updated kd-point with the SAME index and never removes
        the old one (vector_database.c:174) -- one more log entry at the tail per update (collective)."""
        rows = np.ascontiguousarray(rows, dtype=np.float64).reshape(-1, self.D)
        index = np.ascontiguousarray(index, dtype=np.uint64).reshape(-1)
        keep = index < self.n_total                     # :171 out of range: silent no-op
        self._append_log(rows[keep], index[keep], new_rows=False)

    def _append_log(self, rows, index, new_rows: bool) -> None:
        if len(rows) == 0:
            return
        if self.replicated:
            if not self._sealed:
                self._seal()
            self.engine.append_kdpoints(rows[:, :self.K], index)
        elif self.world == 1:
            if new_rows:
                self.engine.insert(rows)
            else:
                self.engine.update(index, rows)
        elif self.rank == self.world - 1:
            # the tail shard: its sequence numbers are the largest, so global log order is preserved
            if new_rows and int(index[0]) == self.lo + self.engine.size:
                self.engine.insert(rows)                # keeps the shard's rows indexable as well
            else:
                self.engine.append_kdpoints(rows[:, :self.K], index)      # reported verbatim: global row numbers

    def _seal(self) -> None:
        """Replicated mode: all-gather the kd-points of every shard and append them in global row order."""
        t = self.torch
        dev = t.device("cuda", self.device)
        mine = t.cat(self._pending) if self._pending else t.empty((0, self.K), dtype=t.float64, device=dev)
        self._pending, self._sealed = [], True
        assert mine.shape[0] == self.hi - self.lo, "a shard must be ingested completely before the first query"
        per = -(-self.n_total // self.world)
        padded = t.zeros((per, self.K), dtype=t.float64, device=dev)
        padded[:mine.shape[0]] = mine
        everything = t.empty((self.world, per, self.K), dtype=t.float64, device=dev)
        t.distributed.all_gather_into_tensor(everything, padded, group=self.group)
        t.cuda.current_stream(self.device).synchronize()
        for r in range(self.world):
            lo, hi = shard_range(self.n_total, self.world, r)
            if hi > lo:
                self.engine.append_kdpoints_device(everything[r].data_ptr(), lo, hi - lo, self.K)
        self.engine.flush()

    def _query_span(self, nq: int):
        """Replicated mode: which queries this rank answers (all of them when there are too few to split)."""
        return query_span(nq, self.world, self.rank)

    def _buffers(self, nq: int, k: int):
        key = (nq, k)
        if key not in self._bufs:
            t = self.torch
            dev = t.device("cuda", self.device)
            local = t.zeros((nq, k, 4), dtype=t.int64, device=dev)
            gathered = t.zeros((self.world, nq, k, 4), dtype=t.int64, device=dev) if self.world > 1 else None
            merged = t.zeros((nq, k, 4), dtype=t.int64, device=dev) if self.world > 1 else local
            host = t.zeros((nq, k, 4), dtype=t.int64).pin_memory()
            self._bufs[key] = (local, gathered, merged, host)
        return self._bufs[key]

    def nearest_device(self, dq, k: int, mode: int = B.MODE_AUTO):
        """dq: float64 CUDA tensor [nq, >=K]. Returns the merged [nq, k, 4] int64 CUDA tensor
        (a view of svdb_candidate records); everything is enqueued on the current stream."""
        nq = dq.shape[0]
        if self.replicated:
            if not self._sealed:
                self._seal()                      # a rank whose shard is empty never saw an ingest call
            return self._nearest_device_replicated(dq, k, mode)
        local, gathered, merged, _ = self._buffers(nq, k)
        if self.world > 1 and self.xch is not None and nq * k <= self.max_records:
            # scan + peer-memory exchange + merge inside the library (one launch for a single query)
            self.engine.nearest_device_sharded(self.xch, dq.data_ptr(), nq, dq.stride(0), k, merged.data_ptr(), mode)
            return merged
        self.engine.nearest_device(dq.data_ptr(), nq, dq.stride(0), k, local.data_ptr(), mode)
        if self.world > 1:
            stream = self.torch.cuda.current_stream(self.device).cuda_stream
            self.torch.distributed.all_gather_into_tensor(gathered, local, group=self.group)
            B.merge_candidates_device(self.device, stream, gathered.data_ptr(), self.world, nq, k, merged.data_ptr())
            self.merge_launches += 1
        return merged

    def _nearest_device_replicated(self, dq, k: int, mode: int):
        t = self.torch
        nq = dq.shape[0]
        lo, hi, per = self._query_span(nq)
        key = ("rep", nq, k)
        if key not in self._bufs:
            dev = t.device("cuda", self.device)
            mine = t.zeros((per if per else nq, k, 4), dtype=t.int64, device=dev)
            everything = t.zeros((self.world, per, k, 4), dtype=t.int64, device=dev) if per else None
            self._bufs[key] = (mine, everything)
        mine, everything = self._bufs[key]
        if hi > lo:
            self.engine.nearest_device(dq[lo:hi].data_ptr(), hi - lo, dq.stride(0), k, mine.data_ptr(), mode)
        if per == 0:
            return mine[:nq]
        t.distributed.all_gather_into_tensor(everything, mine, group=self.group)
        return everything.view(self.world * per, k, 4)[:nq]

    def nearest(self, q_host, k: int):
        """End-to-end call a user makes: q_host is a float64 CPU tensor (or numpy array) [nq, >=K];
        returns a numpy array of svdb_candidate [nq, k] with the MERGED answers (`seq` is the global
        row for insert-only data).  Host query in, host result out.
        With the peer-memory exchange this is ONE C-ABI call (svdb_nearest_batch_sharded): scan, exchange
        and merge are enqueued -- and from the second call of a shape on replayed as one CUDA graph --
        inside the library.  With exchange="nccl" the steps are driven from here."""
        nq = q_host.shape[0]
        if self.replicated and not self._sealed:
            self._seal()
        if self.replicated and self._query_span(nq)[2] == 0:
            # too few queries to split: every replica answers them itself, nothing is exchanged
            q_np = q_host.numpy() if hasattr(q_host, "numpy") else np.asarray(q_host)
            idx, dist, seq = self.engine.nearest(q_np[:, :self.K], k)
            res = np.zeros((nq, k), dtype=B.candidate_dtype)
            res["index"], res["dist"], res["seq"] = idx, dist, seq
            return res
        if self.xch is not None and nq * k <= self.max_records:
            q_np = q_host.numpy() if hasattr(q_host, "numpy") else np.asarray(q_host)
            idx, dist, seq = self.engine.nearest_sharded(self.xch, q_np, k)
            res = np.zeros((nq, k), dtype=B.candidate_dtype)
            res["index"], res["dist"], res["seq"] = idx, dist, seq
            return res
        t = self.torch
        if not hasattr(q_host, "to"):
            q_host = t.from_numpy(np.ascontiguousarray(q_host))
        dq = q_host.to(t.device("cuda", self.device), non_blocking=True)
        merged = self.nearest_device(dq, k)
        if self.replicated:
            if ("host", nq, k) not in self._bufs:
                self._bufs[("host", nq, k)] = t.zeros((nq, k, 4), dtype=t.int64).pin_memory()
            host = self._bufs[("host", nq, k)]
        else:
            host = self._buffers(dq.shape[0], k)[3]
        host.copy_(merged, non_blocking=True)
        t.cuda.current_stream(self.device).synchronize()
        res = host.numpy().view(B.candidate_dtype).reshape(dq.shape[0], k)
        if np.any(res["flags"] & B.CAND_UNSAFE):
            # every rank sees the same merged flags, so every rank takes this branch together
            merged = self.nearest_device(dq, k, mode=B.MODE_EXACT)
            host.copy_(merged, non_blocking=True)
            t.cuda.current_stream(self.device).synchronize()
            res = host.numpy().view(B.candidate_dtype).reshape(dq.shape[0], k)
        res = res.copy()
        if self.world > 1 and np.any(res["flags"][:, 0] & B.CAND_TIE):
            # distinct kd-points at exactly the minimal distance: all shards walk the global tree's path
            # together (svdb_resolve_ties_sharded); the flags are the same on every rank
            q_np = np.ascontiguousarray(q_host.numpy() if hasattr(q_host, "numpy") else q_host)
            self.engine.resolve_ties_sharded(self.rank, self.world, q_np, res, allgather=self._allgather_host)
        return res

    def _allgather_host(self, send: np.ndarray, recv: np.ndarray) -> None:
        """Host bytes in, host bytes of every rank out, over the process group (NCCL moves device tensors)."""
        t = self.torch
        dev = t.device("cuda", self.device)
        mine = t.from_numpy(np.array(send, copy=True)).to(dev)
        out = t.empty(self.world * mine.numel(), dtype=t.uint8, device=dev)
        t.distributed.all_gather_into_tensor(out, mine, group=self.group)
        recv[:] = out.cpu().numpy()


class ReplicatedCompare:
    """/compare over N GPUs: pairs are independent, so every rank holds a replica of
    the rows and answers a contiguous slice of the pairs; one all-gather returns the whole result.
    No exchange inside the kernel -- this is the "replicas, no collective" case."""

    def __init__(self, engine: "B.Engine", rank: int = 0, world: int = 1, group=None):
        import torch
        self.torch, self.e, self.rank, self.world, self.group = torch, engine, rank, world, group

    def compare(self, metric: int, i1, i2):
        """i1, i2: int64 CUDA tensors of n pair indices (same on every rank). Returns float32 [n] (or
        [n, 3] for metric 3) on every rank."""
        t = self.torch
        n = i1.numel()
        per = -(-n // self.world)
        lo, hi = min(n, self.rank * per), min(n, (self.rank + 1) * per)
        width = 3 if metric == B.ALL_METRICS else 1
        mine = t.full((per, width), -1.0, dtype=t.float32, device=i1.device)
        if hi > lo:
            a, b = i1[lo:hi].contiguous(), i2[lo:hi].contiguous()
            self.e.compare_device(metric, a.data_ptr(), b.data_ptr(), hi - lo, mine.data_ptr())
        if self.world == 1:
            out = mine
        else:
            out = t.empty((self.world * per, width), dtype=t.float32, device=i1.device)
            t.distributed.all_gather_into_tensor(out, mine, group=self.group)
        out = out[:n]
        return out if width == 3 else out[:, 0]
