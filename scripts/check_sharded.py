#!/usr/bin/env python
"""Multi-GPU parity check, run under torchrun with one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 scripts/check_sharded.py

Every rank owns a row-range shard; the merged answer of the sharded index (scan + NCCL
all-gather + K7 merge) must be identical -- sequence numbers and distance bits -- to a
single-engine scan of all rows (done on rank 0's GPU) for top-1 and top-10 -- including on lattice
data where distinct points tie exactly and the winner is the one the reference's tree reaches first
(the single engine keeps that tree; the shards walk its path together).  Prints one JSON line from rank 0.
"""
from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "simple-vector-db_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from svdb import binding as B  # noqa: E402
from svdb.sharded import ReplicatedCompare, ShardedIndex  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    ok = True
    report = []
    # the last three are coarse lattices: distinct kd-points at exactly equal distance everywhere (and, with
    # 0/1 coordinates, dozens of them per query), so the answer depends on the order of the reference's tree
    for (n, D, K, levels, replicate) in ((400_003, 128, 128, 0, True), (50_000, 768, 768, 0, True), (300_000, 16, 3, 0, True),
                                         (200_000, 16, 3, 7, True), (200_000, 16, 3, 7, False), (100_000, 64, 64, 2, True),
                                         (60_000, 768, 768, 2, True)):
        g = torch.Generator().manual_seed(n + levels)
        if levels:
            rows = torch.randint(0, levels, (n, D), generator=g).to(torch.float64) / 2      # same on every rank
            Q = torch.randint(0, levels, (16, D), generator=g).to(torch.float64) / 2
            Q[:, 0] += 0.25                       # between lattice planes: no exact hit, the minimum is shared
            Q = Q.pin_memory()
        else:
            rows = torch.rand((n, D), dtype=torch.float64, generator=g)
            Q = torch.rand((16, D), dtype=torch.float64, generator=g).pin_memory()
        idx = ShardedIndex(D, K, n, rank, world, local, exchange=os.environ.get("SVDB_EXCHANGE", "p2p"),
                           replicate_thin=replicate)
        idx.bind_current_stream()
        idx.ingest_device(rows[idx.lo:idx.hi].to(dev).contiguous())
        ref = None
        if rank == 0:
            ref = B.Engine(D, K, device=local)
            ref.insert_device(rows.to(dev).data_ptr(), n, D)
        gm = torch.Generator().manual_seed(99 + n)
        for phase in ("bulk", "after inserts+updates"):
            if phase != "bulk":
                # vector_db_insert x 5 and vector_db_update x 7 (stale kd-points stay searchable): the log grows at its tail
                mk = (lambda m: torch.randint(0, levels, (m, D), generator=gm).to(torch.float64) / 2) if levels else \
                     (lambda m: torch.rand((m, D), dtype=torch.float64, generator=gm))
                new_rows, upd_rows = mk(5).numpy(), mk(7).numpy()
                upd_ids = torch.randint(0, n, (7,), generator=gm).numpy().astype(np.uint64)
                first = idx.append_rows(new_rows)
                idx.append_update(upd_ids, upd_rows)
                if rank == 0:
                    assert ref.insert(new_rows) == first
                    ref.update(upd_ids, upd_rows)
                Q = torch.cat([Q[:9], torch.from_numpy(new_rows[:3]), torch.from_numpy(upd_rows[:4])]).pin_memory()
            for k in (1, 10):
                got = idx.nearest(Q, k)
                few = idx.nearest(Q[:3], k)       # too few queries to split over replicas / one small pass per shard
                # ONE query per call: the whole step is a single scan launch whose last CTA re-ranks, stores to the peers,
                # waits and merges (scan_tail) -- through the host entry point and through the device entry point
                ones = [idx.nearest(Q[i:i + 1], k) for i in range(4)]
                # TWO queries per call: where the kd-points are wide enough the pair shares one scan (K13, NQ = 2) whose tail
                # exchanges and merges both answers
                pairs = [idx.nearest(Q[i:i + 2], k) for i in (0, 2)]
                qd = Q.to(dev)
                pairs_dev = []
                for i in (0, 2):
                    r = idx.nearest_device(qd[i:i + 2], k).clone()
                    torch.cuda.synchronize()
                    pairs_dev.append(r.cpu().numpy().view(B.candidate_dtype).reshape(2, k))
                ones_dev = []
                for i in range(4):
                    r = idx.nearest_device(qd[i:i + 1], k).clone()
                    torch.cuda.synchronize()
                    ones_dev.append(r.cpu().numpy().view(B.candidate_dtype).reshape(1, k))
                if rank == 0:
                    widx, wdist, wseq = ref.nearest(Q.numpy(), k)
                    same = np.array_equal(got["seq"], wseq) and np.array_equal(got["dist"].view(np.uint64), wdist.view(np.uint64))
                    same = same and np.array_equal(got["index"], widx)
                    same = same and np.array_equal(few["seq"], wseq[:3]) and np.array_equal(few["dist"].view(np.uint64), wdist[:3].view(np.uint64))
                    single = all(np.array_equal(o["seq"][0], wseq[i]) and np.array_equal(o["index"][0], widx[i]) and
                                 np.array_equal(o["dist"][0].view(np.uint64), wdist[i].view(np.uint64)) for i, o in enumerate(ones))
                    # the device entry point leaves exact ties between distinct points (SVDB_CAND_TIE) and answers it could not
                    # prove complete (SVDB_CAND_UNSAFE: rerun with SVDB_MODE_FP64 / EXACT) to its caller
                    single_dev = all((o["flags"][0, 0] & (B.CAND_TIE | B.CAND_UNSAFE)) or
                                     (np.array_equal(o["seq"][0], wseq[i]) and np.array_equal(o["dist"][0].view(np.uint64), wdist[i].view(np.uint64)))
                                     for i, o in enumerate(ones_dev))
                    pair_ok = all(np.array_equal(o["seq"], wseq[i:i + 2]) and np.array_equal(o["index"], widx[i:i + 2]) and
                                  np.array_equal(o["dist"].view(np.uint64), wdist[i:i + 2].view(np.uint64)) for i, o in zip((0, 2), pairs))
                    pair_ok = pair_ok and all(all((o["flags"][j, 0] & (B.CAND_TIE | B.CAND_UNSAFE)) or
                                                  (np.array_equal(o["seq"][j], wseq[i + j]) and
                                                   np.array_equal(o["dist"][j].view(np.uint64), wdist[i + j].view(np.uint64))) for j in range(2))
                                              for i, o in zip((0, 2), pairs_dev))
                    same = same and single and single_dev and pair_ok
                    # ... and against the CPU oracle itself (flat (distance, seq) order is the reference's answer wherever
                    # the minimum is not shared by distinct points: the non-lattice cases)
                    oracle_ok = None
                    if not levels and phase == "bulk" and n * K <= 60_000_000:
                        from oracle import binding as OB
                        from oracle.binding import PortDB
                        db = PortDB(OB.load_port(), D, K)
                        for r in rows.numpy():
                            db.insert(r)
                        oracle_ok = True
                        for i, q in enumerate(Q.numpy()):
                            oseq, oidx, od = db.topk(q, k)
                            oracle_ok &= np.array_equal(got["index"][i].astype(np.uint64), np.asarray(oidx, dtype=np.uint64)) and \
                                np.array_equal(got["dist"][i].view(np.uint64), od.view(np.uint64))
                        db.close()
                        same = same and bool(oracle_ok)
                    ok &= bool(same)
                    report.append({"rows": n, "dim": D, "kd_dim": K, "k": k, "lattice_levels": levels, "phase": phase,
                                   "layout": "replicated kd log, queries split" if idx.replicated else "row shards",
                                   "identical_to_single_gpu": bool(same), "single_query_calls_identical": bool(single and single_dev), "two_query_calls_identical": bool(pair_ok),
                                   "identical_to_cpu_oracle": oracle_ok, "tie_events": idx.engine.stats()["tie_events"],
                                   "tie_levels": idx.engine.stats()["tie_levels"]})
        if ref is not None:
            ref.close()
        idx.close()
    # /compare: replicas, pairs split over the ranks, results all-gathered
    n, D = 20_000, 256
    g = torch.Generator().manual_seed(7)
    rows = torch.rand((n, D), dtype=torch.float64, generator=g).to(dev)
    i1 = torch.randint(0, n, (5001,), generator=g).to(dev)
    i2 = torch.randint(0, n, (5001,), generator=g).to(dev)
    with B.Engine(D, 1, device=local, flags=B.FLAG_NO_LOG) as e:
        e.set_stream(torch.cuda.current_stream().cuda_stream)
        e.insert_device(rows.data_ptr(), n, D)
        rc = ReplicatedCompare(e, rank, world)
        for metric in (0, 1, 2, 3):
            got = rc.compare(metric, i1, i2)
            full = torch.empty_like(got)
            e.compare_device(metric, i1.data_ptr(), i2.data_ptr(), i1.numel(), full.data_ptr())
            torch.cuda.synchronize()
            same = bool(torch.equal(got.view(torch.int32), full.view(torch.int32)))
            ok &= same
            if rank == 0:
                report.append({"compare_metric": metric, "pairs": i1.numel(), "identical_to_single_gpu": same})
    dist.barrier()
    if rank == 0:
        line = json.dumps({"check": "sharded_vs_single", "world": world, "ok": ok, "cases": report})
        print(line, flush=True)
        if os.environ.get("SVDB_CHECK_OUT"):      # stdout also carries NCCL's version banner
            with open(os.environ["SVDB_CHECK_OUT"], "w") as f:
                f.write(line + "\n")
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
