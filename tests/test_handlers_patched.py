"""SURVEY.md s8f rows 3 and 4 as an ENDPOINT: integration/f3_f4_handlers.patch adds `/nearest?number=k`, a batched
`POST /compare/<metric>` and the O(D) JSON walk to the reference's own handler code (src/compare_handler.c:263-441,
README.md:287).  oracle/_ref/handler_driver_patched_{ref,ours} are the patched handlers driven in-process by
tests/c/fake_http, over the reference's L1 code and over libsvdb_b200.so (built by `make -C oracle handlers_patched`
where /root/reference exists; the binaries travel to the GPU box).

  * k = 1 / no `number`: byte-identical to the UNPATCHED handlers, on both L1s;
  * k > 1 (drop-in only; the reference's L1 has no top-k and answers 501): the oracle's (distance, seq) order;
  * batched compare: the same values as one GET per pair, identical on both L1s.
"""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

from svdb import synth
from test_handlers_e2e import REF_DRV, OUR_DRV, make_script, run

PREF_DRV = os.path.join(ROOT, "oracle", "_ref", "handler_driver_patched_ref")
POUR_DRV = os.path.join(ROOT, "oracle", "_ref", "handler_driver_patched_ours")

needs_ref = pytest.mark.skipif(not (os.path.exists(REF_DRV) and os.path.exists(PREF_DRV)),
                               reason="oracle/_ref handler drivers not built")


def body(line):
    return json.loads(line.split(" ", 2)[2])


def extra_script(path, base_script, n, D, seed, ks=(1,)):
    """base requests + `number=` queries + batched compares (and the per-pair GETs they must agree with)"""
    rng = np.random.Generator(np.random.PCG64(seed + 100))
    lines = open(base_script).read().splitlines()
    lines = [ln for ln in lines if ln.startswith("POST /vector")][:n]           # inserts only: indices stay put
    queries = [list(rng.random(D) * 9 + 1) for _ in range(6)]
    marks = {"nearest": [], "batch": [], "single": []}
    for q in queries:
        for k in ks:
            marks["nearest"].append((len(lines), k, q))
            lines.append(f"POST /nearest number={k} " + json.dumps(q))
    pairs = [[int(a), int(b)] for a, b in rng.integers(0, n, (12, 2))] + [[0, n + 3]]
    for m in ("cosine_similarity", "euclidean_distance", "dot_product"):
        marks["batch"].append((len(lines), m, pairs))
        lines.append(f"POST /compare/{m} - " + json.dumps(pairs))
        for a, b in pairs[:-1]:
            marks["single"].append((len(lines), m))
            lines.append(f"GET /compare/{m} index1={a}&index2={b} -")
    lines.append("POST /compare/dot_product - [[1,2,3]]")
    lines.append("POST /compare/dot_product - {\"a\": 1}")
    lines.append("POST /compare/manhattan - [[0,1]]")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    return marks


@needs_ref
def test_patched_handlers_k1_identical_to_unpatched_on_reference_l1(tmp_path):
    """CPU: the patch changes nothing for requests the unpatched server understands."""
    script = str(tmp_path / "s.txt")
    make_script(script, 60, 12, 5, True)
    assert run(REF_DRV, script, 3, 12) == run(PREF_DRV, script, 3, 12)
    # `number=1` and `number=0` take the old path too
    with open(script, "a") as f:
        f.write("POST /nearest number=1 " + json.dumps([1.5] * 12) + "\n")
        f.write("POST /nearest number=0 " + json.dumps([1.5] * 12) + "\n")
        f.write("POST /nearest - " + json.dumps([1.5] * 12) + "\n")
    out = run(PREF_DRV, script, 3, 12).splitlines()
    assert out[-4].split(" ", 1)[1] == out[-3].split(" ", 1)[1] == out[-2].split(" ", 1)[1]


@needs_ref
def test_patched_handlers_batched_compare_and_topk_stub_on_reference_l1(tmp_path):
    """CPU: batched compare == one GET per pair; number > 1 needs the batched L1 (501 on the reference's)."""
    base, script = str(tmp_path / "b.txt"), str(tmp_path / "s.txt")
    make_script(base, 40, 8, 6, False)
    marks = extra_script(script, base, 40, 8, 6, ks=(1, 3))
    out = run(PREF_DRV, script, 3, 8).splitlines()
    singles = iter(marks["single"])
    for line_no, metric, pairs in marks["batch"]:
        vals = body(out[line_no])[metric]
        assert len(vals) == len(pairs) and vals[-1] == -1          # out-of-range pair: the metrics' own sentinel
        for v in vals[:-1]:
            ln, m = next(singles)
            assert m == metric and body(out[ln])[metric] == v
    for line_no, k, _ in marks["nearest"]:
        status = out[line_no].split()[1]
        assert status == ("200" if k == 1 else "501")
    assert [ln.split()[1] for ln in out[-4:-1]] == ["400", "400", "400"]


@pytest.mark.gpu
@pytest.mark.parametrize("n,D,K,seed", [(150, 16, 3, 1), (300, 32, 32, 2)])
def test_patched_handlers_on_the_dropin(tmp_path, port, n, D, K, seed):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not all(os.path.exists(p) for p in (REF_DRV, OUR_DRV, PREF_DRV, POUR_DRV)):
        pytest.skip("oracle/_ref handler drivers were not shipped")
    base, script = str(tmp_path / "b.txt"), str(tmp_path / "s.txt")
    nreq = make_script(base, n, D, seed, False)
    # 1. everything the unpatched server understands: identical bytes, patched or not, on the drop-in
    assert run(OUR_DRV, base, K, D) == run(POUR_DRV, base, K, D) == run(PREF_DRV, base, K, D)
    # 2. the new endpoints
    marks = extra_script(script, base, n, D, seed, ks=(1, 2, 10))
    ours = run(POUR_DRV, script, K, D).splitlines()
    ref = run(PREF_DRV, script, K, D).splitlines()
    rows = np.array([json.loads(ln.split(" ", 3)[3])["vector"] for ln in open(script).read().splitlines()[:n]])
    from test_gpu_parity import oracle_topk
    for line_no, k, q in marks["nearest"]:
        if k == 1:
            assert ours[line_no] == ref[line_no]                   # same single answer, same bytes
            continue
        (wseq, widx, wd), = oracle_topk(port, rows, K, np.array([q]), k)
        got = body(ours[line_no])["neighbors"]
        assert [g["index"] for g in got] == [int(i) for i in widx]
        assert [np.float64(g["distance"]).view(np.uint64) for g in got] == [d.view(np.uint64) for d in wd]
        assert all(g["vector"] == list(rows[g["index"]]) for g in got)
    for line_no, metric, pairs in marks["batch"]:
        assert ours[line_no] == ref[line_no]                       # batched L1 call == the reference's per-pair loop
    assert ours[-4:] == ref[-4:]
