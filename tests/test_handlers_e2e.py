"""HTTP-level parity through the reference's OWN handler code.

oracle/_ref/handler_driver_{ref,ours} (built by `make -C oracle handlers` in the build container)
contain the reference's unmodified src/*_handler.c driven in-process by tests/c/fake_http:
the same scripted requests (POST /vector, GET /vector, GET /compare/*, POST /nearest, PUT, DELETE)
go through the same handler code; only the L1 library underneath differs -- the reference's
vector_database.c + kdtree.c, or libsvdb_b200.so.  Every status and JSON body must be identical."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

from svdb import synth

REF_DRV = os.path.join(ROOT, "oracle", "_ref", "handler_driver_ref")
OUR_DRV = os.path.join(ROOT, "oracle", "_ref", "handler_driver_ours")


def make_script(path, n, D, seed, coarse):
    rng = np.random.Generator(np.random.PCG64(seed))
    gen = (lambda: synth.script_values(int(rng.integers(1, 1 << 30)), (D,))) if coarse else (lambda: rng.random(D) * 9 + 1)
    lines = []
    for i in range(n):
        lines.append("POST /vector - " + json.dumps({"uuid": f"00000000-0000-0000-0000-{i:012d}", "vector": list(gen())}))
    lines.append("GET /vector index=3 -")
    lines.append(f"GET /vector uuid=00000000-0000-0000-0000-{7:012d} -")
    lines.append("GET /vector index=99999 -")
    for _ in range(10):
        a, b = rng.integers(0, n, 2)
        for m in ("cosine_similarity", "euclidean_distance", "dot_product"):
            lines.append(f"GET /compare/{m} index1={a}&index2={b} -")
    lines.append(f"GET /compare/dot_product index1=2&index2={n + 5} -")
    for _ in range(25):
        lines.append("POST /nearest - " + json.dumps(list(gen())))
    for _ in range(6):
        lines.append(f"PUT /vector index={int(rng.integers(0, n))} " + json.dumps(list(gen())))
        lines.append(f"DELETE /vector index={int(rng.integers(0, n - 10))} -")
        for _ in range(4):
            lines.append("POST /nearest - " + json.dumps(list(gen())))
    lines.append("POST /nearest - [1,2,3]")
    lines.append("POST /nearest - not json")
    lines.append("POST /vector - " + json.dumps({"uuid": "short", "vector": [1.0, 2.0]}))
    lines.append("PATCH /vector - -")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    return len(lines)


def run(driver, script, kd_dim, D):
    return subprocess.run([driver, script, str(kd_dim), str(D)], capture_output=True, text=True, check=True, timeout=300).stdout


@pytest.mark.skipif(not os.path.exists(REF_DRV), reason="oracle/_ref/handler_driver_ref not built")
def test_reference_handlers_run_in_process(tmp_path):
    """CPU: the fake HTTP layer drives the reference's handlers over the reference's own library."""
    script = str(tmp_path / "s.txt")
    n = make_script(script, 40, 8, 1, False)
    out = run(REF_DRV, script, 3, 8).splitlines()
    assert len(out) == n + 1 and out[-1].startswith("final size=")
    assert out[0].split()[1] == "200" and '"index":3' in out[40]
    assert '"error"' in out[42]


@pytest.mark.gpu
@pytest.mark.parametrize("n,D,K,coarse,seed", [(120, 16, 3, True, 1), (200, 32, 32, False, 2), (150, 6, 2, True, 3)])
def test_http_level_parity_through_reference_handlers(tmp_path, n, D, K, coarse, seed):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not (os.path.exists(REF_DRV) and os.path.exists(OUR_DRV)):
        pytest.skip("oracle/_ref handler drivers were not shipped")
    script = str(tmp_path / "s.txt")
    nreq = make_script(script, n, D, seed, coarse)
    a = run(REF_DRV, script, K, D).splitlines()
    b = run(OUR_DRV, script, K, D).splitlines()
    assert len(a) == len(b) == nreq + 1
    diff = [(x, y) for x, y in zip(a, b) if x != y]
    assert not diff, diff[:3]
