"""Two engines on two GPUs in ONE process (the opt-in to > 48 KB of dynamic shared memory is per device; a
process-wide cache of it made every wide scan of the second device fail with cudaErrorInvalidValue)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

from svdb import binding as B  # noqa: E402
from svdb import synth  # noqa: E402
from test_gpu_parity import assert_topk_equal, oracle_topk  # noqa: E402


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process")
def test_engines_on_two_devices_in_one_process(port):
    D, k = 768, 10                                            # 6 KB rows: every scan needs the opt-in
    rows = synth.uniform_rows(1, 3000, D)
    Q = synth.uniform_rows(2, 3, D)
    Qb = synth.uniform_rows(3, 70, D)
    want, want_b = oracle_topk(port, rows, D, Q, k), oracle_topk(port, rows, D, Qb, k)
    with B.Engine(D, D, device=0) as e0, B.Engine(D, D, device=1) as e1:
        for e in (e0, e1):
            e.insert(rows)
        for plane in (2, 1, 0):
            for e in (e0, e1, e0):                            # device 0 first, then 1, then 0 again
                e.set_option("scan.plane", plane)
                assert_topk_equal(e.nearest(Q, k), want, k)
        for e in (e0, e1):                                    # K2 / K10 and compare as well
            assert_topk_equal(e.nearest(Qb, k), want_b, k)
            e.set_option("nearest.umma_min_queries", 0)
            assert_topk_equal(e.nearest(Qb[:20], k), want_b[:20], k)
            i1, i2 = synth.index_pairs(4, 64, 3000)
            got = e.compare(B.ALL_METRICS, i1, i2)
            assert got.shape == (64, 3) and np.all(np.isfinite(got))
