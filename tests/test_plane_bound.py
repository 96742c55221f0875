"""The square-root-form key error bound of the single-plane scans (csrc/plane_scan.cu, csrc/tail.cuh: FinalArgs::sq_mode),
checked in numpy.  K12 forms key = |x^ - fl32(q)|^2 in fp32 from the bf16 hi plane x^ = bf16(fl32(x)); K13 forms
key = |x^ - q^|^2 EXACTLY (integer dot products) from a one-byte plane x^ = lo + step * u and a 16-bit query grid.  Both hand
finalize   | sqrt(key) - sqrt(d) | <= E + gamma (sqrt(d) + E),   E = max_r |x_r - x^_r|_2 + |q - q^|_2,
from which it derives the re-rank window and the completeness proof.  A bound that were too small would mean silently
wrong answers, so the inequality AND the two decisions finalize takes from it are pinned here, on benign and on
adversarial inputs (the CUDA paths are pinned against the oracle in tests/test_gpu_shadow_scan.py)."""
import numpy as np
import pytest

from test_umma_bound import split


def make(kind, n, nq, K, rng):
    if kind == "uniform":
        return rng.random((n, K)), rng.random((nq, K))
    if kind == "normal":
        return rng.standard_normal((n, K)), rng.standard_normal((nq, K))
    if kind == "near_query":
        q = rng.random((nq, K))
        return np.repeat(q, n // nq, axis=0) + 1e-3 * rng.standard_normal((n, K)), q
    if kind == "offset":
        return 1000.0 + rng.random((n, K)), 1000.0 + rng.random((nq, K))
    if kind == "cauchy":
        return rng.standard_cauchy((n, K)).clip(-1e3, 1e3), rng.standard_cauchy((nq, K)).clip(-1e3, 1e3)
    if kind == "sparse":
        return (np.where(rng.random((n, K)) < 0.05, rng.standard_normal((n, K)) * 100, 0.0),
                np.where(rng.random((nq, K)) < 0.05, rng.standard_normal((nq, K)) * 100, 0.0))
    if kind == "small_integers":
        return rng.integers(-3, 4, (n, K)).astype(float), rng.integers(-3, 4, (nq, K)).astype(float)
    if kind == "query_outside":      # queries far outside the range the plane's grid covers
        return rng.random((n, K)), 3.0 + 5.0 * rng.random((nq, K))
    mag = 10.0 ** rng.integers(-3, 3, size=K).astype(float)
    return rng.standard_normal((n, K)) * mag, rng.standard_normal((nq, K)) * mag


KINDS = ["uniform", "normal", "near_query", "offset", "cauchy", "sparse", "small_integers", "query_outside", "mixed_magnitudes"]


# ---- K12: bf16 hi plane, fp32 arithmetic ---------------------------------------------------------------------------
def k12_keys(x, q):
    """fp32 emulation of scan_plane_kernel: lane l owns coordinates trip * 256 + l * 8 .. + 7, two FMA chains per slice."""
    n, K = x.shape
    Kp = -(-K // 64) * 64
    trips = -(-Kp // 256)
    xh = np.zeros((n, trips * 256), np.float32)
    xh[:, :K] = split(x)[0].astype(np.float32)
    out = np.empty((n, q.shape[0]))
    for j in range(q.shape[0]):
        qf = np.zeros(trips * 256, np.float32)
        qf[:K] = q[j].astype(np.float32)
        d = (xh - qf).astype(np.float32).reshape(n, trips, 32, 8)
        acc = np.zeros((n, 32), np.float32)
        for t in range(trips):
            a0, a1 = acc.copy(), np.zeros((n, 32), np.float32)
            for c in range(0, 8, 2):
                a0 = (d[:, t, :, c].astype(np.float64) ** 2 + a0).astype(np.float32)         # fmaf
                a1 = (d[:, t, :, c + 1].astype(np.float64) ** 2 + a1).astype(np.float32)
            acc = (a0 + a1).astype(np.float32)
        m = 16
        while m >= 1:
            acc = (acc + acc[:, np.arange(32) ^ m]).astype(np.float32)
            m >>= 1
        out[:, j] = acc[:, 0]
    return out, Kp


def k12_terms(x, q, Kp):
    xh = split(x)[0].astype(np.float64)
    ex = np.sqrt(((x - xh) ** 2).sum(1)).max() * (1 + 1e-12)
    eq = np.sqrt(((q - q.astype(np.float32)) ** 2).sum(1)) * (1 + 1e-9)
    return ex + eq + np.sqrt(x.shape[1]) * 1e-22, (Kp / 32.0 + 12.0) * 2.0 ** -24


# ---- K13: one-byte plane, exact integer keys ------------------------------------------------------------------------
def k13_grid(x):
    lo, hi = float(x.min()), float(x.max())
    step = (hi - lo) / 255.0 if hi > lo else 1.0
    return lo, step


def k13_keys(x, q, lo, step):
    u = np.clip(np.rint((x - lo) / step), 0, 255).astype(np.int64)
    Q = np.clip(np.rint(256.0 * ((q - lo) / step)), 0, 65535).astype(np.int64)
    out = np.empty((x.shape[0], q.shape[0]))
    for j in range(q.shape[0]):
        a, b = Q[j] >> 8, Q[j] & 255
        key_int = 65536 * (u * u).sum(1) - 512 * (256 * (u * a).sum(1) + (u * b).sum(1)) + (Q[j] * Q[j]).sum()
        assert np.all(key_int == ((256 * u - Q[j]) ** 2).sum(1))                              # the kernel's decomposition is exact
        out[:, j] = key_int.astype(np.float64) * (step / 256.0) ** 2
    xhat = lo + step * u
    qhat = lo + step * (Q / 256.0)
    ex = np.sqrt(((x - xhat) ** 2).sum(1)).max() * (1 + 1e-12)
    eq = np.sqrt(((q - qhat) ** 2).sum(1)) * (1 + 1e-9)
    return out, ex + eq, 2.0 ** -50


def check_sqrt_form(key, d, E, gamma):
    lhs = np.abs(np.sqrt(key) - np.sqrt(d))
    rhs = E[None, :] + gamma * (np.sqrt(d) + E[None, :])
    assert np.all(lhs <= rhs * (1 + 1e-9) + 1e-300), float((lhs / rhs).max())


def check_finalize_decisions(key, d, E, gamma, k, K, rng):
    """What tail.cuh does with the bound: (1) the candidates are the entries with key <= window(dk), dk = k-th smallest key
    -- every member of the TRUE top-k must be among them; (2) entries whose key is >= `bound` have d >= lb^2."""
    eps64 = 4.0 * (K + 2) * 2.0 ** -53
    for j in range(key.shape[1]):
        dk = np.sort(key[:, j])[k - 1]
        U = np.sqrt(dk) / (1.0 - gamma) + E[j]
        lim = ((U * (1.0 + eps64) + E[j]) * (1.0 + gamma)) ** 2 * (1.0 + 1e-12)
        cand = key[:, j] <= lim
        true_topk = np.argsort(d[:, j], kind="stable")[:k]
        assert np.all(cand[true_topk]), "a true top-k row fell outside the re-rank window"
        bound = np.quantile(key[:, j], rng.random())                                           # any cut: the proof is per entry
        lb = np.sqrt(bound) / (1.0 + gamma) - E[j]
        if lb > 0:
            assert np.all(d[key[:, j] >= bound, j] >= lb * lb * (1.0 - eps64 - 1e-12))


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("K,seed", [(768, 1), (100, 2), (320, 3)])
def test_k12_hi_plane_key_bound(kind, K, seed):
    rng = np.random.default_rng(seed + 17 * KINDS.index(kind))
    x, q = make(kind, 200, 4, K, rng)
    d = ((x[:, None, :] - q[None, :, :]) ** 2).sum(-1)
    key, Kp = k12_keys(x, q)
    E, gamma = k12_terms(x, q, Kp)
    check_sqrt_form(key, d, E, gamma)
    for k in (1, 5):
        check_finalize_decisions(key, d, E, gamma, k, K, rng)
    if kind == "uniform" and K == 768:      # not vacuous: the window around a distance of ~100 is a fraction of a unit wide
        assert E.max() < 0.03


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("K,seed", [(768, 1), (256, 2), (300, 3)])
def test_k13_byte_plane_key_bound(kind, K, seed):
    rng = np.random.default_rng(seed + 31 * KINDS.index(kind))
    x, q = make(kind, 200, 4, K, rng)
    d = ((x[:, None, :] - q[None, :, :]) ** 2).sum(-1)
    lo, step = k13_grid(x)
    key, E, gamma = k13_keys(x, q, lo, step)
    check_sqrt_form(key, d, E, gamma)
    for k in (1, 5):
        check_finalize_decisions(key, d, E, gamma, k, K, rng)
    if kind == "uniform" and K == 768:
        assert E.max() < 0.04
    # rows appended AFTER the grid was fixed may lie outside it: they are clamped and the measured plane error grows
    x2 = np.vstack([x, x[:5] * 3.0 + 1.0])
    d2 = ((x2[:, None, :] - q[None, :, :]) ** 2).sum(-1)
    key2, E2, _ = k13_keys(x2, q, lo, step)
    check_sqrt_form(key2, d2, E2, gamma)


# ---- two selection shortcuts the kernels take: both only need an UPPER bound of an order statistic --------------------------
@pytest.mark.parametrize("seed", range(6))
def test_kth_smallest_of_thread_minima_bounds_the_kth_smallest(seed):
    """tail.cuh, selection path: every thread keeps the smallest key among the entries it looked at; the k-th smallest of
    those T minima stands in for the k-th smallest key overall.  They are k distinct entries, so it can only be larger --
    which only widens the re-rank window -- and for k = 1 it is the minimum itself."""
    rng = np.random.default_rng(seed)
    T = 128
    for total, k in ((2664, 1), (2664, 4), (4440, 10), (4440, 24), (300, 24), (40, 24)):
        keys = rng.random(total) if seed % 2 else np.round(rng.random(total), 2)          # with and without ties
        valid = rng.random(total) < (0.9 if total > 100 else 0.5)
        mins = np.full(T, np.inf)
        for t in range(T):
            mine = keys[t::T][valid[t::T]]
            if mine.size:
                mins[t] = mine.min()
        kth_of_minima = np.sort(mins)[k - 1]
        v = np.sort(keys[valid])
        if v.size >= k and np.isfinite(kth_of_minima):
            assert kth_of_minima >= v[k - 1]
            if k == 1:
                assert kth_of_minima == v[0]


@pytest.mark.parametrize("seed", range(4))
def test_cap_th_smallest_of_group_minima_bounds_the_groups_cap_th_smallest(seed):
    """umma_filter.cu (K10): every CTA of a query group publishes the smallest key it has kept; CTAs own disjoint rows, so the
    cap-th smallest of the published minima is the key of a row with at least cap - 1 rows below it: an upper bound of the
    group's cap-th smallest key (what the filter may use as a threshold), and a far tighter one than the smallest of the CTAs'
    own cap-th smallest keys, which is what the thresholds came from before."""
    rng = np.random.default_rng(seed)
    streams, cap = 37, 24
    for rows_per_cta in (128, 1280, 20000):
        keys = rng.standard_normal((streams, rows_per_cta)) ** 2
        truth = np.sort(keys.ravel())[cap - 1]
        from_minima = np.sort(keys.min(axis=1))[cap - 1]
        own_cap_th = np.sort(keys, axis=1)[:, cap - 1].min()
        assert from_minima >= truth
        assert own_cap_th >= truth
        if rows_per_cta >= 1280:
            # keys that pass the threshold (= survivors to append, sort, prune): several times fewer
            assert (keys < from_minima).sum() * 3 < (keys < own_cap_th).sum()


def test_k13_pair_shares_the_squares():
    """scan_plane8_kernel<NQ = 2>: sum u^2 is formed once per row and used for both queries of the pair."""
    rng = np.random.default_rng(3)
    x = rng.random((50, 300))
    q = rng.random((2, 300))
    lo, step = k13_grid(x)
    both, _, _ = k13_keys(x, q, lo, step)
    for j in range(2):
        one, _, _ = k13_keys(x, q[j:j + 1], lo, step)
        assert np.array_equal(both[:, j], one[:, 0])


def test_device_quantisers_equal_the_numpy_model(tmp_path):
    """The two expressions that put rows and queries on K13's grids (csrc/common.cuh: p8_quant_x, p8_quant_q) are plain C:
    compiled here for the host from the very source text the kernels include and compared with the numpy model the bound
    tests above are written in -- rounding mode, clamps, NaN and infinities included."""
    import ctypes
    import os
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "simple-vector-db_b200", "csrc", "common.cuh")).read()
    fns = re.findall(r"__device__ __forceinline__ (unsigned p8_quant_[xq]\(double v, double lo, double step\) \{.*?\n\})", src, re.S)
    assert len(fns) == 2
    c = "#include <math.h>\n" + "\n".join("static " + f for f in fns) + """
void quant(const double *v, int n, double lo, double step, unsigned *ux, unsigned *uq) {
    for (int i = 0; i < n; i++) { ux[i] = p8_quant_x(v[i], lo, step); uq[i] = p8_quant_q(v[i], lo, step); }
}
"""
    (tmp_path / "q.c").write_text(c)
    so = str(tmp_path / "q.so")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", str(tmp_path / "q.c"), "-o", so, "-lm"], check=True)
    lib = ctypes.CDLL(so)
    rng = np.random.default_rng(5)
    for lo, step in ((0.0, 1.0 / 255.0), (-3.7, 0.031), (1000.0, 1e-3), (0.25, 7.0)):
        v = np.concatenate([lo + step * 255.0 * rng.random(4000), lo + step * (rng.integers(0, 512, 500) / 2.0),      # incl. exact halves
                            lo + step * 300.0 * rng.standard_normal(500), [np.nan, np.inf, -np.inf, lo, lo - step, lo + 255 * step, 1e300, -1e300]])
        ux = np.zeros(len(v), np.uint32)
        uq = np.zeros(len(v), np.uint32)
        lib.quant(v.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), len(v), ctypes.c_double(lo), ctypes.c_double(step),
                  ux.ctypes.data_as(ctypes.POINTER(ctypes.c_uint)), uq.ctypes.data_as(ctypes.POINTER(ctypes.c_uint)))
        with np.errstate(invalid="ignore"):
            tx = np.rint((v - lo) / step)
            tq = np.rint(256.0 * ((v - lo) / step))
        wx = np.where(np.isnan(tx), 0, np.clip(tx, 0, 255)).astype(np.uint32)
        wq = np.where(np.isnan(tq), 0, np.clip(tq, 0, 65535)).astype(np.uint32)
        assert np.array_equal(ux, wx) and np.array_equal(uq, wq)


def test_bound_coefficients_in_the_sources_are_the_ones_the_tests_use(tmp_path):
    """plane_gamma (K12), umma_eabs_coef (K10) and shadow_eps (K11) are one-line host functions next to their kernels; the
    numpy bound tests restate them.  Compiled from the source text and compared, so that the two cannot drift apart."""
    import ctypes
    import os
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(root, "simple-vector-db_b200", "csrc")
    bodies = []
    for fname, sig in (("plane_scan.cu", r"double plane_gamma\(int Kp\) \{[^\n]*\}"), ("umma_filter.cu", r"double umma_eabs_coef\(int K\) \{[^\n]*\}"),
                       ("scan_kernels.cu", r"double shadow_eps\(int K\) \{[^\n]*\}")):
        m = re.search(sig, open(os.path.join(csrc, fname)).read())
        assert m, sig
        bodies.append(m.group(0))
    (tmp_path / "c.c").write_text("#include <math.h>\n" + "\n".join(bodies) + "\n")
    so = str(tmp_path / "c.so")
    subprocess.run(["gcc", "-O2", "-shared", "-fPIC", str(tmp_path / "c.c"), "-o", so, "-lm"], check=True)
    lib = ctypes.CDLL(so)
    for f in (lib.plane_gamma, lib.umma_eabs_coef, lib.shadow_eps):
        f.restype = ctypes.c_double
        f.argtypes = [ctypes.c_int]
    for K in (64, 100, 128, 320, 768, 1024):
        Kp = -(-K // 64) * 64
        assert lib.plane_gamma(Kp) == (Kp / 32.0 + 12.0) * 2.0 ** -24 == k12_terms(np.zeros((1, K)), np.zeros((1, K)), Kp)[1]
        assert lib.umma_eabs_coef(K) == 3.2 * 2.0 ** -16 + (3.0 * K / 16.0) * 2.0 ** -21 + 8.0 * 2.0 ** -20
        assert lib.shadow_eps(K) == 2.0 ** -13 + (K / 32.0 + 12.0) * 2.0 ** -24
