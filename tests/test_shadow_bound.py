"""The key-error bound K11 hands to finalize (csrc/scan_kernels.cu: shadow_eps, shadow_eabs_coef),
checked in numpy: keys formed the way scan_shadow_kernel forms them -- x^ = hi + lo in fp32, diff = x^ - fl32(q),
fp32 multiply-accumulate, one accumulator per lane of 8-coordinate slices, butterfly over 32 lanes -- stay within
eps * d + eabs * (max|x|^2 + |q|^2) of the exact squared distance, on benign and on adversarial inputs.  A bound that
were too small would mean silently wrong answers, so this is pinned on the CPU as well as on the device."""
import numpy as np
import pytest

from test_umma_bound import split


def shadow_keys(x: np.ndarray, q: np.ndarray) -> np.ndarray:
    """fp32 emulation of scan_shadow_kernel's key for every (row, query)."""
    n, K = x.shape
    Kp = -(-K // 256) * 256
    hi, lo = split(x)
    xh = np.zeros((n, Kp), np.float32)
    xh[:, :K] = (hi.astype(np.float32) + lo.astype(np.float32))           # exact in fp32
    qf = np.zeros((q.shape[0], Kp), np.float32)
    qf[:, :K] = q.astype(np.float32)
    out = np.empty((n, q.shape[0]), np.float64)
    for j in range(q.shape[0]):
        d = xh - qf[j]                                                     # fp32 subtract
        sq = (d * d).astype(np.float32)                                    # the kernel fuses this rounding away (FFMA): fewer roundings there
        lanes = sq.reshape(n, Kp // 256, 32, 8)                            # trip, lane, coordinate within the lane's slice
        acc = np.zeros((n, 32), np.float32)
        for t in range(Kp // 256):
            for c in range(8):
                acc = (acc + lanes[:, t, :, c]).astype(np.float32)         # sequential per lane
        m = 16
        while m >= 1:                                                      # butterfly
            acc = (acc + acc[:, np.arange(32) ^ m]).astype(np.float32)
            m >>= 1
        out[:, j] = acc[:, 0]
    return out


def bound(K, x, q, d):
    eps = 2.0 ** -13 + (K / 32.0 + 12.0) * 2.0 ** -24
    delta = 2.0 ** -16 + 2.0 ** -24
    eabs_coef = (1.0 + 8192.0) * 2.0 * delta * delta * 1.01
    scale = (x ** 2).sum(1).max() + (q ** 2).sum(1)[None, :]
    return eps * d + eabs_coef * scale


@pytest.mark.parametrize("kind,K,seed", [
    ("uniform", 768, 1), ("uniform", 100, 2), ("normal", 320, 3),
    ("near_query", 256, 4),        # rows within 1e-3 of the query: d tiny against the norms (the absolute term must carry it)
    ("offset", 128, 5),            # 1000 + U[0,1): cancellation in the differences
    ("mixed_magnitudes", 512, 6),  # coordinates spread over six decades
    ("cauchy", 200, 7),            # heavy tails: a few coordinates carry the norms
    ("sparse", 1000, 8),           # 5 % non-zeros
    ("small_integers", 50, 9),     # exact ties galore
])
def test_shadow_key_error_within_bound(kind, K, seed):
    rng = np.random.default_rng(seed)
    n, nq = 200, 4
    if kind == "uniform":
        x, q = rng.random((n, K)), rng.random((nq, K))
    elif kind == "normal":
        x, q = rng.standard_normal((n, K)), rng.standard_normal((nq, K))
    elif kind == "near_query":
        q = rng.random((nq, K))
        x = np.repeat(q, n // nq, axis=0) + 1e-3 * rng.standard_normal((n, K))
    elif kind == "offset":
        x, q = 1000.0 + rng.random((n, K)), 1000.0 + rng.random((nq, K))
    elif kind == "cauchy":
        x, q = rng.standard_cauchy((n, K)).clip(-1e3, 1e3), rng.standard_cauchy((nq, K)).clip(-1e3, 1e3)
    elif kind == "sparse":
        x = np.where(rng.random((n, K)) < 0.05, rng.standard_normal((n, K)) * 100, 0.0)
        q = np.where(rng.random((nq, K)) < 0.05, rng.standard_normal((nq, K)) * 100, 0.0)
    elif kind == "small_integers":
        x, q = rng.integers(-3, 4, (n, K)).astype(float), rng.integers(-3, 4, (nq, K)).astype(float)
    else:
        mag = 10.0 ** rng.integers(-3, 3, size=K).astype(float)
        x, q = rng.standard_normal((n, K)) * mag, rng.standard_normal((nq, K)) * mag
    d = ((x[:, None, :] - q[None, :, :]) ** 2).sum(-1)
    err = np.abs(shadow_keys(x, q) - d)
    b = bound(K, x, q, d)
    assert np.all(err <= b), float((err / b).max())
    # and the bound is not vacuous on benign data: a few percent of the gaps between neighbours at most
    if kind == "uniform" and K == 768:
        assert b.max() < 0.03 and err.max() < b.max() / 5
