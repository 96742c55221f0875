"""The arithmetic K10's error bound rests on (csrc/umma_filter.cu header), checked in numpy with exact
(float64) evaluation of the bf16 products: x = xh + xl + r with |r| <= (2^-16 + 2^-24)|x|, and the three products the
tensor cores form differ from <x, q> by at most 3.1 * 2^-16 * sum|x_i q_i|.  The accumulation term of the bound is
measured on the device (tests/test_gpu_umma.py::test_umma_key_error_is_inside_the_bound)."""
import numpy as np
import pytest


def bf16_rne(f32: np.ndarray) -> np.ndarray:
    """float32 -> nearest bfloat16 (ties to even), returned as float32."""
    u = f32.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def split(x: np.ndarray):
    f = x.astype(np.float32)
    hi = bf16_rne(f)
    lo = bf16_rne(f - hi)
    return hi.astype(np.float64), lo.astype(np.float64)


@pytest.mark.parametrize("scale,seed", [(1.0, 1), (1e-6, 2), (1e6, 3), (1e12, 4)])
def test_split_residual_and_product_bound(scale, seed):
    rng = np.random.default_rng(seed)
    K = 768
    x = (rng.random((64, K)) - 0.3) * scale
    q = rng.standard_normal((16, K)) * scale
    xh, xl = split(x)
    qh, ql = split(q)
    assert np.all(np.abs(x - xh - xl) <= (2.0 ** -16 + 2.0 ** -24) * np.abs(x))
    assert np.all(np.abs(xl) <= 2.0 ** -8 * np.abs(x) * (1 + 2.0 ** -8))
    approx = xh @ qh.T + xh @ ql.T + xl @ qh.T
    exact = x @ q.T
    bound = 3.1 * 2.0 ** -16 * (np.abs(x) @ np.abs(q).T)
    assert np.all(np.abs(approx - exact) <= bound)
    # and with sum|x_i q_i| <= (|x|^2 + |q|^2) / 2 the key error 2 * |approx - exact| stays below the first term of umma_eabs_coef
    coef_repr = 3.2 * 2.0 ** -16
    scale2 = (x ** 2).sum(1)[:, None] + (q ** 2).sum(1)[None, :]
    assert np.all(2 * np.abs(approx - exact) <= coef_repr * scale2)


@pytest.mark.parametrize("kind,seed", [("uniform", 1), ("cauchy", 2), ("near_query", 3), ("scaled", 4)])
def test_key_error_with_sequential_fp32_accumulation_within_umma_eabs_coef(kind, seed):
    """The whole K10 key, emulated with one running fp32 sum over all 3K bf16 products (3K roundings; the tensor core
    rounds once per instruction of 16 products -- measured, scripts/umma_accumulator_probe.py): a sanity check on realistic
    data that stays inside coef * (max|x|^2 + |q|^2), coef = umma_eabs_coef(K).  The bound itself rests on the GPU
    measurement pinned by tests/test_gpu_umma.py::test_tcgen05_accumulator_loss_is_inside_the_budget."""
    rng = np.random.default_rng(seed)
    n, nq, K = 48, 3, 256
    if kind == "uniform":
        x, q = rng.random((n, K)), rng.random((nq, K))
    elif kind == "cauchy":
        x, q = rng.standard_cauchy((n, K)).clip(-1e3, 1e3), rng.standard_cauchy((nq, K)).clip(-1e3, 1e3)
    elif kind == "near_query":
        q = rng.standard_normal((nq, K))
        x = np.repeat(q, n // nq, axis=0) * (1 + 1e-5 * rng.standard_normal((n, K)))
    else:
        x, q = rng.standard_normal((n, K)) * 1e7, rng.standard_normal((nq, K)) * 1e7
    xh, xl = split(x)
    qh, ql = split(q)
    prod = np.zeros((n, nq), np.float32)
    for i in range(K):
        for a, b in ((xh, ql), (xl, qh), (xh, qh)):
            prod = (prod + a[:, i:i + 1].astype(np.float32) * b[:, i].astype(np.float32)[None, :]).astype(np.float32)
    xn, qn = (x ** 2).sum(1).astype(np.float32), (q ** 2).sum(1).astype(np.float32)
    key = (xn[:, None] + qn[None, :]).astype(np.float32) + np.float32(-2) * prod
    d = ((x[:, None, :] - q[None, :, :]) ** 2).sum(-1)
    coef = 3.2 * 2.0 ** -16 + (3.0 * K / 16.0) * 2.0 ** -21 + 8.0 * 2.0 ** -20
    scale = (x ** 2).sum(1).max() + (q ** 2).sum(1)[None, :]
    assert np.all(np.abs(key.astype(np.float64) - d) <= coef * scale)
