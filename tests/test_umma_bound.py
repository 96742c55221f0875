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
