"""What does the tcgen05 fp32 accumulator (TMEM) lose?  K10 (csrc/umma_filter.cu) budgets an ABSOLUTE key error
E = umma_eabs_coef(K) * (max|x|^2 + |q|^2); round 1 took the accumulation part of it from a "deliberately loose model" of
the tensor core's adder.  This probe MEASURES it: adversarial rows and queries whose coordinates are all +-powers of two
(exact in bf16: the lo planes are zero, no product term is dropped, every product is exact), so that the only error sources
are the fp32 accumulation inside and between the K/16 chained `tcgen05.mma kind::f16` instructions of a key and the three
fp32 roundings outside it (norms, their sum, the final fma: <= 4 * 2^-24 of the scale).  Keys are dumped by the kernel
itself (option umma.debug_keys: the first 128 log rows x the first query group) and compared with exact arithmetic.

Patterns (x = row, q = query; products p_i = x_i q_i):
  big_first(t) / big_last(t) / big_mid(t): one product of magnitude 1, K-1 same-sign products of 2^t, t = -30 .. -16
      -- what a truncating aligned adder without guard bits loses entirely (IEEE fp32 summed sequentially loses them too
      while t <= -25 + ...; a wide internal adder keeps them)
  big_first_neg(t): the small ones with the opposite sign
  block_heads(t): a big product at the head of every 16-coordinate MMA step, 15 small ones behind it
  alternating / block_alternating: +1, -1, ... (cancellation inside an instruction / between instructions)
  ramps: 2^-(i mod m), all positive
  random: random signs, exponents -12 .. 0
against queries q = s * (+1 ...) and s * (+1, -1, ...) for eight scales s.

    python scripts/umma_accumulator_probe.py            # one JSON line per K (needs a GPU)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "simple-vector-db_b200")]
from svdb import binding as B  # noqa: E402

TS = list(range(-30, -15, 2))          # 8 values x 5 families = 40 patterns


def patterns(K: int, rng):
    rows, names = [], []

    def add(name, v):
        rows.append(np.asarray(v, dtype=np.float64))
        names.append(name)

    for t in TS:
        s = 2.0 ** t
        v = np.full(K, s); v[0] = 1.0; add(f"big_first({t})", v)
        v = np.full(K, s); v[K - 1] = 1.0; add(f"big_last({t})", v)
        v = np.full(K, s); v[K // 2] = 1.0; add(f"big_mid({t})", v)
        v = np.full(K, -s); v[0] = 1.0; add(f"big_first_neg({t})", v)
        v = np.full(K, s); v[::16] = 1.0; add(f"block_heads({t})", v)
    add("alternating", np.where(np.arange(K) % 2 == 0, 1.0, -1.0))
    add("block_alternating", np.where((np.arange(K) // 16) % 2 == 0, 1.0, -1.0))
    for m in (8, 16, 24):
        add(f"ramp({m})", 2.0 ** -(np.arange(K) % m))
    add("ramp_down", 2.0 ** -np.floor(np.arange(K) * 24.0 / K))
    add("ramp_up", 2.0 ** -np.floor((K - 1 - np.arange(K)) * 24.0 / K))
    while len(rows) < 128:
        add("random", rng.choice([-1.0, 1.0], K) * 2.0 ** rng.integers(-12, 1, K))
    return np.stack(rows[:128]), names[:128]


def queries(K: int):
    qs = []
    for s in (2.0 ** e for e in (-6, -3, -1, 0, 1, 2, 3, 5)):
        for shape in range(8):
            if shape == 0:
                v = np.ones(K)
            elif shape == 1:
                v = np.where(np.arange(K) % 2 == 0, 1.0, -1.0)
            elif shape == 2:
                v = np.where((np.arange(K) // 16) % 2 == 0, 1.0, -1.0)
            elif shape == 3:
                v = -np.ones(K)
            elif shape == 4:
                v = 2.0 ** -(np.arange(K) % 8)
            elif shape == 5:
                v = np.where(np.arange(K) % 3 == 0, 1.0, 0.5)
            elif shape == 6:
                v = np.where(np.arange(K) < K // 2, 1.0, -1.0)
            else:
                v = 2.0 ** -(np.arange(K) % 2)
            qs.append(s * v)
    return np.stack(qs)            # 64 queries


def probe(K: int, seed: int = 1):
    rng = np.random.Generator(np.random.PCG64(seed))
    X, names = patterns(K, rng)
    Q = queries(K)
    with B.Engine(K, K) as e:
        e.insert(X)
        e.set_option("nearest.umma_min_queries", 1)
        e.set_option("nearest.umma_min_kd_dim", 1)
        e.set_option("umma.debug_keys", 1)
        e.nearest(Q, 1)
        keys = e.debug_filter_keys(64).astype(np.float64)           # [128 rows][64 queries]
        coef = None
    xn, qn = (X ** 2).sum(1), (Q ** 2).sum(1)
    dot = X @ Q.T                                                   # exact: dyadic terms spanning < 53 bits
    exact = xn[:, None] + qn[None, :] - 2.0 * dot
    err = np.abs(keys - exact)
    scale = xn.max() + qn[None, :]                                  # what finalize multiplies the coefficient with
    sabs = np.abs(X) @ np.abs(Q).T                                  # sum |x_i q_i|
    outside = 4.0 * 2.0 ** -24 * (xn[:, None] + qn[None, :])        # roundings outside the accumulator
    acc_err = np.maximum(err - outside, 0.0) / 2.0                  # the key carries -2 * acc
    per_mma = acc_err / sabs / (K / 16.0)
    r, c = np.unravel_index(np.argmax(err / scale), err.shape)
    r2, c2 = np.unravel_index(np.argmax(per_mma), per_mma.shape)
    by_family = {}
    for i, nm in enumerate(names):
        fam = nm.split("(")[0]
        by_family[fam] = max(by_family.get(fam, 0.0), float((acc_err[i] / sabs[i]).max()))
    return {"K": K, "chained_mma_steps": K // 16, "rows": 128, "queries": 64,
            "key_err_over_scale_max": float((err / scale).max()), "worst_pattern": names[r], "worst_query": int(c),
            "acc_err_over_sum_abs_products_max": float((acc_err / sabs).max()),
            "acc_err_over_sum_abs_products_per_mma_max": float(per_mma.max()), "worst_per_mma_pattern": names[r2],
            "in_units_of_2^-24_per_mma": float(per_mma.max() * 2.0 ** 24),
            "acc_err_over_sum_abs_products_by_family": by_family}


def main():
    for K in (64, 256, 768, 1024):
        print(json.dumps(probe(K)), flush=True)


if __name__ == "__main__":
    main()
