"""Synthetic inputs for the workloads BASELINE.json names.

Pure numpy; no reference code is needed to regenerate them.
"""
from __future__ import annotations

import numpy as np


def script_values(seed: int, shape) -> np.ndarray:
    """Values distributed like the reference's load generator writes them.

    test/add_vectors.sh:15-27: integer part uniform in 1..9, then 0..6 decimal
    digits, each uniform in 0..9, printed as text and parsed by the server's JSON
    reader.  Few distinct short decimals => exact duplicate kd-points and exact
    distance ties do occur, which is why the tie rule is tested on this data.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    n = int(np.prod(shape))
    ip = rng.integers(1, 10, size=n)
    places = rng.integers(0, 7, size=n)
    digits = rng.integers(0, 10, size=(n, 6))
    # "i.d1..dp" parsed by strtod == correctly rounded (i*10^p + d1..dp) / 10^p: both operands are
    # exact doubles and IEEE division rounds correctly, so this equals float(text) bit for bit
    frac = np.zeros(n, dtype=np.int64)
    for j in range(6):
        frac = np.where(j < places, frac * 10 + digits[:, j], frac)
    scale = 10.0 ** places
    out = (ip * scale + frac) / scale
    return out.reshape(shape)


def uniform_rows(seed: int, n: int, D: int) -> np.ndarray:
    """U[0,1) fp64 rows (configs 2, 3, 5)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.random((n, D), dtype=np.float64)


def normal_rows(seed: int, n: int, D: int) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.standard_normal((n, D))


def index_pairs(seed: int, n_pairs: int, n_rows: int):
    rng = np.random.Generator(np.random.PCG64(seed))
    return (rng.integers(0, n_rows, size=n_pairs, dtype=np.uint64),
            rng.integers(0, n_rows, size=n_pairs, dtype=np.uint64))
