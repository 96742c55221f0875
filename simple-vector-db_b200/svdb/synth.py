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
    out = np.empty(n, dtype=np.float64)
    for i in range(n):
        p = int(places[i])
        if p:
            out[i] = float(f"{ip[i]}." + "".join(str(int(d)) for d in digits[i, :p]))
        else:
            out[i] = float(ip[i])
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
