"""The K8/K9 kernel SOURCE (csrc/median_tree.cu: radix-select median build + balanced-tree traversal) executed on
the CPU under a small emulation of the CUDA execution model (tests/cusim/: CTAs as std::threads, __syncthreads,
warp collectives with lane masks).  Checks the build invariants (permutation, valid median split at every node,
leaves of <= 32) and the traversal's answers and tie flags against brute force in the reference's arithmetic
(kdtree.c:134-137), for every lanes-per-query variant.  The GPU parity tests run the same code on the device."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "simple-vector-db_b200", "csrc")
SIM = os.path.join(ROOT, "tests", "cusim")


@pytest.fixture(scope="module")
def sim_binary(tmp_path_factory):
    tmp = tmp_path_factory.mktemp("mtree_sim")
    src = open(os.path.join(CSRC, "median_tree.cu")).read()
    src = src.replace('#include "common.cuh"', '#include "cusim_common.h"')
    # kernel<<<grid, block, smem, stream>>>(args)  ->  cusim::launch(grid, block, kernel, args)
    src, n = re.subn(r"(\w+(?:<[\w, ]+>)?)<<<(.*?), (\w+), 0, st>>>\(", r"cusim::launch(\2, \3, \1, ", src)
    assert n == 13, n
    assert "<<<" not in src
    (tmp / "median_tree_sim.inc").write_text(src)
    exe = str(tmp / "mtree_sim")
    subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-ffp-contract=off", "-I", str(tmp), "-I", SIM, "-I", CSRC,
                    os.path.join(SIM, "mtree_sim_main.cpp"), "-o", exe], check=True)
    return exe


# n, K, tail, nq, distribution (0 uniform, 1 coarse grid, 2 sorted, 3 non-finite rows), seed
CASES = [
    (0, 3, 0, 4, 0, 1),            # empty
    (0, 3, 40, 6, 1, 2),           # tail only
    (1, 1, 0, 4, 0, 3),
    (33, 2, 5, 12, 1, 4),          # one split, ties
    (700, 3, 70, 24, 0, 5),        # one CTA, several levels
    (2048, 3, 0, 16, 1, 6),        # largest single-CTA build, coarse grid: duplicates and distinct ties
    (2049, 2, 33, 16, 0, 7),       # one radix-select level, uneven halves
    (9001, 3, 100, 24, 0, 8),      # three radix-select levels
    (4600, 3, 0, 16, 1, 9),        # two radix-select levels on a grid: medians inside long runs of equal keys
    (4200, 8, 10, 8, 0, 10),       # K = 8
    (4200, 2, 0, 12, 2, 11),       # sorted insertion order
    (4500, 3, 64, 16, 3, 12),      # NaN / +-inf coordinates
]


@pytest.mark.parametrize("n,K,tail,nq,dist,seed", CASES)
def test_kernel_source_under_cpu_emulation(sim_binary, n, K, tail, nq, dist, seed):
    blk = "1" if seed % 4 == 0 else "3"          # split values in heap order / in 64-byte blocks of three levels
    r = subprocess.run([sim_binary, str(n), str(K), str(tail), str(nq), str(dist), str(seed), blk], capture_output=True,
                       text=True, timeout=180)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr
