"""K10's selection SOURCE (csrc/umma_select.cuh: pruning a candidate buffer to its cap smallest keys by bitwise
selection, the conservative threshold, the final bitonic sort) executed on the CPU under the emulation of the CUDA
execution model in tests/cusim/ and checked against std::sort.  The tensor-core half of K10 only runs on a GPU
(tests/test_gpu_umma.py)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "simple-vector-db_b200", "csrc")
SIM = os.path.join(ROOT, "tests", "cusim")


@pytest.fixture(scope="module")
def sim_binary(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("umma_sim") / "umma_select_sim")
    subprocess.run(["g++", "-std=c++20", "-O1", "-pthread", "-I", SIM, "-I", CSRC,
                    os.path.join(SIM, "umma_select_sim_main.cpp"), "-o", exe], check=True)
    return exe


@pytest.mark.parametrize("seed,nbuf,cap,mode", [
    (1, 12, 24, 0),      # the bench shape: cap = k + 14 with k = 10
    (2, 12, 32, 0),      # the largest cap
    (3, 12, 15, 1),      # five distinct key values: the cap-th smallest sits inside a long run of ties
    (4, 12, 24, 2),      # negative keys and the -FLT_MAX stand-in of non-finite keys
    (5, 10, 1, 1),       # cap = 1
])
def test_selection_source_under_cpu_emulation(sim_binary, seed, nbuf, cap, mode):
    r = subprocess.run([sim_binary, str(seed), str(nbuf), str(cap), str(mode)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr
