"""The reference handlers' call pattern, in C (tests/c/callsite_harness.c): ONE object file,
compiled against compat/vector_database.h, linked against the reference's own library and
against the CUDA drop-in.  CPU: it compiles and links against both.  GPU: both runs print the
same bytes."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, PKG

from svdb import synth

HARNESS = os.path.join(ROOT, "tests", "c", "callsite_harness.c")
LIBDIR = os.path.join(PKG, "lib")
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def _build(tmp_path, libdir, libname, tag):
    obj = tmp_path / "harness.o"
    if not obj.exists():
        subprocess.run(["gcc", "-O1", "-Wall", "-I", os.path.join(ROOT, "compat"), "-c", HARNESS, "-o", str(obj)], check=True)
    exe = tmp_path / f"harness_{tag}"
    subprocess.run(["gcc", str(obj), "-o", str(exe), f"-L{libdir}", f"-l{libname}", f"-Wl,-rpath,{libdir}", "-lm", "-pthread"],
                   check=True)
    return str(exe)


def _write_input(path, n, D, K, nq, npairs, nops, seed, coarse):
    rng = np.random.Generator(np.random.PCG64(seed))
    gen = (lambda shape: synth.script_values(int(rng.integers(1, 1 << 30)), shape)) if coarse else (lambda shape: rng.random(shape))
    rows, Q = gen((n, D)), gen((nq, D))
    pairs = rng.integers(0, n + 2, size=(npairs, 2), dtype=np.uint64)   # a few out of bounds
    with open(path, "wb") as f:
        np.array([n, D, K, nq, npairs, nops], dtype=np.uint64).tofile(f)
        rows.tofile(f)
        Q.tofile(f)
        pairs.tofile(f)
        for i in range(nops):
            code = 1 if rng.random() < 0.6 else 2
            np.array([code, rng.integers(0, n)], dtype=np.uint64).tofile(f)
            gen((D,)).tofile(f)


@pytest.fixture(scope="module")
def built_lib():
    import importlib.util
    spec = importlib.util.spec_from_file_location("svdb_build", os.path.join(PKG, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()


def test_harness_compiles_and_links_against_both(tmp_path, built_lib, ref):
    _build(tmp_path, LIBDIR, "svdb_b200", "ours")
    _build(tmp_path, REFDIR, "svdb_ref", "ref")


@pytest.mark.gpu
@pytest.mark.parametrize("n,D,K,coarse", [(600, 24, 3, False), (400, 40, 40, False), (500, 6, 3, True)])
def test_same_output_as_reference(tmp_path, built_lib, n, D, K, coarse):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if not os.path.exists(os.path.join(REFDIR, "libsvdb_ref.so")):
        pytest.skip("oracle/_ref was not shipped")
    ours = _build(tmp_path, LIBDIR, "svdb_b200", "ours")
    refx = _build(tmp_path, REFDIR, "svdb_ref", "ref")
    inp = str(tmp_path / "in.bin")
    _write_input(inp, n, D, K, nq=40, npairs=30, nops=60, seed=n + D, coarse=coarse)
    a = subprocess.run([refx, inp, str(tmp_path / "ref.db")], capture_output=True, text=True, check=True).stdout
    b = subprocess.run([ours, inp, str(tmp_path / "ours.db")], capture_output=True, text=True, check=True).stdout
    la, lb = a.splitlines(), b.splitlines()
    assert len(la) == len(lb) and len(la) > 100
    diff = [(x, y) for x, y in zip(la, lb) if x != y]
    assert not diff, diff[:5]      # ids (ties included), uuids, metric bits, sizes: all identical
    assert open(tmp_path / "ref.db", "rb").read() == open(tmp_path / "ours.db", "rb").read()
