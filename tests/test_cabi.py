"""CPU-side checks of the C boundary: the library builds, loads without a GPU, exports every
symbol the headers declare, and refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, PKG

from svdb import binding as B


@pytest.fixture(scope="module")
def lib():
    import importlib.util
    spec = importlib.util.spec_from_file_location("svdb_build", os.path.join(PKG, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    return B.lib()


def _declared(header: str):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b((?:svdb|kdtree|vector_db)_[a-z0-9_]+|cosine_similarity|euclidean_distance|dot_product)\s*\(", text))
    return names


def test_exports_every_declared_symbol(lib):
    declared = _declared("svdb_b200.h") | _declared("svdb_dropin.h")
    assert declared >= set(B.NATIVE_SYMBOLS) | set(B.DROPIN_SYMBOLS)
    out = subprocess.run(["nm", "-D", "--defined-only", B.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = sorted(declared - exported)
    assert not missing, f"declared in include/ but not exported: {missing}"


def test_no_dependency_on_driver_or_oracle(lib):
    out = subprocess.run(["ldd", B.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "libcuda.so" not in out and "oracle" not in out and "svdb_ref" not in out


def test_struct_layouts_match_reference_prefix():
    """Vector is 56 bytes (uuid[37] + pad, size_t, double*); the db/tree prefixes line up."""
    from oracle.binding import Vector, VectorDatabaseS, KDTreeS
    assert C.sizeof(Vector) == 56 and Vector.dimension.offset == 40 and Vector.data.offset == 48
    assert VectorDatabaseS.size.offset == 8 and VectorDatabaseS.kdtree.offset == 24
    assert KDTreeS.dimension.offset == 8
    assert B.candidate_dtype.itemsize == 32


def test_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(B.SvdbError) as ei:
        B.Engine(16, 4)
    assert "no CUDA device" in str(ei.value) or "CPU path" in str(ei.value)
    with pytest.raises(B.SvdbError):
        B.compare_vectors(B.DOT, np.ones(4), np.ones(4))


def test_argument_checks_precede_device_use(lib):
    cfg = B.Config(4, 8, 0, 0, 0, 0)   # kd_dim > dimension
    h = C.c_void_p()
    assert lib.svdb_engine_create(C.byref(cfg), C.byref(h)) == -1
    assert b"kd_dim" in lib.svdb_last_error()
    assert lib.svdb_engine_create(None, C.byref(h)) == -1


def test_dropin_without_gpu_keeps_reference_sentinels(lib, capfd):
    """kdtree_nearest on an empty tree is (size_t)-1 (kdtree.c:172-177) and needs no device."""
    from oracle.binding import RefApi, NONE
    api = RefApi(B.LIB_PATH)
    t = api.lib.kdtree_create(3)
    q = np.zeros(3)
    assert api.lib.kdtree_nearest(t, q.ctypes.data_as(C.POINTER(C.c_double))) == NONE
    api.lib.kdtree_free(t)
    assert api.metric(0, np.ones(3), np.ones(4)) == np.float32(-1.0)   # vector_database.c:302-305
    db = api.lib.vector_db_init(0, 3)
    assert db.contents.size == 0 and db.contents.capacity == 10 and bool(db.contents.kdtree)
    assert not api.lib.vector_db_read(db, 0)
    api.lib.vector_db_update(db, 5, api.make_vector(np.ones(3), own=False))   # out of range: silent no-op
    api.lib.vector_db_delete(db, 5)
    api.lib.vector_db_free(db)


def test_every_engine_option_is_documented():
    """Every name svdb_set_option accepts appears in the public header, INTEGRATION.md or DESIGN.md."""
    import re
    src = open(os.path.join(ROOT, "simple-vector-db_b200", "csrc", "engine.cu")).read()
    names = sorted(set(re.findall(r'n == "([a-z0-9_.]+)"', src)))
    assert len(names) > 30
    docs = "".join(open(os.path.join(ROOT, f)).read() for f in ("include/svdb_b200.h", "INTEGRATION.md", "DESIGN.md"))
    missing = [n for n in names if n not in docs]
    assert not missing, missing
