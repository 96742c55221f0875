"""The drop-in claim, checked against the reference's OWN callers: its unmodified
src/*_handler.c and src/main.c are compiled (with their own include/*.h) and linked against
libsvdb_b200.so instead of src/vector_database.c + src/kdtree.c.  libmicrohttpd and cJSON are
not installed here, so tests/c/stubs/ provides header + link stand-ins for them (never run).
Runs only where the reference tree exists (the build container)."""
import os
import subprocess

import pytest

from conftest import ROOT, PKG

REF = "/root/reference"
STUBS = os.path.join(ROOT, "tests", "c", "stubs")
LIBDIR = os.path.join(PKG, "lib")
CALLERS = ["compare_handler", "get_handler", "post_handler", "put_handler", "delete_handler", "main"]
L1_API = {"vector_db_init", "vector_db_free", "vector_db_insert", "vector_db_read", "vector_db_read_by_uuid",
          "vector_db_update", "vector_db_delete", "vector_db_save", "vector_db_load", "cosine_similarity",
          "euclidean_distance", "dot_product", "kdtree_create", "kdtree_nearest"}


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference tree not present")
def test_reference_callers_compile_and_link_against_the_dropin(tmp_path):
    import importlib.util
    spec = importlib.util.spec_from_file_location("svdb_build", os.path.join(PKG, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    objs = []
    for name in CALLERS:
        obj = str(tmp_path / f"{name}.o")
        # -include stdint.h: the reference relies on a transitive include that glibc does not provide
        subprocess.run(["gcc", "-c", "-pthread", "-include", "stdint.h", f"-I{STUBS}", f"-I{REF}/include",
                        f"{REF}/src/{name}.c", "-o", obj], check=True, capture_output=True)
        objs.append(obj)
    stub_obj = str(tmp_path / "stubs.o")
    subprocess.run(["gcc", "-c", f"-I{STUBS}", os.path.join(STUBS, "stubs.c"), "-o", stub_obj], check=True)
    exe = str(tmp_path / "vector_db_server")
    subprocess.run(["gcc", *objs, stub_obj, "-o", exe, f"-L{LIBDIR}", "-lsvdb_b200", f"-Wl,-rpath,{LIBDIR}", "-pthread", "-lm"],
                   check=True, capture_output=True)
    undefined = subprocess.run(["nm", "-u", exe], capture_output=True, text=True, check=True).stdout
    wanted = {line.split()[-1] for line in undefined.splitlines()} & L1_API
    assert wanted == L1_API, f"callers did not pick these from the library: {sorted(L1_API - wanted)}"
    ldd = subprocess.run(["ldd", exe], capture_output=True, text=True, check=True).stdout
    assert "libsvdb_b200.so" in ldd and "not found" not in ldd


def test_link_line_patch_applies_to_the_reference_makefile(tmp_path):
    """integration/link_line.patch is the whole binding a maintainer adds: the reference's Makefile stops compiling its own
    vector_database.c / kdtree.c and links libsvdb_b200.so instead.  Applied to a scratch copy, `make -n` must show exactly that."""
    import shutil
    import subprocess
    if not os.path.exists(os.path.join(REF, "Makefile")):
        pytest.skip("reference tree not present")
    if shutil.which("patch") is None or shutil.which("make") is None:
        pytest.skip("patch / make not installed")
    shutil.copy(os.path.join(REF, "Makefile"), tmp_path / "Makefile")
    os.makedirs(tmp_path / "src")
    for name in ("get_handler", "post_handler", "put_handler", "delete_handler", "compare_handler", "main", "vector_database", "kdtree"):
        (tmp_path / "src" / (name + ".c")).write_text("")
    patch = os.path.join(ROOT, "integration", "link_line.patch")
    subprocess.run(["patch", "-p1", "-i", patch], cwd=tmp_path, check=True, capture_output=True)
    out = subprocess.run(["make", "-n", "SVDB_B200=" + ROOT], cwd=tmp_path, check=True, capture_output=True, text=True).stdout
    link = [line for line in out.splitlines() if "-o executable/vector_db_server" in line]
    assert len(link) == 1
    assert "-lsvdb_b200" in link[0] and os.path.join(ROOT, "simple-vector-db_b200", "lib") in link[0]
    assert "kdtree.o" not in link[0] and "vector_database.o" not in link[0]
    assert "src/kdtree.c" not in out and "src/vector_database.c" not in out
    for name in ("get_handler", "post_handler", "put_handler", "delete_handler", "compare_handler", "main"):
        assert f"src/{name}.c" in out
