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
