"""Build libsvdb_b200.so (CUDA kernels + C-ABI) for sm_100a with nvcc, in-tree.

    python simple-vector-db_b200/build.py [--force]

The result lands in simple-vector-db_b200/lib/ so that it travels with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libsvdb_b200.so")
SOURCES = ["scan_kernels.cu", "plane_scan.cu", "compare_kernels.cu", "tree_kernels.cu", "median_tree.cu", "mma_kernels.cu", "umma_filter.cu", "arena.cu", "engine.cu", "exchange.cu", "tie_protocol.cu", "dropin.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-cudart", "static"]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def _stamp() -> str:
    """Hash of everything the library depends on: sources, headers, the source list, the flags and the compiler."""
    h = hashlib.sha256()
    h.update(repr((SOURCES, ARCH, COMMON_FLAGS)).encode())
    try:
        h.update(subprocess.run([nvcc(), "--version"], capture_output=True, text=True).stdout.encode())
    except Exception:  # noqa: BLE001
        pass
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h")):
                h.update(name.encode())
                h.update(open(os.path.join(root, name), "rb").read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    stamp_file = os.path.join(LIBDIR, "build.stamp")
    stamp = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    cc = nvcc()
    common = [cc, *ARCH, *COMMON_FLAGS]
    if verbose:
        common += ["-Xptxas", "-v"]
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        objs.append(obj)
        procs.append((src, subprocess.Popen(common + ["-c", os.path.join(CSRC, src), "-o", obj],
                                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- {src} ---\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    link = [cc, *ARCH, "-shared", "-cudart", "static", "-o", LIB, *objs, "-Xlinker", "-Bsymbolic",
            "-lpthread", "-ldl", "-lrt"]
    subprocess.run(link, check=True)
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
