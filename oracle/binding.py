"""ctypes bindings for the CPU checkers.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module; the product package never does.

Two libraries, one uniform ``cpu_*`` driver surface:
  * oracle/_ref/libsvdb_ref.so  -- the reference's own src/kdtree.c +
    src/vector_database.c compiled unmodified (oracle/Makefile), kind "reference";
  * oracle/libsvdb_oracle.so    -- our restatement (svdb_oracle.c), kind "port".
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libsvdb_ref.so")
REF_O0_SO = os.path.join(HERE, "_ref", "libsvdb_ref_O0.so")
PORT_SO = os.path.join(HERE, "libsvdb_oracle.so")

UUID_SIZE = 37
NONE = (1 << 64) - 1  # (size_t)-1

_dp = C.POINTER(C.c_double)
_zp = C.POINTER(C.c_size_t)


def build(ref_root: str = "/root/reference") -> None:
    """Compile the checkers (port always; _ref when the reference tree is present)."""
    subprocess.run(["make", "-s", "-C", HERE, "all", f"REF={ref_root}"], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    # the reference's handlers driven in-process (needs the CUDA library to link the second binary)
    subprocess.run(["make", "-s", "-C", HERE, "handlers", f"REF={ref_root}"], check=False,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    # ... and the same handlers with integration/f3_f4_handlers.patch applied (top-k /nearest, batched /compare)
    subprocess.run(["make", "-s", "-C", HERE, "handlers_patched", f"REF={ref_root}"], check=False,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def _as_f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a: np.ndarray, t=_dp):
    return a.ctypes.data_as(t)


class Vector(C.Structure):
    """include/vector_database.h:14-18 of the reference."""
    _fields_ = [("uuid", C.c_char * UUID_SIZE), ("dimension", C.c_size_t), ("data", _dp)]


class KDTreeS(C.Structure):
    _fields_ = [("root", C.c_void_p), ("dimension", C.c_size_t)]


class VectorDatabaseS(C.Structure):
    """Prefix of include/vector_database.h:24-30 (mutex omitted: never touched here)."""
    _fields_ = [("vectors", C.POINTER(Vector)), ("size", C.c_size_t), ("capacity", C.c_size_t),
                ("kdtree", C.POINTER(KDTreeS))]


class CpuDriver:
    """The uniform cpu_* surface both libraries export."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.path = path
        self.lib = L = C.CDLL(path)
        L.cpu_build.restype = C.c_void_p
        L.cpu_build.argtypes = [_dp, C.c_size_t, C.c_size_t, C.c_size_t]
        L.cpu_free.argtypes = [C.c_void_p]
        L.cpu_nearest_batch.argtypes = [C.c_void_p, _dp, C.c_size_t, C.c_size_t, C.c_size_t, _zp]
        L.cpu_compare_batch.argtypes = [C.c_void_p, C.c_int, _zp, _zp, C.c_size_t, C.c_size_t,
                                        C.POINTER(C.c_float)]
        L.cpu_kind.restype = C.c_char_p
        self.kind = L.cpu_kind().decode()

    def build(self, rows, K: int):
        rows = _as_f64(rows)
        n, D = rows.shape
        h = self.lib.cpu_build(_ptr(rows), n, D, K)
        if not h:
            raise MemoryError("cpu_build failed")
        return h

    def free(self, h) -> None:
        self.lib.cpu_free(h)

    def nearest_batch(self, h, Q, nthreads: int = 1) -> np.ndarray:
        Q = _as_f64(Q)
        nq, stride = Q.shape
        out = np.empty(nq, dtype=np.uint64)
        self.lib.cpu_nearest_batch(h, _ptr(Q), nq, stride, nthreads, _ptr(out, _zp))
        return out

    def compare_batch(self, h, metric: int, i1, i2, nthreads: int = 1) -> np.ndarray:
        i1 = np.ascontiguousarray(i1, dtype=np.uint64)
        i2 = np.ascontiguousarray(i2, dtype=np.uint64)
        out = np.empty(len(i1), dtype=np.float32)
        self.lib.cpu_compare_batch(h, metric, _ptr(i1, _zp), _ptr(i2, _zp), len(i1), nthreads,
                                   _ptr(out, C.POINTER(C.c_float)))
        return out


class RefApi:
    """The reference's L1 C API (include/vector_database.h:39-135, include/kdtree.h:32-57)
    bound on ANY shared library that exports it: oracle/_ref (the reference itself) or the
    product's libsvdb_b200.so (the drop-in) -- the parity tests drive both through this class."""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.path = path
        self.lib = L = getattr(self, "lib", None) or C.CDLL(path)
        pdb = C.POINTER(VectorDatabaseS)
        L.vector_db_init.restype = pdb
        L.vector_db_init.argtypes = [C.c_size_t, C.c_size_t]
        L.vector_db_free.argtypes = [pdb]
        L.vector_db_insert.restype = C.c_size_t
        L.vector_db_insert.argtypes = [pdb, Vector]
        L.vector_db_read.restype = C.POINTER(Vector)
        L.vector_db_read.argtypes = [pdb, C.c_size_t]
        L.vector_db_read_by_uuid.restype = C.POINTER(Vector)
        L.vector_db_read_by_uuid.argtypes = [pdb, C.c_char_p]
        L.vector_db_update.argtypes = [pdb, C.c_size_t, Vector]
        L.vector_db_delete.argtypes = [pdb, C.c_size_t]
        L.vector_db_save.argtypes = [pdb, C.c_char_p]
        L.vector_db_load.restype = pdb
        L.vector_db_load.argtypes = [C.c_char_p, C.c_size_t]
        for f in (L.cosine_similarity, L.euclidean_distance, L.dot_product):
            f.restype = C.c_float
            f.argtypes = [Vector, Vector]
        L.kdtree_create.restype = C.POINTER(KDTreeS)
        L.kdtree_create.argtypes = [C.c_size_t]
        L.kdtree_insert.argtypes = [C.POINTER(KDTreeS), _dp, C.c_size_t]
        L.kdtree_free.argtypes = [C.POINTER(KDTreeS)]
        L.kdtree_nearest.restype = C.c_size_t
        L.kdtree_nearest.argtypes = [C.POINTER(KDTreeS), _dp]
        self._libc = C.CDLL(None)
        self._libc.malloc.restype = C.c_void_p
        self._libc.malloc.argtypes = [C.c_size_t]

    def make_vector(self, data, uuid: str = "", own: bool = True) -> Vector:
        """A Vector whose data is malloc'ed (insert/update take ownership: post_handler.c:244)."""
        data = _as_f64(data)
        v = Vector()
        v.uuid = uuid.encode()[:36]
        v.dimension = len(data)
        if own:
            p = self._libc.malloc(max(1, data.nbytes))
            C.memmove(p, data.ctypes.data, data.nbytes)
            v.data = C.cast(p, _dp)
        else:
            v._keep = data
            v.data = _ptr(data)
        return v

    def metric(self, which: int, a, b) -> np.float32:
        va, vb = self.make_vector(a, own=False), self.make_vector(b, own=False)
        f = (self.lib.cosine_similarity, self.lib.euclidean_distance, self.lib.dot_product)[which]
        return np.float32(f(va, vb))

    def nearest(self, db, q) -> int:
        q = _as_f64(q)
        return self.lib.kdtree_nearest(db.contents.kdtree, _ptr(q))


class RefLib(CpuDriver, RefApi):
    """oracle/_ref: the reference's own code plus our cpu_* batch driver."""

    def __init__(self, path: str = REF_SO):
        CpuDriver.__init__(self, path)
        RefApi.__init__(self, path)


class PortLib(CpuDriver):
    """Our restatement's own entry points (oracle/svdb_oracle.c)."""

    def __init__(self, path: str = PORT_SO):
        super().__init__(path)
        L = self.lib
        L.orc_log_create.restype = C.c_void_p
        L.orc_log_create.argtypes = [C.c_size_t]
        L.orc_log_free.argtypes = [C.c_void_p]
        L.orc_log_size.restype = C.c_size_t
        L.orc_log_size.argtypes = [C.c_void_p]
        L.orc_log_append.restype = C.c_int64
        L.orc_log_append.argtypes = [C.c_void_p, _dp, C.c_size_t]
        L.orc_sqdist.restype = C.c_double
        L.orc_sqdist.argtypes = [_dp, _dp, C.c_size_t]
        L.orc_tree_nearest.restype = C.c_size_t
        L.orc_tree_nearest.argtypes = [C.c_void_p, _dp]
        L.orc_tree_nearest_seq.restype = C.c_int64
        L.orc_tree_nearest_seq.argtypes = [C.c_void_p, _dp, _dp, _zp]
        L.orc_flat_topk.restype = C.c_size_t
        L.orc_flat_topk.argtypes = [C.c_void_p, _dp, C.c_size_t, C.POINTER(C.c_int64), _zp, _dp]
        for f in (L.orc_cosine_similarity, L.orc_euclidean_distance, L.orc_dot_product, L.orc_dot_f):
            f.restype = C.c_float
            f.argtypes = [_dp, _dp, C.c_size_t]
        L.orc_db_create.restype = C.c_void_p
        L.orc_db_create.argtypes = [C.c_size_t, C.c_size_t]
        L.orc_db_free.argtypes = [C.c_void_p]
        L.orc_db_size.restype = C.c_size_t
        L.orc_db_size.argtypes = [C.c_void_p]
        L.orc_db_log.restype = C.c_void_p
        L.orc_db_log.argtypes = [C.c_void_p]
        L.orc_db_row.restype = _dp
        L.orc_db_row.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_db_insert.restype = C.c_size_t
        L.orc_db_insert.argtypes = [C.c_void_p, _dp]
        L.orc_db_update.argtypes = [C.c_void_p, C.c_size_t, _dp]
        L.orc_db_delete.argtypes = [C.c_void_p, C.c_size_t]
        L.orc_db_nearest.restype = C.c_size_t
        L.orc_db_nearest.argtypes = [C.c_void_p, _dp]
        L.orc_db_compare.restype = C.c_float
        L.orc_db_compare.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t]

    # -- convenience ------------------------------------------------------
    def metric(self, which: int, a, b) -> np.float32:
        a, b = _as_f64(a), _as_f64(b)
        f = (self.lib.orc_cosine_similarity, self.lib.orc_euclidean_distance, self.lib.orc_dot_product)[which]
        return np.float32(f(_ptr(a), _ptr(b), len(a)))

    def selfdot_f(self, a) -> np.float32:
        a = _as_f64(a)
        return np.float32(self.lib.orc_dot_f(_ptr(a), _ptr(a), len(a)))

    def sqdist(self, p, q) -> float:
        p, q = _as_f64(p), _as_f64(q)
        return float(self.lib.orc_sqdist(_ptr(p), _ptr(q), min(len(p), len(q))))

    def flat_topk(self, log, q, k: int):
        q = _as_f64(q)
        seq = np.empty(k, dtype=np.int64)
        idx = np.empty(k, dtype=np.uint64)
        d = np.empty(k, dtype=np.float64)
        m = self.lib.orc_flat_topk(log, _ptr(q), k, _ptr(seq, C.POINTER(C.c_int64)), _ptr(idx, _zp), _ptr(d))
        return seq[:m], idx[:m], d[:m]

    def tree_nearest_seq(self, log, q):
        q = _as_f64(q)
        best = C.c_double()
        visited = C.c_size_t()
        s = self.lib.orc_tree_nearest_seq(log, _ptr(q), C.byref(best), C.byref(visited))
        return int(s), best.value, visited.value


class PortDB:
    """Store-semantics model (insert / update / delete / nearest / compare) on the port."""

    def __init__(self, port: PortLib, D: int, K: int):
        self.p, self.D, self.K = port, D, K
        self.h = port.lib.orc_db_create(D, K)

    def close(self):
        if self.h:
            self.p.lib.orc_db_free(self.h)
            self.h = None

    __del__ = close

    @property
    def size(self) -> int:
        return self.p.lib.orc_db_size(self.h)

    @property
    def log(self):
        return self.p.lib.orc_db_log(self.h)

    def insert(self, v) -> int:
        v = _as_f64(v)
        return self.p.lib.orc_db_insert(self.h, _ptr(v))

    def update(self, i: int, v) -> None:
        v = _as_f64(v)
        self.p.lib.orc_db_update(self.h, i, _ptr(v))

    def delete(self, i: int) -> None:
        self.p.lib.orc_db_delete(self.h, i)

    def nearest(self, q) -> int:
        q = _as_f64(q)
        return self.p.lib.orc_db_nearest(self.h, _ptr(q))

    def topk(self, q, k: int):
        return self.p.flat_topk(self.log, q, k)

    def compare(self, metric: int, i1: int, i2: int) -> np.float32:
        return np.float32(self.p.lib.orc_db_compare(self.h, metric, i1, i2))

    def row(self, i: int) -> np.ndarray:
        p = self.p.lib.orc_db_row(self.h, i)
        return np.ctypeslib.as_array(p, shape=(self.D,)).copy()


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def load_port() -> PortLib:
    if not os.path.exists(PORT_SO):
        build()
    return PortLib()


def load_ref(o0: bool = False) -> RefLib:
    return RefLib(REF_O0_SO if o0 else REF_SO)


def load_cpu_driver() -> CpuDriver:
    """Prefer the compiled reference (kind 'reference'), else the port."""
    return load_ref() if have_ref() else load_port()
