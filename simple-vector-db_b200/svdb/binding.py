"""ctypes binding of libsvdb_b200.so (include/svdb_b200.h, include/svdb_dropin.h).

This is the Python mirror of the C boundary; it contains no compute.  If the CUDA
library has not been built, importing the engine raises -- there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(HERE), "lib", "libsvdb_b200.so")

NONE = (1 << 64) - 1
MAX_K = 24
COSINE, EUCLIDEAN, DOT, ALL_METRICS = 0, 1, 2, 3
FLAG_LOG_ONLY, FLAG_NO_LOG, FLAG_SHARD = 1, 2, 4
MODE_AUTO, MODE_EXACT, MODE_TREE, MODE_MTREE, MODE_FP64 = 0, 1, 2, 3, 4
CAND_UNSAFE, CAND_TIE = 1, 2

_dp = C.POINTER(C.c_double)
_zp = C.POINTER(C.c_size_t)
_fp = C.POINTER(C.c_float)
_u64p = C.POINTER(C.c_uint64)

candidate_dtype = np.dtype([("dist", "<f8"), ("seq", "<u8"), ("index", "<u8"), ("flags", "<u8")])


class Config(C.Structure):
    _fields_ = [("dimension", C.c_size_t), ("kd_dim", C.c_size_t), ("device", C.c_int),
                ("seq_base", C.c_uint64), ("reserve_rows", C.c_size_t), ("flags", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("kernels_launched", C.c_uint64), ("exact_reruns", C.c_uint64), ("tree_reruns", C.c_uint64),
                ("tree_rounds", C.c_uint64), ("coalesced_calls", C.c_uint64), ("coalesced_passes", C.c_uint64),
                ("hbm_bytes_mapped", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("tie_events", C.c_uint64), ("tie_levels", C.c_uint64), ("mtree_builds", C.c_uint64),
                ("mtree_levels", C.c_uint64), ("mtree_rows", C.c_uint64), ("fp64_reruns", C.c_uint64),
                ("scan_plane_last", C.c_uint64), ("tree_dropped", C.c_uint64)]


# ---- exact ties on a sharded store (svdb_tie_resolve / svdb_resolve_ties_sharded) ----
tie_first_dtype = np.dtype([("seq", "<u8"), ("index", "<u8"), ("v", "<f8"), ("tied", "<u8")])
tie_split_dtype = np.dtype([("n", "<u8", (2,)), ("min_seq", "<u8", (2,)), ("min_index", "<u8", (2,))])
ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t)
TIE_COLLECT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p)
TIE_FIRST_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                           C.c_void_p, C.c_void_p)
TIE_SPLIT_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                           C.c_void_p, C.c_void_p)


class TieBackend(C.Structure):
    _fields_ = [("ctx", C.c_void_p), ("world", C.c_int), ("rank", C.c_int), ("kd_dim", C.c_size_t),
                ("allgather", ALLGATHER_FN), ("allgather_ctx", C.c_void_p), ("collect", TIE_COLLECT_FN),
                ("first_in_cell", TIE_FIRST_FN), ("split", TIE_SPLIT_FN)]


class SvdbError(RuntimeError):
    pass


_lib = None

# every symbol include/svdb_b200.h declares
NATIVE_SYMBOLS = [
    "svdb_last_error", "svdb_version", "svdb_device_count", "svdb_engine_create", "svdb_engine_destroy",
    "svdb_set_stream", "svdb_insert_batch", "svdb_update_batch", "svdb_delete_batch", "svdb_append_kdpoints",
    "svdb_insert_batch_device", "svdb_append_kdpoints_device", "svdb_flush", "svdb_size", "svdb_log_size", "svdb_dimension", "svdb_kd_dim",
    "svdb_read_row", "svdb_nearest_batch", "svdb_nearest_batch_device", "svdb_merge_candidates_device",
    "svdb_compare_batch", "svdb_compare_batch_all", "svdb_compare_batch_device", "svdb_compare_vectors",
    "svdb_get_stats", "svdb_set_option", "svdb_time_scan", "svdb_take_scan_time", "svdb_debug_filter_keys",
    "svdb_engine_load_file", "svdb_save_file", "svdb_get_uuid", "svdb_set_uuid",
    "svdb_exchange_create", "svdb_exchange_connect", "svdb_exchange_destroy", "svdb_exchange_merge",
    "svdb_nearest_batch_sharded", "svdb_tie_resolve", "svdb_resolve_ties_sharded", "svdb_nearest_batch_device_sharded", "svdb_debug_tail_times", "svdb_debug_plane8",
]
# every symbol include/svdb_dropin.h declares (the reference's L1 API + two batched extensions)
DROPIN_SYMBOLS = [
    "kdtree_create", "kdtree_insert", "kdtree_free", "kdtree_nearest", "vector_db_init", "vector_db_free",
    "vector_db_insert", "vector_db_read", "vector_db_read_by_uuid", "vector_db_update", "vector_db_delete",
    "vector_db_save", "vector_db_load", "cosine_similarity", "euclidean_distance", "dot_product",
    "kdtree_nearest_batch", "vector_db_compare_batch",
]


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SvdbError(f"{LIB_PATH} is missing: build it with `python simple-vector-db_b200/build.py` "
                        "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.svdb_last_error.restype = C.c_char_p
    L.svdb_version.restype = C.c_char_p
    L.svdb_engine_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
    L.svdb_engine_destroy.argtypes = [C.c_void_p]
    L.svdb_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.svdb_insert_batch.argtypes = [C.c_void_p, _dp, C.c_size_t, C.c_size_t, _zp]
    L.svdb_update_batch.argtypes = [C.c_void_p, _zp, _dp, C.c_size_t, C.c_size_t]
    L.svdb_delete_batch.argtypes = [C.c_void_p, _zp, C.c_size_t]
    L.svdb_append_kdpoints.argtypes = [C.c_void_p, _dp, _zp, C.c_size_t, C.c_size_t]
    L.svdb_insert_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, _zp]
    L.svdb_append_kdpoints_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t]
    L.svdb_flush.argtypes = [C.c_void_p]
    for f in (L.svdb_size, L.svdb_log_size, L.svdb_dimension, L.svdb_kd_dim):
        f.restype = C.c_size_t
        f.argtypes = [C.c_void_p]
    L.svdb_read_row.argtypes = [C.c_void_p, C.c_size_t, _dp]
    L.svdb_nearest_batch.argtypes = [C.c_void_p, _dp, C.c_size_t, C.c_size_t, C.c_size_t, _zp, _dp, _u64p]
    L.svdb_nearest_batch_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                            C.c_void_p, C.c_int]
    L.svdb_merge_candidates_device.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t,
                                               C.c_size_t, C.c_void_p]
    L.svdb_compare_batch.argtypes = [C.c_void_p, C.c_int, _zp, _zp, C.c_size_t, _fp]
    L.svdb_compare_batch_all.argtypes = [C.c_void_p, _zp, _zp, C.c_size_t, _fp]
    L.svdb_compare_batch_device.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    L.svdb_compare_vectors.argtypes = [C.c_int, C.c_int, _dp, _dp, C.c_size_t, _fp]
    L.svdb_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    L.svdb_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_long]
    L.svdb_time_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, _fp]
    L.svdb_take_scan_time.argtypes = [C.c_void_p, _fp, _u64p]
    L.svdb_debug_filter_keys.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.svdb_debug_tail_times.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.svdb_debug_plane8.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    L.svdb_engine_load_file.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_uint32, C.POINTER(C.c_void_p)]
    L.svdb_save_file.argtypes = [C.c_void_p, C.c_char_p]
    L.svdb_exchange_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_size_t, C.POINTER(C.c_void_p), C.c_char_p]
    L.svdb_exchange_connect.argtypes = [C.c_void_p, C.c_char_p]
    L.svdb_exchange_destroy.argtypes = [C.c_void_p]
    L.svdb_exchange_merge.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]
    L.svdb_nearest_batch_sharded.argtypes = [C.c_void_p, C.c_void_p, _dp, C.c_size_t, C.c_size_t, C.c_size_t, _zp, _dp, _u64p]
    L.svdb_nearest_batch_device_sharded.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                                    C.c_void_p, C.c_int]
    L.svdb_tie_resolve.argtypes = [C.POINTER(TieBackend), C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, _u64p]
    L.svdb_resolve_ties_sharded.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, ALLGATHER_FN, C.c_void_p, C.c_void_p,
                                            C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t]
    L.svdb_get_uuid.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p]
    L.svdb_set_uuid.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p]
    _lib = L
    return L


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise SvdbError(f"{what} failed ({rc}): {lib().svdb_last_error().decode()}")


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _u64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64)


class Engine:
    """One shard of the store in one GPU's HBM (svdb_engine)."""

    def __init__(self, dimension: int, kd_dim: int | None = None, device: int = 0, seq_base: int = 0,
                 reserve_rows: int = 0, flags: int = 0):
        self.L = lib()
        cfg = Config(dimension, kd_dim if kd_dim is not None else dimension, device, seq_base, reserve_rows, flags)
        h = C.c_void_p()
        _check(self.L.svdb_engine_create(C.byref(cfg), C.byref(h)), "svdb_engine_create")
        self.h = h
        self.D, self.K, self.device = dimension, cfg.kd_dim, device

    @classmethod
    def load_file(cls, path: str, kd_dim: int, device: int = 0, flags: int = 0) -> "Engine":
        """Bulk load of the reference's save file straight into HBM (svdb_engine_load_file)."""
        self = cls.__new__(cls)
        self.L = lib()
        h = C.c_void_p()
        _check(self.L.svdb_engine_load_file(path.encode(), kd_dim, device, flags, C.byref(h)), "svdb_engine_load_file")
        self.h = h
        self.D, self.K, self.device = self.L.svdb_dimension(h), self.L.svdb_kd_dim(h), device
        return self

    def save_file(self, path: str) -> None:
        _check(self.L.svdb_save_file(self.h, path.encode()), "svdb_save_file")

    def get_uuid(self, index: int) -> str:
        buf = C.create_string_buffer(37)
        _check(self.L.svdb_get_uuid(self.h, index, buf), "svdb_get_uuid")
        return buf.value.decode()

    def set_uuid(self, index: int, uuid: str) -> None:
        _check(self.L.svdb_set_uuid(self.h, index, uuid.encode()), "svdb_set_uuid")

    def close(self) -> None:
        if getattr(self, "h", None):
            self.L.svdb_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- store --------------------------------------------------------------
    def insert(self, rows) -> int:
        rows = _f64(rows).reshape(-1, self.D)
        first = C.c_size_t()
        _check(self.L.svdb_insert_batch(self.h, rows.ctypes.data_as(_dp), len(rows), self.D, C.byref(first)),
               "svdb_insert_batch")
        return first.value

    def insert_device(self, ptr: int, n: int, ld: int) -> int:
        first = C.c_size_t()
        _check(self.L.svdb_insert_batch_device(self.h, C.c_void_p(ptr), n, ld, C.byref(first)),
               "svdb_insert_batch_device")
        return first.value

    def append_kdpoints_device(self, ptr: int, first_index: int, n: int, ld: int) -> None:
        _check(self.L.svdb_append_kdpoints_device(self.h, C.c_void_p(ptr), first_index, n, ld), "svdb_append_kdpoints_device")

    def update(self, index, rows) -> None:
        index = _u64(np.atleast_1d(index))
        rows = _f64(rows).reshape(-1, self.D)
        _check(self.L.svdb_update_batch(self.h, index.ctypes.data_as(_zp), rows.ctypes.data_as(_dp), len(index), self.D),
               "svdb_update_batch")

    def delete(self, index) -> None:
        index = _u64(np.atleast_1d(index))
        _check(self.L.svdb_delete_batch(self.h, index.ctypes.data_as(_zp), len(index)), "svdb_delete_batch")

    def append_kdpoints(self, pts, index) -> None:
        pts = _f64(pts)
        pts = pts.reshape(-1, pts.shape[-1])
        index = _u64(np.atleast_1d(index))
        _check(self.L.svdb_append_kdpoints(self.h, pts.ctypes.data_as(_dp), index.ctypes.data_as(_zp), len(index),
                                           pts.shape[1]), "svdb_append_kdpoints")

    def flush(self) -> None:
        _check(self.L.svdb_flush(self.h), "svdb_flush")

    @property
    def size(self) -> int:
        return self.L.svdb_size(self.h)

    @property
    def log_size(self) -> int:
        return self.L.svdb_log_size(self.h)

    def read_row(self, index: int) -> np.ndarray:
        out = np.empty(self.D, dtype=np.float64)
        _check(self.L.svdb_read_row(self.h, index, out.ctypes.data_as(_dp)), "svdb_read_row")
        return out

    # -- nearest ------------------------------------------------------------
    def nearest(self, Q, k: int = 1):
        """(index, dist, seq), each nq x k, for host queries Q (nq x >=K)."""
        Q = _f64(Q)
        Q = Q.reshape(-1, Q.shape[-1])
        nq = len(Q)
        idx = np.empty((nq, k), dtype=np.uint64)
        dist = np.empty((nq, k), dtype=np.float64)
        seq = np.empty((nq, k), dtype=np.uint64)
        _check(self.L.svdb_nearest_batch(self.h, Q.ctypes.data_as(_dp), nq, Q.shape[1], k, idx.ctypes.data_as(_zp),
                                         dist.ctypes.data_as(_dp), seq.ctypes.data_as(_u64p)), "svdb_nearest_batch")
        return idx, dist, seq

    def nearest_sharded(self, xch: "Exchange", Q, k: int = 1):
        """This shard's part of a sharded query (svdb_nearest_batch_sharded): merged (index, dist, seq)."""
        Q = _f64(Q)
        Q = Q.reshape(-1, Q.shape[-1])
        nq = len(Q)
        idx = np.empty((nq, k), dtype=np.uint64)
        dist = np.empty((nq, k), dtype=np.float64)
        seq = np.empty((nq, k), dtype=np.uint64)
        _check(self.L.svdb_nearest_batch_sharded(self.h, xch.h, Q.ctypes.data_as(_dp), nq, Q.shape[1], k,
                                                 idx.ctypes.data_as(_zp), dist.ctypes.data_as(_dp),
                                                 seq.ctypes.data_as(_u64p)), "svdb_nearest_batch_sharded")
        return idx, dist, seq

    def resolve_ties_sharded(self, rank: int, world: int, Q, merged: np.ndarray, allgather=None, xch: "Exchange" = None):
        """Position 0 of every SVDB_CAND_TIE-flagged row of `merged` (nq x k candidates, the same on every rank)
        becomes the entry the reference's global tree reaches first; in place.  Collective.  allgather(send, recv):
        numpy uint8 arrays, recv = world blocks of len(send) in rank order; or pass the peer-memory exchange."""
        Q = _f64(Q)
        Q = Q.reshape(-1, Q.shape[-1])
        assert merged.dtype == candidate_dtype and merged.flags["C_CONTIGUOUS"] and merged.shape[0] == len(Q)
        cb = make_allgather_cb(allgather, world) if allgather is not None else C.cast(None, ALLGATHER_FN)
        _check(self.L.svdb_resolve_ties_sharded(self.h, xch.h if xch is not None else None, rank, world, cb, None,
                                                C.c_void_p(Q.ctypes.data), len(Q), Q.shape[1],
                                                C.c_void_p(merged.ctypes.data), merged.shape[1]),
               "svdb_resolve_ties_sharded")
        return merged

    def nearest_device(self, q_ptr: int, nq: int, ldq: int, k: int, out_ptr: int, mode: int = 0) -> None:
        _check(self.L.svdb_nearest_batch_device(self.h, C.c_void_p(q_ptr), nq, ldq, k, C.c_void_p(out_ptr), int(mode)),
               "svdb_nearest_batch_device")

    def nearest_device_sharded(self, xch: "Exchange", q_ptr: int, nq: int, ldq: int, k: int, out_ptr: int, mode: int = 0) -> None:
        """Scan of this shard + peer-memory exchange + merge (collective); out receives the merged candidates."""
        _check(self.L.svdb_nearest_batch_device_sharded(self.h, xch.h, C.c_void_p(q_ptr), nq, ldq, k, C.c_void_p(out_ptr),
                                                        int(mode)), "svdb_nearest_batch_device_sharded")

    # -- compare ------------------------------------------------------------
    def compare(self, metric: int, i1, i2) -> np.ndarray:
        i1, i2 = _u64(i1), _u64(i2)
        n = len(i1)
        if metric == ALL_METRICS:
            out = np.empty((n, 3), dtype=np.float32)
            _check(self.L.svdb_compare_batch_all(self.h, i1.ctypes.data_as(_zp), i2.ctypes.data_as(_zp), n,
                                                 out.ctypes.data_as(_fp)), "svdb_compare_batch_all")
            return out
        out = np.empty(n, dtype=np.float32)
        _check(self.L.svdb_compare_batch(self.h, metric, i1.ctypes.data_as(_zp), i2.ctypes.data_as(_zp), n,
                                         out.ctypes.data_as(_fp)), "svdb_compare_batch")
        return out

    def compare_device(self, metric: int, i1_ptr: int, i2_ptr: int, n: int, out_ptr: int) -> None:
        _check(self.L.svdb_compare_batch_device(self.h, metric, C.c_void_p(i1_ptr), C.c_void_p(i2_ptr), n,
                                                C.c_void_p(out_ptr)), "svdb_compare_batch_device")

    # -- plumbing -----------------------------------------------------------
    def set_stream(self, stream_handle: int | None) -> None:
        """A cudaStream_t handle (0 = CUDA's legacy default stream); None = the engine's own stream."""
        h = C.c_void_p(-1) if stream_handle is None else C.c_void_p(stream_handle)
        _check(self.L.svdb_set_stream(self.h, h), "svdb_set_stream")

    def set_option(self, name: str, value: int) -> None:
        _check(self.L.svdb_set_option(self.h, name.encode(), int(value)), "svdb_set_option")

    def stats(self) -> dict:
        s = Stats()
        _check(self.L.svdb_get_stats(self.h, C.byref(s)), "svdb_get_stats")
        return {k: getattr(s, k) for k, _ in Stats._fields_}

    def time_scan(self, q_ptr: int, nq: int, ldq: int, k: int, iters: int) -> float:
        ms = C.c_float()
        _check(self.L.svdb_time_scan(self.h, C.c_void_p(q_ptr), nq, ldq, k, iters, C.byref(ms)), "svdb_time_scan")
        return ms.value

    def debug_filter_cycles(self) -> np.ndarray:
        """16 floats behind the K10 key dump: kilocycles CTA 0 spent per role (csrc/umma_filter.cu, option umma.debug_keys)."""
        out = np.empty(128 * 256 + 16, dtype=np.float32)
        _check(self.L.svdb_debug_filter_keys(self.h, C.c_void_p(out.ctypes.data), out.size), "svdb_debug_filter_keys")
        return out[128 * 256:]

    def debug_filter_keys(self, bn: int) -> np.ndarray:
        out = np.empty((128, bn), dtype=np.float32)
        _check(self.L.svdb_debug_filter_keys(self.h, C.c_void_p(out.ctypes.data), out.size), "svdb_debug_filter_keys")
        return out

    def debug_plane8(self, nbytes: int = 0):
        """K13's grid (lo, step), the measured plane error, and the first nbytes bytes of the plane."""
        par = np.zeros(3, dtype=np.float64)
        raw = np.zeros(max(1, nbytes), dtype=np.uint8)
        _check(self.L.svdb_debug_plane8(self.h, C.c_void_p(par.ctypes.data), C.c_void_p(raw.ctypes.data), nbytes), "svdb_debug_plane8")
        return par[0], par[1], par[2], raw[:nbytes]

    def debug_tail_times(self, n_ctas: int = 296) -> np.ndarray:
        out = np.zeros(32 + n_ctas, dtype=np.uint64)
        _check(self.L.svdb_debug_tail_times(self.h, C.c_void_p(out.ctypes.data), out.size), "svdb_debug_tail_times")
        return out

    def take_scan_time(self):
        ms = C.c_float()
        n = C.c_uint64()
        _check(self.L.svdb_take_scan_time(self.h, C.byref(ms), C.byref(n)), "svdb_take_scan_time")
        return ms.value, n.value


def compare_vectors(metric: int, a, b, device: int = 0) -> np.float32:
    a, b = _f64(a), _f64(b)
    out = C.c_float()
    _check(lib().svdb_compare_vectors(device, metric, a.ctypes.data_as(_dp), b.ctypes.data_as(_dp), len(a), C.byref(out)),
           "svdb_compare_vectors")
    return np.float32(out.value)


def _np_at(ptr, dtype, count):
    """numpy view of `count` items of `dtype` at a raw address handed to a callback."""
    if count == 0:
        return np.empty(0, dtype=dtype)
    dt = np.dtype(dtype)
    buf = (C.c_ubyte * (count * dt.itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dt, count=count)


def make_allgather_cb(fn, world: int):
    """Wrap fn(send_u8, recv_u8) as a C svdb_allgather_fn; exceptions become a non-zero return code."""
    def cb(_ctx, send, recv, nbytes):
        try:
            fn(_np_at(send, np.uint8, nbytes), _np_at(recv, np.uint8, nbytes * world))
            return 0
        except Exception as ex:                  # never unwind through the C frames
            import traceback
            traceback.print_exc()
            return -5
    return ALLGATHER_FN(cb)


def tie_resolve(backend: TieBackend, Q, merged: np.ndarray) -> int:
    """svdb_tie_resolve with a caller-supplied backend (tests drive it with a numpy backend). Returns levels walked."""
    Q = _f64(Q)
    Q = Q.reshape(-1, Q.shape[-1])
    levels = C.c_uint64()
    _check(lib().svdb_tie_resolve(C.byref(backend), C.c_void_p(Q.ctypes.data), len(Q), Q.shape[1],
                                  C.c_void_p(merged.ctypes.data), merged.shape[1], C.byref(levels)), "svdb_tie_resolve")
    return levels.value


def merge_candidates_device(device: int, stream: int | None, in_ptr: int, nshards: int, nq: int, k: int,
                            out_ptr: int) -> None:
    _check(lib().svdb_merge_candidates_device(device, C.c_void_p(stream or 0), C.c_void_p(in_ptr), nshards, nq, k,
                                              C.c_void_p(out_ptr)), "svdb_merge_candidates_device")


class Exchange:
    """Peer-memory candidate exchange (svdb_exchange): create -> all-gather handles -> connect."""

    def __init__(self, device: int, rank: int, world: int, max_records: int):
        self.L = lib()
        h = C.c_void_p()
        buf = C.create_string_buffer(64)
        _check(self.L.svdb_exchange_create(device, rank, world, max_records, C.byref(h), buf), "svdb_exchange_create")
        self.h, self.handle, self.world = h, buf.raw, world

    def connect(self, all_handles: bytes) -> None:
        assert len(all_handles) == 64 * self.world
        _check(self.L.svdb_exchange_connect(self.h, all_handles), "svdb_exchange_connect")

    def merge(self, stream: int, local_ptr: int, nq: int, k: int, out_ptr: int) -> None:
        _check(self.L.svdb_exchange_merge(self.h, C.c_void_p(stream), C.c_void_p(local_ptr), nq, k, C.c_void_p(out_ptr)),
               "svdb_exchange_merge")

    def close(self) -> None:
        if getattr(self, "h", None):
            self.L.svdb_exchange_destroy(self.h)
            self.h = None
