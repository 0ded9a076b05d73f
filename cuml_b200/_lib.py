"""ctypes binding of the C-ABI (include/cuml_b200/kmeans_c.h).

This is the stub the reference's Cython layer (python/cuml/cuml/cluster/cpp/kmeans.pxd:15-174)
would be replaced by.  There is no fallback: if the shared library is missing or cannot be
loaded, importing the estimator raises.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_LIB = None


class KMeansParams(C.Structure):
    """C mirror of ML::kmeans::KMeansParams (reference cpp/include/cuml/cluster/kmeans_params.hpp:17-32)."""
    _fields_ = [
        ("metric", C.c_int32),
        ("n_clusters", C.c_int32),
        ("init", C.c_int32),
        ("max_iter", C.c_int32),
        ("tol", C.c_double),
        ("verbosity", C.c_int32),
        ("rng_seed", C.c_uint64),
        ("rng_base_subsequence", C.c_uint64),
        ("rng_type", C.c_int32),
        ("n_init", C.c_int32),
        ("oversampling_factor", C.c_double),
        ("batch_samples", C.c_int32),
        ("batch_centroids", C.c_int32),
        ("init_size", C.c_int64),
        ("device_buffer_samples", C.c_int64),
    ]


INIT_KMEANS_PLUS_PLUS, INIT_RANDOM, INIT_ARRAY = 0, 1, 2
L2_EXPANDED, L2_SQRT_EXPANDED = 0, 1

# every symbol include/cuml_b200/kmeans_c.h declares
EXPORTED_SYMBOLS = [
    "cuml_b200_kmeans_params_default", "cuml_b200_handle_create", "cuml_b200_handle_destroy",
    "cuml_b200_handle_sync", "cuml_b200_handle_stream", "cuml_b200_last_error", "cuml_b200_version",
    "cuml_b200_nccl_unique_id", "cuml_b200_handle_init_comm",
    "cuml_b200_peer_window_create", "cuml_b200_peer_window_attach",
    "cuml_b200_kmeans_fit_f32_i32", "cuml_b200_kmeans_fit_f64_i32", "cuml_b200_kmeans_fit_f32_i64",
    "cuml_b200_kmeans_fit_f64_i64", "cuml_b200_kmeans_fit_parts_f32", "cuml_b200_kmeans_fit_parts_f64",
    "cuml_b200_kmeans_fit_parts_labels_f32", "cuml_b200_kmeans_fit_parts_labels_f64",
    "cuml_b200_kmeans_predict_f32_i32", "cuml_b200_kmeans_predict_f64_i32", "cuml_b200_kmeans_predict_f32_i64",
    "cuml_b200_kmeans_predict_f64_i64", "cuml_b200_kmeans_transform_f32_i32", "cuml_b200_kmeans_transform_f64_i32",
    "cuml_b200_kmeans_transform_f32_i64", "cuml_b200_kmeans_transform_f64_i64",
    "cuml_b200_kmeans_lloyd_step_f32", "cuml_b200_kmeans_assign_f32", "cuml_b200_launch_count_reset",
    "cuml_b200_launch_count", "cuml_b200_kernel_timing_enable", "cuml_b200_kernel_timing_read",
    "cuml_b200_kmeans_tc_supported", "cuml_b200_kmeans_debug_dots_f32", "cuml_b200_kmeans_estep_variant",
    "cuml_b200_kmeans_fused_update",
]


class CumlB200Error(RuntimeError):
    pass


def lib_path():
    return _build.lib_path()


def load(build_if_missing=True):
    """Load libcuml_b200.so (building it in-tree first if it is missing)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        if not build_if_missing:
            raise CumlB200Error("libcuml_b200.so is missing: run `python -m cuml_b200.build`")
        _build.build()
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    P = C.POINTER
    lib.cuml_b200_last_error.restype = C.c_char_p
    lib.cuml_b200_version.restype = C.c_char_p
    lib.cuml_b200_handle_stream.restype = vp
    lib.cuml_b200_handle_stream.argtypes = [vp]
    lib.cuml_b200_kmeans_params_default.argtypes = [P(KMeansParams)]
    lib.cuml_b200_kmeans_params_default.restype = None
    lib.cuml_b200_handle_create.argtypes = [P(vp), vp, vp, C.c_int, C.c_int]
    lib.cuml_b200_handle_destroy.argtypes = [vp]
    lib.cuml_b200_handle_sync.argtypes = [vp]
    lib.cuml_b200_nccl_unique_id.argtypes = [vp]
    lib.cuml_b200_handle_init_comm.argtypes = [vp, vp, C.c_int, C.c_int]
    lib.cuml_b200_peer_window_create.argtypes = [vp, C.c_int, C.c_size_t, vp]
    lib.cuml_b200_peer_window_attach.argtypes = [vp, vp, C.c_int, C.c_int]
    for t in ("f32", "f64"):
        for ix, it in (("i32", i32), ("i64", i64)):
            getattr(lib, f"cuml_b200_kmeans_fit_{t}_{ix}").argtypes = [vp, P(KMeansParams), vp, it, it, vp, vp, vp, vp]
            getattr(lib, f"cuml_b200_kmeans_predict_{t}_{ix}").argtypes = [vp, P(KMeansParams), vp, vp, it, it, vp,
                                                                           C.c_int, vp, vp]
            getattr(lib, f"cuml_b200_kmeans_transform_{t}_{ix}").argtypes = [vp, P(KMeansParams), vp, vp, it, it, vp]
        getattr(lib, f"cuml_b200_kmeans_fit_parts_{t}").argtypes = [vp, P(KMeansParams), vp, vp, i64, i64, vp, vp, vp, vp]
        getattr(lib, f"cuml_b200_kmeans_fit_parts_labels_{t}").argtypes = [vp, P(KMeansParams), vp, vp, i64, i64, vp, vp,
                                                                           vp, vp, vp]
    lib.cuml_b200_kmeans_lloyd_step_f32.argtypes = [vp, vp, i64, i64, vp, i32, vp, vp, vp, vp, C.c_int]
    lib.cuml_b200_kmeans_assign_f32.argtypes = [vp, vp, i64, i64, i32, vp, vp, C.c_int]
    lib.cuml_b200_kmeans_debug_dots_f32.argtypes = [vp, vp, i64, i64, i32, vp, vp, vp, P(i64)]
    lib.cuml_b200_launch_count.restype = i64
    lib.cuml_b200_launch_count_reset.restype = None
    lib.cuml_b200_kernel_timing_enable.argtypes = [vp, C.c_int]
    lib.cuml_b200_kernel_timing_read.argtypes = [vp, P(dbl), P(i64), P(dbl), P(i64)]
    lib.cuml_b200_kmeans_tc_supported.argtypes = [i64, i32]
    lib.cuml_b200_kmeans_estep_variant.argtypes = [vp, i64, i32]
    lib.cuml_b200_kmeans_fused_update.argtypes = [vp, i64, i32]
    _LIB = lib
    return lib


_STATUS_EXC = {1: ValueError, 2: CumlB200Error, 3: CumlB200Error, 4: CumlB200Error}


def check(status):
    """translate a C status into the exception the reference's `except +` would surface
    (python/cuml/cuml/cluster/cpp/kmeans.pxd:46: logic_error -> ValueError, others -> RuntimeError)."""
    if status != 0:
        msg = load().cuml_b200_last_error().decode("utf-8", "replace")
        raise _STATUS_EXC.get(status, CumlB200Error)(msg)


def default_params():
    p = KMeansParams()
    load().cuml_b200_kmeans_params_default(C.byref(p))
    return p


class Handle:
    """The sliver of pylibraft.common.handle.Handle the k-means path uses: a CUDA stream and
    (multi-GPU) an NCCL communicator injected by the caller."""

    def __init__(self, stream=None, n_ranks=1, rank=0):
        self._lib = load()
        self._h = C.c_void_p()
        check(self._lib.cuml_b200_handle_create(C.byref(self._h), C.c_void_p(stream or 0), None, rank, n_ranks))
        self.rank, self.n_ranks = rank, n_ranks
        self.comm_kind = None

    def getHandle(self):
        return self._h.value

    @property
    def ptr(self):
        return self._h

    def sync(self):
        check(self._lib.cuml_b200_handle_sync(self._h))

    def init_comm(self, unique_id: bytes, rank: int, n_ranks: int):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        check(self._lib.cuml_b200_handle_init_comm(self._h, buf, rank, n_ranks))
        self.rank, self.n_ranks = rank, n_ranks
        self.comm_kind = "nccl"

    def peer_window_create(self, n_ranks: int, slot_bytes: int = 0) -> bytes:
        """allocate this rank's exchange window of the peer-memory communicator; returns its 64-byte CUDA IPC handle"""
        buf = C.create_string_buffer(64)
        check(self._lib.cuml_b200_peer_window_create(self._h, n_ranks, slot_bytes, buf))
        return buf.raw

    def peer_window_attach(self, ipc_handles, rank: int, n_ranks: int):
        """map the windows of all ranks (`ipc_handles`: n_ranks handles in rank order)"""
        blob = b"".join(bytes(x) for x in ipc_handles)
        assert len(blob) == 64 * n_ranks
        buf = C.create_string_buffer(blob, len(blob))
        check(self._lib.cuml_b200_peer_window_attach(self._h, buf, rank, n_ranks))
        self.rank, self.n_ranks = rank, n_ranks
        self.comm_kind = "peer"

    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        check(load().cuml_b200_nccl_unique_id(buf))
        return buf.raw

    def close(self):
        if self._h:
            self._lib.cuml_b200_handle_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
