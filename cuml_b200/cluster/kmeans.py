"""``KMeans`` estimator -- the host-side mirror of ``cuml.cluster.KMeans``
(reference python/cuml/cuml/cluster/kmeans.pyx:437-1190) over the C-ABI.

Same constructor arguments, methods, fitted attributes and error strings as the reference;
torch is used only for device memory and streams (cupy is not in the image).
"""
from __future__ import annotations

import ctypes as C
import threading
from numbers import Integral

import numpy as np

from .. import _lib

_INT_MAX = 2**31 - 1
_tls = threading.local()


def _torch():
    import torch
    return torch


def get_handle():
    """thread-local handle on torch's current stream (reference internals/base.py:23-51)."""
    torch = _torch()
    if not torch.cuda.is_available():
        raise _lib.CumlB200Error("cuml_b200 needs a CUDA device: there is no CPU fallback")
    stream = torch.cuda.current_stream()
    key = (torch.cuda.current_device(), stream.cuda_stream)
    cache = getattr(_tls, "handles", None)
    if cache is None:
        cache = _tls.handles = {}
    h = cache.get(key)
    if h is None:
        # torch's default stream is the legacy NULL stream: pass cudaStreamLegacy (0x1) explicitly so the
        # library's work is ordered with torch's (a NULL argument would mean "handle-owned stream")
        h = cache[key] = _lib.Handle(stream=stream.cuda_stream or 1)
    return h


def _indices_i32(n_rows, n_cols):
    # reference kmeans.pyx:30-40
    return n_rows * n_cols <= _INT_MAX - 1


def check_random_seed(seed):
    # reference internals/validation.py:58-93
    if seed is None:
        return int(np.random.randint(0, 2**32 - 1))
    if isinstance(seed, (Integral, np.integer)):
        return int(seed) % (2**32)
    if isinstance(seed, np.random.RandomState):
        return int(seed.randint(0, 2**32 - 1))
    raise ValueError(f"{seed!r} cannot be used to seed the random number generator")


class _Input:
    """validated input: a C-contiguous 2-D float32/float64 torch CUDA tensor + how to hand results back"""

    def __init__(self, t, kind):
        self.t = t
        self.kind = kind  # "numpy" | "torch" | "cai"


def _as_device_matrix(X, dtype=None, name="X", ndim=2, device=None):
    torch = _torch()
    kind = "numpy"
    if isinstance(X, torch.Tensor):
        kind = "torch"
        t = X
    elif hasattr(X, "__cuda_array_interface__"):
        kind = "cai"
        t = torch.as_tensor(X, device="cuda")
    else:
        a = np.asarray(X)
        if a.dtype == object:
            raise ValueError(f"{name} must be numeric")
        t = torch.from_numpy(np.ascontiguousarray(a))
    if t.dim() != ndim:
        if ndim == 2:
            raise ValueError(f"Expected 2D array, got {t.dim()}D array instead")
        t = t.reshape(-1)
    if not t.dtype.is_floating_point or t.dtype in (torch.float16, torch.bfloat16):
        t = t.to(torch.float32)  # int inputs -> fp32 (xfail-list.yaml:693-699)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    t = t.to(dev).contiguous()
    return _Input(t, kind)


class KMeans:
    """KMeans (reference python/cuml/cuml/cluster/kmeans.pyx:437-1190).

    Parameters mirror the reference (kmeans.pyx:691-717).  ``tol`` follows the reference's GPU
    semantics: iteration stops when the raw squared centroid shift drops below ``tol``;
    ``tol=0`` runs all ``max_iter`` iterations.
    """

    _multi_gpu = False
    _cpu_class_path = "sklearn.cluster.KMeans"

    def __init__(self, *, n_clusters=8, max_iter=300, tol=1e-4, verbose=False, random_state=None,
                 init="scalable-k-means++", n_init="auto", oversampling_factor=2.0,
                 max_samples_per_batch=1 << 15, device_buffer_samples=0, init_size=0, output_type=None):
        self.n_clusters = n_clusters
        self.max_iter = max_iter
        self.tol = tol
        self.verbose = verbose
        self.random_state = random_state
        self.init = init
        self.n_init = n_init
        self.oversampling_factor = oversampling_factor
        self.max_samples_per_batch = max_samples_per_batch
        self.device_buffer_samples = device_buffer_samples
        self.init_size = init_size
        self.output_type = output_type

    # ---- sklearn plumbing -------------------------------------------------------------------
    @classmethod
    def _get_param_names(cls):
        return ["n_init", "oversampling_factor", "max_samples_per_batch", "device_buffer_samples", "init_size",
                "init", "max_iter", "n_clusters", "random_state", "tol", "verbose", "output_type"]

    def get_params(self, deep=True):
        return {k: getattr(self, k) for k in self._get_param_names()}

    def set_params(self, **params):
        for k, v in params.items():
            if k not in self._get_param_names():
                raise ValueError(f"Invalid parameter {k!r} for estimator {type(self).__name__}")
            setattr(self, k, v)
        return self

    @property
    def _n_features_out(self):
        return self.n_clusters

    # ---- parameter mapping (reference kmeans.pyx:43-90) --------------------------------------
    def _c_params(self):
        p = _lib.default_params()
        p.n_clusters = int(self.n_clusters)
        p.max_iter = int(self.max_iter)
        p.tol = float(self.tol)
        # reference internals/logger.pyx:30-37 (_verbose_to_level): True -> debug, False -> info, int v -> level 6 - v
        # in rapids_logger's numbering (trace 0, debug 1, info 2, warn 3, error 4, critical 5, off 6)
        if isinstance(self.verbose, bool):
            p.verbosity = 1 if self.verbose else 2
        else:
            p.verbosity = min(max(6 - int(self.verbose), 0), 6)
        p.metric = _lib.L2_EXPANDED
        p.batch_samples = int(self.max_samples_per_batch)
        p.init_size = int(self.init_size)
        p.device_buffer_samples = int(self.device_buffer_samples)
        p.oversampling_factor = float(self.oversampling_factor)
        if self._multi_gpu and self.random_state is None:
            raise ValueError(
                "KMeansMG requires `random_state != None`, please select a consistent "
                "non-None `random_state` to use across all partitions when calling "
                "KMeansMG")
        p.rng_seed = check_random_seed(self.random_state)
        if isinstance(self.init, str):
            if self.init == "k-means++":
                p.oversampling_factor = 0.0
                p.init = _lib.INIT_KMEANS_PLUS_PLUS
            elif self.init in ("scalable-k-means++", "k-means||"):
                p.init = _lib.INIT_KMEANS_PLUS_PLUS
            elif self.init == "random":
                p.init = _lib.INIT_RANDOM
            else:
                raise ValueError(f"init={self.init!r} is not supported")
        else:
            p.init = _lib.INIT_ARRAY
        if self._multi_gpu and p.oversampling_factor == 0:
            raise ValueError("init='k-means++' or oversampling_factor=0 not supported for KMeansMG")
        if self.n_init == "auto":
            p.n_init = 1 if (isinstance(self.init, str) and p.init == _lib.INIT_KMEANS_PLUS_PLUS) else 10
        else:
            p.n_init = int(self.n_init)
        return p

    def _validate_fit_params(self):
        if not isinstance(self.n_clusters, Integral) or self.n_clusters <= 0:
            raise ValueError(f"n_clusters={self.n_clusters} should be a positive integer.")
        if int(self.device_buffer_samples) < 0:
            raise ValueError(f"device_buffer_samples must be >= 0, got {int(self.device_buffer_samples)}.")

    def _validate_fit_row_constraints(self, n_rows):
        if not self._multi_gpu and n_rows < self.n_clusters:
            raise ValueError(f"n_samples={n_rows} should be >= n_clusters={self.n_clusters}.")

    def _check_is_fitted(self):
        if not hasattr(self, "_centers"):
            raise RuntimeError("This KMeans instance is not fitted yet. Call 'fit' first.")

    # ---- output conversion --------------------------------------------------------------------
    def _out(self, t, kind):
        ot = self.output_type
        if ot in (None, "input"):
            ot = "numpy" if kind == "numpy" else "torch"
        if ot == "numpy":
            return t.detach().cpu().numpy()
        return t

    @property
    def cluster_centers_(self):
        self._check_is_fitted()
        return self._out(self._centers, self._in_kind)

    @property
    def labels_(self):
        self._check_is_fitted()
        return self._out(self._labels, self._in_kind)

    # ---- fit ------------------------------------------------------------------------------------
    def _prepare_centers(self, X):
        torch = _torch()
        k, d = int(self.n_clusters), X.shape[1]
        if isinstance(self.init, str):
            return torch.zeros((k, d), dtype=X.dtype, device=X.device)
        centers = _as_device_matrix(self.init, dtype=X.dtype, name="init", device=X.device).t.clone()
        if centers.shape[0] != k:
            raise ValueError(f"The shape of the initial centers {tuple(centers.shape)} does not "
                             f"match the number of clusters {k}.")
        if centers.shape[1] != d:
            raise ValueError(f"The shape of the initial centers {tuple(centers.shape)} does not "
                             f"match the number of features of the data {d}.")
        return centers

    def fit(self, X, y=None, sample_weight=None):
        """Compute k-means clustering with X (reference kmeans.pyx:725-821)."""
        torch = _torch()
        if int(self.device_buffer_samples) > 0 and self._multi_gpu:
            raise ValueError(f"device_buffer_samples={int(self.device_buffer_samples)} is not supported for the "
                             f"multi-GPU KMeans fit path; set device_buffer_samples=0.")
        self._validate_fit_params()
        if self._streams_from_host(X):
            return self._fit_out_of_core(X, sample_weight)
        xin = _as_device_matrix(X)
        Xd = xin.t
        if Xd.shape[0] < 1 or Xd.shape[1] < 1:
            raise ValueError(f"Found array with {Xd.shape[0]} sample(s) and {Xd.shape[1]} feature(s) while a "
                             f"minimum of 1 is required.")
        if not bool(torch.isfinite(Xd).all()):
            raise ValueError("Input X contains NaN or infinity.")
        wd = None
        if sample_weight is not None:
            wd = _as_device_matrix(sample_weight, dtype=Xd.dtype, name="sample_weight", ndim=1, device=Xd.device).t
            if wd.shape[0] != Xd.shape[0]:
                raise ValueError("sample_weight.shape == {}, expected {}!".format(tuple(wd.shape), (Xd.shape[0],)))
        n_rows, n_cols = Xd.shape
        self.n_features_in_ = n_cols
        self._validate_fit_row_constraints(n_rows)
        centers = self._prepare_centers(Xd)
        handle = self.handle if self._multi_gpu else get_handle()
        params = self._c_params()
        # one library call: fit + the labels / inertia of its own final assignment pass (the reference runs a second,
        # redundant E-step here, kmeans.pyx:803-812)
        n_iter, labels, inertia = self._c_fit_labels(handle, params, Xd.data_ptr(), n_rows, n_cols,
                                                     wd.data_ptr() if wd is not None else None, centers)
        handle.sync()
        self._centers = centers
        self._labels = labels
        self._in_kind = xin.kind
        self.inertia_ = inertia
        self.n_iter_ = n_iter
        return self

    # ---- out-of-core fit: host X streamed through a device buffer of device_buffer_samples rows -------------
    def _streams_from_host(self, X):
        buf = int(self.device_buffer_samples)
        return (not self._multi_gpu and buf > 0 and isinstance(X, np.ndarray) and X.ndim == 2
                and X.shape[0] > buf and X.dtype in (np.float32, np.float64))

    def _fit_out_of_core(self, X, sample_weight):
        """Host-resident X larger than ``device_buffer_samples``: the library walks it batch by batch every
        iteration (reference host-data path, kmeans_fit.cu:167-231); X is never resident on the device."""
        torch = _torch()
        lib = _lib.load()
        Xh = np.ascontiguousarray(X)
        n_rows, n_cols = Xh.shape
        if not np.isfinite(Xh).all():
            raise ValueError("Input X contains NaN or infinity.")
        wh = None
        if sample_weight is not None:
            wh = np.ascontiguousarray(np.asarray(sample_weight, dtype=Xh.dtype).reshape(-1))
            if wh.shape[0] != n_rows:
                raise ValueError("sample_weight.shape == {}, expected {}!".format(wh.shape, (n_rows,)))
        self.n_features_in_ = n_cols
        self._validate_fit_row_constraints(n_rows)
        tdt = torch.float32 if Xh.dtype == np.float32 else torch.float64
        probe = torch.empty((0, n_cols), dtype=tdt, device=torch.device("cuda", torch.cuda.current_device()))
        centers = self._prepare_centers(probe)
        handle = get_handle()
        params = self._c_params()
        f32 = Xh.dtype == np.float32
        torch.cuda.synchronize()
        # the library streams the host rows itself and hands back the labels of its final pass
        n_iter, labels, inertia = self._c_fit_labels(handle, params, Xh.ctypes.data, n_rows, n_cols,
                                                     wh.ctypes.data if wh is not None else None, centers)
        handle.sync()
        self._centers = centers
        self._labels = labels
        self._in_kind = "numpy"
        self.inertia_ = inertia
        self.n_iter_ = n_iter
        return self

    def _c_fit_labels(self, handle, params, x_ptr, n, d, w_ptr, centers):
        """cuml_b200_kmeans_fit_parts_labels_*: (n_iter, labels [n] on the centres' device, inertia).  x_ptr / w_ptr may
        be host or device addresses (the library probes the residency like ML::is_device_or_managed_type)."""
        torch = _torch()
        lib = _lib.load()
        f32 = centers.dtype == torch.float32
        k = centers.shape[0]
        labels = torch.zeros(n, dtype=torch.int32, device=centers.device)
        fn = lib.cuml_b200_kmeans_fit_parts_labels_f32 if f32 else lib.cuml_b200_kmeans_fit_parts_labels_f64
        inertia = (C.c_float if f32 else C.c_double)()
        n_iter = C.c_int64()
        xp = (C.c_void_p * 1)(x_ptr)
        rows = (C.c_int64 * 1)(n)
        wp = (C.c_void_p * 1)(w_ptr) if w_ptr is not None else None
        lp = (C.c_void_p * 1)(labels.data_ptr())
        _lib.check(fn(handle.ptr, C.byref(params), xp, rows, 1, d, wp, centers.data_ptr(), C.byref(inertia),
                      C.byref(n_iter), lp))
        if not (_indices_i32(n, d) and _indices_i32(k, d)):     # reference kmeans.pyx:277-281: int64 labels then
            labels = labels.to(torch.int64)
        return int(n_iter.value), labels, float(inertia.value)

    def _c_fit(self, handle, params, X, w, centers):
        lib = _lib.load()
        f32 = X.dtype == _torch().float32
        n, d = X.shape
        i32 = _indices_i32(n, d)
        fn = getattr(lib, "cuml_b200_kmeans_fit_%s_%s" % ("f32" if f32 else "f64", "i32" if i32 else "i64"))
        inertia = (C.c_float if f32 else C.c_double)()
        n_iter = (C.c_int32 if i32 else C.c_int64)()
        _lib.check(fn(handle.ptr, C.byref(params), X.data_ptr(), n, d, w.data_ptr() if w is not None else None,
                      centers.data_ptr(), C.byref(inertia), C.byref(n_iter)))
        return int(n_iter.value)

    def _c_predict(self, handle, params, X, w, centers, normalize_weights=True):
        torch = _torch()
        lib = _lib.load()
        f32 = X.dtype == torch.float32
        n, d = X.shape
        k = centers.shape[0]
        i32 = _indices_i32(n, d) and _indices_i32(k, d)  # reference kmeans.pyx:277-281
        labels = torch.zeros(n, dtype=torch.int32 if i32 else torch.int64, device=X.device)
        fn = getattr(lib, "cuml_b200_kmeans_predict_%s_%s" % ("f32" if f32 else "f64", "i32" if i32 else "i64"))
        inertia = (C.c_float if f32 else C.c_double)()
        _lib.check(fn(handle.ptr, C.byref(params), centers.data_ptr(), X.data_ptr(), n, d,
                      w.data_ptr() if w is not None else None, 1 if normalize_weights else 0, labels.data_ptr(),
                      C.byref(inertia)))
        return labels, float(inertia.value)

    # ---- inference --------------------------------------------------------------------------------
    def fit_predict(self, X, y=None, sample_weight=None):
        return self.fit(X, sample_weight=sample_weight).labels_

    def _predict_host_chunked(self, X, sample_weight=None):
        """Host X larger than ``device_buffer_samples``: predict chunk by chunk (reference
        ``_kmeans_predict_host_chunked``, kmeans.pyx:356-434).  Weights are normalised once over the whole input
        (sum(w) == n_samples), each chunk is then predicted with ``normalize_weights=False`` and the chunk
        inertias add up."""
        torch = _torch()
        Xh = np.ascontiguousarray(X)
        n_rows, n_cols = Xh.shape
        if n_cols != self._centers.shape[1]:
            raise ValueError(f"X has {n_cols} features, but KMeans is expecting "
                             f"{self._centers.shape[1]} features as input.")
        dev, tdt = self._centers.device, self._centers.dtype
        wh = None
        if sample_weight is not None:
            wh = np.asarray(sample_weight, dtype=np.float64).reshape(-1)
            if wh.shape[0] != n_rows:
                raise ValueError("sample_weight.shape == {}, expected {}!".format(wh.shape, (n_rows,)))
            wh = wh * (n_rows / wh.sum())
        handle = get_handle()
        params = self._c_params()
        buf = int(self.device_buffer_samples)
        labels = torch.empty(n_rows, dtype=torch.int32, device=dev)
        inertia = 0.0
        for s0 in range(0, n_rows, buf):
            xb = torch.from_numpy(Xh[s0:s0 + buf]).to(device=dev, dtype=tdt)
            wb = None
            if wh is not None:
                wb = torch.from_numpy(np.ascontiguousarray(wh[s0:s0 + buf])).to(device=dev, dtype=tdt)
            lb, ib = self._c_predict(handle, params, xb, wb, self._centers, normalize_weights=False)
            labels[s0:s0 + xb.shape[0]] = lb.to(torch.int32)
            inertia += ib
        handle.sync()
        return self._out(labels, "numpy"), inertia

    def _predict_labels_inertia(self, X, sample_weight=None):
        torch = _torch()
        self._check_is_fitted()
        if self._streams_from_host(X):
            return self._predict_host_chunked(X, sample_weight)
        xin = _as_device_matrix(X, dtype=self._centers.dtype, device=self._centers.device)
        if xin.t.shape[1] != self._centers.shape[1]:
            raise ValueError(f"X has {xin.t.shape[1]} features, but KMeans is expecting "
                             f"{self._centers.shape[1]} features as input.")
        if sample_weight is None:
            # the reference materialises ones (kmeans.pyx:1038-1039); a null pointer means the same here
            wd = None
        else:
            wd = _as_device_matrix(sample_weight, dtype=xin.t.dtype, name="sample_weight", ndim=1,
                                   device=xin.t.device).t
        handle = get_handle()
        params = self._c_params()
        labels, inertia = self._c_predict(handle, params, xin.t, wd, self._centers)
        handle.sync()
        return self._out(labels, xin.kind), inertia

    def predict(self, X):
        return self._predict_labels_inertia(X)[0]

    def transform(self, X):
        """distances to the cluster centres; L2Expanded => squared (reference kmeans.pyx:51,1074-1162)."""
        torch = _torch()
        self._check_is_fitted()
        xin = _as_device_matrix(X, dtype=self._centers.dtype, device=self._centers.device)
        n, d = xin.t.shape
        if d != self._centers.shape[1]:
            raise ValueError(f"X has {d} features, but KMeans is expecting "
                             f"{self._centers.shape[1]} features as input.")
        k = int(self._centers.shape[0])
        if not _indices_i32(n, k):
            raise NotImplementedError("KMeans.transform does not currently support output shapes "
                                      f"that require int64 indexing. Got output shape ({n}, {k}).")
        out = torch.zeros((n, k), dtype=xin.t.dtype, device=xin.t.device)
        lib = _lib.load()
        f32 = xin.t.dtype == torch.float32
        i32 = _indices_i32(n, d)
        fn = getattr(lib, "cuml_b200_kmeans_transform_%s_%s" % ("f32" if f32 else "f64", "i32" if i32 else "i64"))
        handle = get_handle()
        params = self._c_params()
        params.n_clusters = k
        _lib.check(fn(handle.ptr, C.byref(params), self._centers.data_ptr(), xin.t.data_ptr(), n, d, out.data_ptr()))
        handle.sync()
        return self._out(out, xin.kind)

    def score(self, X, y=None, sample_weight=None):
        return -1 * self._predict_labels_inertia(X, sample_weight=sample_weight)[1]

    def fit_transform(self, X, y=None, sample_weight=None):
        self.fit(X, sample_weight=sample_weight)
        return self.transform(X)

    # ---- sklearn interop (reference kmeans.pyx:604-689, internals/interop.py:206-257) ----------
    def as_sklearn(self):
        from sklearn.cluster._kmeans import KMeans as SkKMeans   # the real class even when accel.install() swapped the public name
        init = self.init
        if not isinstance(init, str):
            init = _as_device_matrix(init).t.cpu().numpy()
        elif init in ("scalable-k-means++", "k-means||"):
            init = "k-means++"
        sk = SkKMeans(n_clusters=self.n_clusters, init=init, n_init=self.n_init, max_iter=self.max_iter,
                      tol=self.tol, random_state=self.random_state)
        if hasattr(self, "_centers"):
            sk.cluster_centers_ = self._centers.cpu().numpy()
            sk.labels_ = self._labels.cpu().numpy()
            sk.inertia_ = self.inertia_
            sk.n_iter_ = self.n_iter_
            sk.n_features_in_ = self.n_features_in_
            sk._n_features_out = self.n_clusters
            try:
                from sklearn.utils._openmp_helpers import _openmp_effective_n_threads
                sk._n_threads = _openmp_effective_n_threads()
            except ImportError:
                sk._n_threads = 1
        return sk

    @classmethod
    def from_sklearn(cls, model):
        if callable(model.init):
            raise ValueError(f"`init={model.init!r}` is not supported")
        init = model.init
        if isinstance(init, str):
            init = {"k-means++": "scalable-k-means++", "random": "random"}[init]
        est = cls(n_clusters=model.n_clusters, init=init, n_init=model.n_init, max_iter=model.max_iter,
                  tol=model.tol, random_state=model.random_state)
        if hasattr(model, "cluster_centers_"):
            torch = _torch()
            est._centers = torch.as_tensor(np.ascontiguousarray(model.cluster_centers_)).cuda()
            est._labels = torch.as_tensor(model.labels_).cuda()
            est._in_kind = "numpy"
            est.inertia_ = model.inertia_
            est.n_iter_ = model.n_iter_
            est.n_features_in_ = model.n_features_in_
        return est

    def __getstate__(self):
        st = dict(self.__dict__)
        for k in ("_centers", "_labels"):
            if k in st:
                st[k] = st[k].cpu()
        st.pop("handle", None)
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)
        torch = _torch()
        if torch.cuda.is_available():
            for k in ("_centers", "_labels"):
                if k in self.__dict__:
                    self.__dict__[k] = self.__dict__[k].cuda()
