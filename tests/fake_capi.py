"""A CPU stand-in for the compute entry points of libcuml_b200 -- TEST INFRASTRUCTURE ONLY.

The product has no CPU path (tests/test_capi_cpu.py::test_no_cpu_fallback).  This module lets the CPU suite run
the *Python* side of the product -- the estimator, ``KMeansMG``, the distributed orchestration, the in-library
callers -- end to end without a GPU: it answers the same C-ABI calls (same argument order, same pointer / ``byref``
conventions as include/cuml_b200/kmeans_c.h) from the fp64 oracle, reading and writing the caller's buffers
through their raw addresses.  What it checks is therefore the glue: argument order, dtype / index-width selection,
buffer ownership, attribute plumbing, error paths.  The arithmetic of the real library is checked on the GPU
(tests/test_kmeans_gpu.py).

``install(monkeypatch)`` swaps the loaded library for the stand-in and keeps tensors on the CPU.
"""
from __future__ import annotations

import ctypes as C
import re

import numpy as np

_CT = {"f32": (C.c_float, np.float32), "f64": (C.c_double, np.float64)}
_IX = {"i32": np.int32, "i64": np.int64}


def _addr(p):
    if p is None:
        return None
    if isinstance(p, int):
        return p
    if hasattr(p, "value"):          # c_void_p
        return p.value
    raise TypeError(f"unexpected pointer argument {p!r}")


def _array(ptr, shape, ctype):
    """numpy view of caller memory at address ptr (no copy)"""
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape, dtype=np.dtype(ctype))
    buf = (ctype * n).from_address(_addr(ptr))
    return np.ctypeslib.as_array(buf).reshape(shape)


def _seed_centers(X, k, params, run, w):
    """seeded inits of the stand-in: sklearn's k-means++ / random rows (the real library has its own kernels)"""
    from sklearn.cluster import kmeans_plusplus
    seed = (int(params.rng_seed) + 7919 * run) % (2**32)
    if params.init == 1:                                   # Random
        idx = np.random.default_rng(seed).choice(X.shape[0], size=k, replace=False)
        return np.asarray(X[idx], dtype=np.float64)
    centers, _ = kmeans_plusplus(np.asarray(X, dtype=np.float64), k, random_state=seed, sample_weight=w)
    return centers


class FakeLib:
    """answers the calls of cuml_b200._lib / the estimators; everything else is forwarded to the real library
    (params_default, last_error, version work without a GPU)"""

    def __init__(self, real):
        self._real = real
        self.calls = []                                    # names of the compute entry points that were called

    # ---- handle ----------------------------------------------------------------------------------------
    def cuml_b200_handle_create(self, out, stream, comm, rank, n_ranks):
        out._obj.value = 0xB200
        return 0

    def cuml_b200_handle_destroy(self, h):
        return 0

    def cuml_b200_handle_sync(self, h):
        return 0

    def __getattr__(self, name):
        m = re.fullmatch(r"cuml_b200_kmeans_(fit|predict|transform)_(f32|f64)_(i32|i64)", name)
        if m:
            op, t, ix = m.groups()
            return lambda *a: self._dispatch(name, op, t, ix, *a)
        m = re.fullmatch(r"cuml_b200_kmeans_fit_parts_(f32|f64)", name)
        if m:
            return lambda *a: self._fit_parts(name, m.group(1), *a)
        m = re.fullmatch(r"cuml_b200_kmeans_fit_parts_labels_(f32|f64)", name)
        if m:
            return lambda *a: self._fit_parts(name, m.group(1), *a[:-1], labels_parts=a[-1])
        return getattr(self._real, name)

    # ---- compute ---------------------------------------------------------------------------------------
    def _dispatch(self, name, op, t, ix, *a):
        self.calls.append(name)
        return getattr(self, "_" + op)(t, ix, *a)

    def _lloyd(self, params, X, w, centers_out):
        from oracle import lloyd
        k = int(params.n_clusters)
        n = X.shape[0]
        if n < k:
            raise ValueError(f"n_samples={n} should be >= n_clusters={k}.")
        n_init = 1 if params.init == 2 else int(params.n_init)
        best = None
        for run in range(n_init):
            C0 = np.array(centers_out, dtype=np.float64) if params.init == 2 else _seed_centers(X, k, params, run, w)
            r = lloyd.fit(X, C0, max_iter=int(params.max_iter), tol=float(params.tol), sample_weight=w, rule="cuvs")
            if best is None or r["inertia"] < best["inertia"]:
                best = r
        centers_out[...] = best["centroids"].astype(centers_out.dtype)
        return best

    def _fit(self, t, ix, h, params, X, n, d, w, centers, inertia, n_iter):
        ct, _ = _CT[t]
        p = params._obj
        Xa = _array(X, (n, d), ct)
        wa = _array(w, (n,), ct) if _addr(w) else None
        Ca = _array(centers, (int(p.n_clusters), d), ct)
        r = self._lloyd(p, Xa, wa, Ca)
        inertia._obj.value = r["inertia"]
        n_iter._obj.value = r["n_iter"]
        return 0

    def _fit_parts(self, name, t, h, params, xp, rows, n_parts, d, wp, centers, inertia, n_iter, labels_parts=None):
        self.calls.append(name)
        ct, _ = _CT[t]
        p = params._obj
        parts = [_array(xp[i], (int(rows[i]), d), ct) for i in range(n_parts)]
        Xa = np.concatenate(parts) if parts else np.zeros((0, d))
        wa = None
        if wp is not None:
            wa = np.concatenate([_array(wp[i], (int(rows[i]),), ct) for i in range(n_parts)])
        Ca = _array(centers, (int(p.n_clusters), d), ct)
        r = self._lloyd(p, Xa, wa, Ca)
        inertia._obj.value = r["inertia"]
        n_iter._obj.value = r["n_iter"]
        if labels_parts is not None:
            off = 0
            for i in range(n_parts):
                if _addr(labels_parts[i]):
                    _array(labels_parts[i], (int(rows[i]),), C.c_int32)[...] = r["labels"][off:off + int(rows[i])]
                off += int(rows[i])
        return 0

    def _predict(self, t, ix, h, params, centers, X, n, d, w, normalize, labels, inertia):
        from oracle import lloyd
        ct, _ = _CT[t]
        p = params._obj
        Xa = _array(X, (n, d), ct)
        Ca = _array(centers, (int(p.n_clusters), d), ct)
        wa = _array(w, (n,), ct) if _addr(w) else None
        lab, inn = lloyd.predict(Xa, Ca, wa, normalize=bool(normalize))
        La = _array(labels, (n,), C.c_int32 if ix == "i32" else C.c_int64)
        La[...] = lab
        inertia._obj.value = inn
        return 0

    def _transform(self, t, ix, h, params, centers, X, n, d, out):
        from oracle import lloyd
        ct, _ = _CT[t]
        p = params._obj
        Xa = _array(X, (n, d), ct)
        Ca = _array(centers, (int(p.n_clusters), d), ct)
        Oa = _array(out, (n, int(p.n_clusters)), ct)
        Oa[...] = lloyd.transform(Xa, Ca, sqrt=(p.metric == 1))
        return 0


class FakeHandle:
    def __init__(self, stream=None, n_ranks=1, rank=0):
        self.ptr = C.c_void_p(0xB200)
        self.rank, self.n_ranks = rank, n_ranks

    def sync(self):
        pass

    def close(self):
        pass


def install(monkeypatch):
    """route the product's Python layer to the stand-in and keep its tensors on the CPU"""
    import torch
    from cuml_b200 import _lib
    from cuml_b200.cluster import kmeans, kmeans_mg
    fake = FakeLib(_lib.load())
    monkeypatch.setattr(_lib, "_LIB", fake)
    monkeypatch.setattr(_lib, "Handle", FakeHandle)
    monkeypatch.setattr(kmeans, "get_handle", lambda: FakeHandle())
    orig = kmeans._as_device_matrix

    def on_cpu(X, dtype=None, name="X", ndim=2, device=None):
        return orig(X, dtype=dtype, name=name, ndim=ndim, device=torch.device("cpu"))

    monkeypatch.setattr(kmeans, "_as_device_matrix", on_cpu)
    monkeypatch.setattr(kmeans_mg, "_as_device_matrix", on_cpu)
    return fake
