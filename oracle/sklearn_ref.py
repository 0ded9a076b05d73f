"""The reference's CPU execution path, run as-is.  Test infrastructure / timed CPU baseline only.

``cuml.cluster.KMeans`` declares ``_cpu_class_path = "sklearn.cluster.KMeans"``
(python/cuml/cuml/cluster/kmeans.pyx:604) and maps parameters with ``_params_to_cpu``
(kmeans.pyx:659-672): whenever the GPU path is unavailable the reference executes exactly
this class.  scikit-learn ships in the image (build container and GPU box alike), so the
reference CPU path is importable on both sides.
"""
from __future__ import annotations

import os
import time
import warnings

import numpy as np


def n_threads():
    try:
        from sklearn.utils._openmp_helpers import _openmp_effective_n_threads
        return int(_openmp_effective_n_threads())
    except Exception:  # pragma: no cover
        return os.cpu_count() or 1


def fit(X, init, max_iter=50, tol=0.0, sample_weight=None, n_init=1, copy_x=True):
    """sklearn KMeans(init=array, algorithm='lloyd'); returns dict like oracle.lloyd.fit.
    copy_x=False lets sklearn centre X in place (no second copy of a matrix that fills a quarter of host memory)."""
    from sklearn.cluster import KMeans
    km = KMeans(n_clusters=init.shape[0], init=np.asarray(init), n_init=n_init, max_iter=max_iter,
                tol=tol, algorithm="lloyd", copy_x=copy_x)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        t0 = time.perf_counter()
        km.fit(X, sample_weight=sample_weight)
        dt = time.perf_counter() - t0
    return dict(centroids=km.cluster_centers_, labels=km.labels_.astype(np.int64),
                inertia=float(km.inertia_), n_iter=int(km.n_iter_), seconds=dt, model=km)


def time_fit(X, init, max_iter, reps=3, copy_x=True):
    """best-of-reps wall time (the reference harness convention,
    python/cuml/cuml/benchmark/runners.py:42-64) -> (seconds, n_iter)."""
    best, n_iter = None, 0
    for _ in range(reps):
        r = fit(X, init, max_iter=max_iter, tol=0.0, copy_x=copy_x)
        if best is None or r["seconds"] < best:
            best, n_iter = r["seconds"], r["n_iter"]
    return best, n_iter


def time_predict(X, centers, reps=3):
    """best-of-reps wall time of the CPU path's ``predict`` (labels of X for fixed centres) -> seconds."""
    from sklearn.cluster import KMeans
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        km = KMeans(n_clusters=centers.shape[0], init=np.asarray(centers), n_init=1, max_iter=1, tol=0.0,
                    algorithm="lloyd").fit(X[:max(centers.shape[0], min(len(X), 4 * centers.shape[0]))])
        km.cluster_centers_ = np.ascontiguousarray(centers, dtype=X.dtype)
        best = None
        for _ in range(reps):
            t0 = time.perf_counter()
            km.predict(X)
            dt = time.perf_counter() - t0
            best = dt if best is None or dt < best else best
    return best
