"""Bin edges of ``KBinsDiscretizer(strategy='kmeans')`` -- the one place the reference's preprocessing code
calls the k-means path (reference python/cuml/cuml/_thirdparty/sklearn/preprocessing/_discretization.py:190-230).

Only that call is mirrored (the discretizer itself is outside SURVEY.md section 8): per feature a 1-D k-means
from a deterministic, uniformly spaced init; sorted centres; edges at the midpoints between neighbouring centres,
closed by the column minimum and maximum; bins narrower than 1e-8 removed.
"""
from __future__ import annotations

import warnings

import numpy as np


def uniform_init(col_min, col_max, n_bins):
    """centres of ``n_bins`` equal-width bins over [col_min, col_max]  (_discretization.py:207-209)"""
    edges = np.linspace(col_min, col_max, n_bins + 1)
    return ((edges[1:] + edges[:-1]) * 0.5)[:, None]


def edges_from_centers(centers, col_min, col_max, n_bins, feature=0):
    """sorted centres -> bin edges, small bins removed  (_discretization.py:216-230)"""
    centers = np.sort(np.asarray(centers, dtype=np.float64).reshape(-1))
    edges = (centers[1:] + centers[:-1]) * 0.5
    edges = np.r_[col_min, edges, col_max]
    mask = np.diff(edges, prepend=-np.inf) > 1e-8
    edges = edges[mask]
    if len(edges) - 1 != n_bins:
        warnings.warn("Bins whose width are too small (i.e., <= 1e-8) in feature %d are removed. Consider "
                      "decreasing the number of bins." % feature)
    return edges


def kmeans_bin_edges(X, n_bins, _estimator=None):
    """Per-feature bin edges of the 'kmeans' strategy.  ``X``: (n_samples, n_features) array-like; ``n_bins``: int
    or one int per feature.  Returns a list of 1-D float64 numpy arrays (edges of feature j)."""
    import torch
    from ..cluster.kmeans import KMeans, _as_device_matrix
    Xd = _as_device_matrix(X).t
    n, d = Xd.shape
    bins = np.broadcast_to(np.asarray(n_bins, dtype=np.int64), (d,)).copy()
    out = []
    for jj in range(d):
        column = Xd[:, jj].contiguous()
        col_min, col_max = float(column.min()), float(column.max())
        if col_min == col_max:
            warnings.warn("Feature %d is constant and will be replaced with 0." % jj)
            out.append(np.array([-np.inf, np.inf]))
            continue
        init = uniform_init(col_min, col_max, int(bins[jj])).astype(np.float32 if Xd.dtype == torch.float32 else np.float64)
        km = (_estimator or KMeans)(n_clusters=int(bins[jj]), init=init, n_init=1, output_type="numpy")
        km.fit(column[:, None])
        out.append(edges_from_centers(km.cluster_centers_[:, 0], col_min, col_max, int(bins[jj]), jj))
    return out
