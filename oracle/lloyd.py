"""fp64 numpy restatement of the k-means path.  Test infrastructure only.

Rules restated (each with the reference evidence it follows):

E-step     label_i = argmin_j (||c_j||^2 - 2 x_i.c_j), ties -> lowest index (first-min, strict <)
           -- sklearn/cluster/_k_means_lloyd.pyx:187-213 (the reference CPU path,
           python/cuml/cuml/cluster/kmeans.pyx:604); GPU path: cuvs fusedL2NN role called from
           cpp/src/kmeans/kmeans_predict.cu:41-42.
M-step     c_j = sum_i w_i x_i / sum_i w_i over members -- _k_means_lloyd.pyx:215-218,
           _k_means_common.pyx:274-296.
empty      rule="cuvs": an empty cluster keeps its previous centroid (no relocation; the
           deviation the reference documents in cuml_accel_tests/upstream/scikit-learn/
           xfail-list.yaml:679-682).  rule="sklearn": relocation is NOT restated here (use
           oracle.sklearn_ref for that behaviour).
stopping   rule="cuvs": stop after the iteration whose raw squared centroid shift
           sum||c_new-c_old||^2 < tol (xfail-list.yaml:707-715); n_iter = iterations executed.
           rule="sklearn": stop when labels repeat or shift <= tol*mean(var(X))
           (sklearn/cluster/_kmeans.py:285-293,715-730).
weights    normalised so sum w = n_samples unless normalize_weights=False
           (python/cuml/cuml/cluster/kmeans.pyx:371-378; xfail-list.yaml:683-692).
inertia    sum_i w_i ||x_i - c_label(i)||^2 with the FINAL centroids, exact difference form
           (_k_means_common.pyx:94-124).
transform  squared distances for L2Expanded, sqrt for L2SqrtExpanded
           (cpp/include/cuml/common/distance_type.hpp:13-36; kmeans.pyx:51).
"""
from __future__ import annotations

import numpy as np

_CHUNK = 1 << 16


def normalize_weights(w, n):
    """scale w so that sum(w) == n (cuVS checkWeight role; kmeans.pyx:371-378)."""
    w = np.asarray(w, dtype=np.float64)
    return w * (float(n) / w.sum())


def e_step(X, C, return_second=False):
    """labels (int64), min squared distance (fp64, exact difference form), optionally the
    second-smallest squared distance (for top-2 gap checks)."""
    C64 = np.asarray(C, dtype=np.float64)
    cn = (C64 * C64).sum(1)
    n = X.shape[0]
    labels = np.empty(n, dtype=np.int64)
    dmin = np.empty(n, dtype=np.float64)
    d2nd = np.empty(n, dtype=np.float64) if return_second else None
    for s in range(0, n, _CHUNK):
        x = np.asarray(X[s:s + _CHUNK], dtype=np.float64)
        part = cn[None, :] - 2.0 * (x @ C64.T)          # ||c||^2 - 2 x.c
        lab = part.argmin(1)                            # numpy argmin = first minimum
        labels[s:s + _CHUNK] = lab
        diff = x - C64[lab]
        dmin[s:s + _CHUNK] = (diff * diff).sum(1)
        if return_second:
            xn = (x * x).sum(1)
            full = part + xn[:, None]
            if C64.shape[0] > 1:
                full[np.arange(len(lab)), lab] = np.inf
                d2nd[s:s + _CHUNK] = np.maximum(full.min(1), 0.0)
            else:
                d2nd[s:s + _CHUNK] = np.inf
    if return_second:
        return labels, dmin, d2nd
    return labels, dmin


def m_step(X, labels, k, w=None, C_old=None):
    """weighted sums [k,d], weights [k]; new centroids with the cuVS empty-cluster rule."""
    n, d = X.shape
    S = np.zeros((k, d), dtype=np.float64)
    W = np.zeros(k, dtype=np.float64)
    for s in range(0, n, _CHUNK):
        x = np.asarray(X[s:s + _CHUNK], dtype=np.float64)
        lab = labels[s:s + _CHUNK]
        ww = np.ones(len(lab)) if w is None else np.asarray(w[s:s + _CHUNK], dtype=np.float64)
        np.add.at(W, lab, ww)
        # sort-free scatter-add per feature block
        order = np.argsort(lab, kind="stable")
        ls = lab[order]
        xs = x[order] * ww[order, None]
        bounds = np.flatnonzero(np.diff(ls)) + 1
        starts = np.concatenate(([0], bounds))
        S[ls[starts]] += np.add.reduceat(xs, starts, axis=0)
    C_new = np.array(C_old, dtype=np.float64, copy=True) if C_old is not None else np.zeros((k, d))
    nz = W > 0
    C_new[nz] = S[nz] / W[nz, None]
    return S, W, C_new


def lloyd_step(X, C, w=None):
    """one full Lloyd iteration: (labels, S, W, C_new, inertia wrt C, shift2)."""
    labels, dmin = e_step(X, C)
    S, W, C_new = m_step(X, labels, C.shape[0], w, C_old=C)
    inertia = float(dmin.sum() if w is None else (dmin * np.asarray(w, dtype=np.float64)).sum())
    shift2 = float(((C_new - np.asarray(C, dtype=np.float64)) ** 2).sum())
    return labels, S, W, C_new, inertia, shift2


def fit(X, C0, max_iter=300, tol=1e-4, sample_weight=None, rule="cuvs", normalize=True):
    """Lloyd fit from init='array' centroids.  Returns dict(centroids, labels, inertia, n_iter)."""
    n = X.shape[0]
    w = None
    if sample_weight is not None:
        w = normalize_weights(sample_weight, n) if (normalize and rule == "cuvs") else np.asarray(sample_weight, np.float64)
    C = np.asarray(C0, dtype=np.float64).copy()
    tol_eff = tol
    if rule == "sklearn":
        tol_eff = tol * float(np.mean(np.var(np.asarray(X, dtype=np.float64), axis=0)))
    prev_labels = None
    n_iter = 0
    for it in range(1, max_iter + 1):
        labels, S, W, C_new, _, shift2 = lloyd_step(X, C, w)
        n_iter = it
        if rule == "sklearn":
            if prev_labels is not None and np.array_equal(labels, prev_labels):
                break                                   # strict convergence; centres NOT updated
            C = C_new
            prev_labels = labels
            if shift2 <= tol_eff:
                break
        else:
            C = C_new
            if shift2 < tol_eff:
                break
    labels, dmin = e_step(X, C)
    inertia = float(dmin.sum() if w is None else (dmin * w).sum())
    return dict(centroids=C, labels=labels, inertia=inertia, n_iter=n_iter)


def predict(X, C, sample_weight=None, normalize=True):
    """labels + weighted inertia (ML::kmeans::predict, cpp/include/cuml/cluster/kmeans.hpp:154-195)."""
    labels, dmin = e_step(X, C)
    if sample_weight is None:
        return labels, float(dmin.sum())
    w = normalize_weights(sample_weight, X.shape[0]) if normalize else np.asarray(sample_weight, np.float64)
    return labels, float((dmin * w).sum())


def transform(X, C, sqrt=False):
    """[n,k] distances (ML::kmeans::transform, kmeans.hpp:213-242); squared unless sqrt."""
    X64 = np.asarray(X, dtype=np.float64)
    C64 = np.asarray(C, dtype=np.float64)
    out = np.empty((X64.shape[0], C64.shape[0]))
    for s in range(0, X64.shape[0], _CHUNK):
        x = X64[s:s + _CHUNK]
        diff = x[:, None, :] - C64[None, :, :] if x.shape[0] * C64.size < (1 << 24) else None
        if diff is not None:
            out[s:s + _CHUNK] = (diff * diff).sum(2)
        else:
            out[s:s + _CHUNK] = np.maximum(
                (x * x).sum(1)[:, None] + (C64 * C64).sum(1)[None, :] - 2.0 * x @ C64.T, 0.0)
    return np.sqrt(out) if sqrt else out


def inertia_of(X, C, labels, w=None):
    """exact fp64 sum_i w_i ||x_i - c_label(i)||^2 for given labels."""
    C64 = np.asarray(C, dtype=np.float64)
    tot = 0.0
    for s in range(0, X.shape[0], _CHUNK):
        x = np.asarray(X[s:s + _CHUNK], dtype=np.float64)
        diff = x - C64[labels[s:s + _CHUNK]]
        dd = (diff * diff).sum(1)
        tot += float(dd.sum() if w is None else (dd * np.asarray(w[s:s + _CHUNK], np.float64)).sum())
    return tot


def label_disagreements_ok(X, C, labels_test, rel_gap_tol):
    """Compare labels with the fp64 E-step.  A disagreement is *excusable* iff the test label's
    exact distance is within rel_gap_tol * (||x||^2 + ||c||^2) of the true minimum (i.e. the
    top-2 gap is below the stated fp32 tolerance).  Returns (agreement_fraction, n_bad)."""
    labels, dmin = e_step(X, C)
    labels_test = np.asarray(labels_test).astype(np.int64)
    bad = np.flatnonzero(labels != labels_test)
    n_bad = 0
    C64 = np.asarray(C, dtype=np.float64)
    for i in bad:
        x = np.asarray(X[i], dtype=np.float64)
        dt = ((x - C64[labels_test[i]]) ** 2).sum()
        scale = (x * x).sum() + (C64[labels_test[i]] ** 2).sum()
        if dt - dmin[i] > rel_gap_tol * scale:
            n_bad += 1
    return 1.0 - len(bad) / max(1, len(labels)), n_bad
