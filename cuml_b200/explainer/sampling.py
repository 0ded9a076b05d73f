"""``kmeans_sampling`` -- dataset summarisation by weighted k-means for the SHAP explainers, a caller of the
k-means path (reference python/cuml/cuml/explainer/sampling.py:14-79).

Same contract as the reference: missing values are imputed with the column mean, ``KMeans(n_clusters=k,
random_state=random_state, n_init="auto")`` is fitted, the ``k`` centres are the summary, and with
``round_values`` every coordinate of a centre is replaced by the nearest value that occurs in that column of
``X`` (first such row on ties, as ``argmin`` does).  torch stands in for cupy; the imputation and the rounding
are plain tensor code (they are not on the hot path), the fit is ``cuml_b200.cluster.KMeans``.
"""
from __future__ import annotations

import numpy as np


def _group_names(X):
    # reference sampling.py:43-50
    cols = getattr(X, "columns", None)
    if cols is not None:
        return [str(c) for c in cols]
    if hasattr(X, "name") and hasattr(X, "to_numpy"):     # pandas Series
        return [str(X.name)]
    shape = getattr(X, "shape", None)
    if shape is not None and len(shape) == 2:
        return [str(i) for i in range(shape[1])]
    return ["0"]


def impute_column_mean(X):
    """NaN -> mean of the non-missing entries of the column (the SimpleImputer(strategy="mean") step,
    reference sampling.py:59-62).  ``X``: 2-D floating torch tensor; returns a new tensor."""
    import torch
    miss = torch.isnan(X)
    if not bool(miss.any()):
        return X.clone()
    filled = torch.where(miss, torch.zeros_like(X), X)
    cnt = (~miss).sum(0).clamp(min=1).to(X.dtype)
    mean = filled.sum(0) / cnt
    return torch.where(miss, mean.expand_as(X), X)


def round_to_column_values(X, summary, chunk=1 << 22):
    """summary[i, j] <- X[argmin_r |X[r, j] - summary[i, j]|, j]   (reference sampling.py:68-73)."""
    import torch
    out = summary.clone()
    k, d = summary.shape
    n = X.shape[0]
    rows = max(1, chunk // max(1, k))
    for j in range(d):
        xj = X[:, j]
        best = torch.full((k,), float("inf"), dtype=X.dtype, device=X.device)
        val = out[:, j].clone()
        for s in range(0, n, rows):                       # [rows, k] blocks bound the temporary
            blk = xj[s:s + rows]
            dist = (blk[:, None] - summary[None, :, j]).abs()
            m, idx = dist.min(0)                          # first row on ties within the block
            take = m < best                               # strict: earlier blocks win ties
            best = torch.where(take, m, best)
            val = torch.where(take, blk[idx], val)
        out[:, j] = val
    return out


def kmeans_sampling(X, k, round_values=True, detailed=False, random_state=0, _estimator=None):
    """Summarise ``X`` (n_samples, n_features) by ``k`` weighted means.

    Returns ``summary`` (k, n_features), or ``(summary, group_names, labels)`` with ``detailed=True``.
    numpy in -> numpy out; torch / ``__cuda_array_interface__`` in -> torch CUDA tensors out.
    """
    import torch
    from ..cluster.kmeans import KMeans, _as_device_matrix
    group_names = _group_names(X)
    as_numpy = not (isinstance(X, torch.Tensor) or hasattr(X, "__cuda_array_interface__"))
    if hasattr(X, "to_numpy"):           # pandas DataFrame / Series
        X = X.to_numpy()
    if not isinstance(X, torch.Tensor) and not hasattr(X, "__cuda_array_interface__"):
        X = np.asarray(X)
        if X.ndim == 1:
            X = X.reshape(-1, 1)
    elif isinstance(X, torch.Tensor) and X.dim() == 1:
        X = X.reshape(-1, 1)
    Xd = _as_device_matrix(X).t
    Xd = impute_column_mean(Xd)
    est = (_estimator or KMeans)(n_clusters=k, random_state=random_state, n_init="auto", output_type="torch")
    est.fit(Xd)
    summary = est.cluster_centers_.clone()
    if round_values:
        summary = round_to_column_values(Xd, summary)
    if as_numpy:
        summary = summary.cpu().numpy()
    if detailed:
        labels = est.labels_
        return summary, group_names, (labels.cpu().numpy() if as_numpy else labels)
    return summary
