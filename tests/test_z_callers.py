"""Callers of the k-means path: ``kmeans_sampling`` (reference explainer/sampling.py:14-79) and
the bin edges of ``KBinsDiscretizer(strategy='kmeans')`` (reference _discretization.py:190-230).

CPU part: the tensor code around the fit (imputation, rounding, edge construction) against scikit-learn / a
loop restatement of the reference.  GPU part: the composed functions with the fit on the device."""
import warnings

import numpy as np
import pytest


def test_impute_column_mean_matches_simple_imputer():
    import torch
    from sklearn.impute import SimpleImputer
    from cuml_b200.explainer.sampling import impute_column_mean
    rng = np.random.default_rng(0)
    X = rng.normal(size=(200, 6))
    X[rng.random(X.shape) < 0.1] = np.nan
    got = impute_column_mean(torch.from_numpy(X)).numpy()
    want = SimpleImputer(missing_values=np.nan, strategy="mean").fit_transform(X)
    np.testing.assert_allclose(got, want, rtol=1e-12)
    clean = rng.normal(size=(10, 3))
    assert np.array_equal(impute_column_mean(torch.from_numpy(clean)).numpy(), clean)


def test_round_to_column_values_matches_reference_loops():
    import torch
    from cuml_b200.explainer.sampling import round_to_column_values
    rng = np.random.default_rng(1)
    X = rng.integers(0, 12, size=(500, 4)).astype(np.float64)      # discrete values: many exact ties
    summary = rng.uniform(0, 11, size=(7, 4))
    want = summary.copy()
    for i in range(summary.shape[0]):                               # reference sampling.py:68-73
        for j in range(X.shape[1]):
            ind = np.argmin(np.abs(X[:, j] - summary[i, j]))
            want[i, j] = X[ind, j]
    for chunk in (1 << 22, 64):                                     # one block / many blocks
        got = round_to_column_values(torch.from_numpy(X), torch.from_numpy(summary), chunk=chunk).numpy()
        assert np.array_equal(got, want)


def test_group_names():
    import pandas as pd
    from cuml_b200.explainer.sampling import _group_names
    assert _group_names(pd.DataFrame({"a": [1.0], "b": [2.0]})) == ["a", "b"]
    assert _group_names(pd.Series([1.0, 2.0], name="s")) == ["s"]
    assert _group_names(np.zeros((3, 2))) == ["0", "1"]
    assert _group_names(np.zeros(3)) == ["0"]


def test_kmeans_edges_match_sklearn_discretizer():
    """uniform init + edge construction == sklearn's KBinsDiscretizer when fed the same 1-D fit"""
    from sklearn.cluster import KMeans as SkKMeans
    from sklearn.preprocessing import KBinsDiscretizer
    from cuml_b200.preprocessing.discretization import edges_from_centers, uniform_init
    rng = np.random.default_rng(2)
    col = np.concatenate([rng.normal(m, 0.3, size=300) for m in (-4.0, 0.0, 1.5, 6.0)])
    kb = KBinsDiscretizer(n_bins=4, strategy="kmeans", encode="ordinal").fit(col[:, None])
    init = uniform_init(col.min(), col.max(), 4)
    km = SkKMeans(n_clusters=4, init=init, n_init=1).fit(col[:, None])
    edges = edges_from_centers(km.cluster_centers_[:, 0], col.min(), col.max(), 4)
    np.testing.assert_allclose(edges, kb.bin_edges_[0], rtol=1e-12)
    # duplicate centres collapse a bin (and warn), as the reference does
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        e = edges_from_centers([1.0, 1.0, 1.0, 3.0], 0.0, 4.0, 4)
    assert e.tolist() == [0.0, 1.0, 2.0, 4.0] and any("too small" in str(w.message) for w in rec)


@pytest.mark.gpu
def test_kmeans_bin_edges_gpu():
    from sklearn.preprocessing import KBinsDiscretizer
    from cuml_b200.preprocessing import kmeans_bin_edges
    rng = np.random.default_rng(2)
    cols = [np.concatenate([rng.normal(m, 0.3, size=400) for m in means])
            for means in ((-4.0, 0.0, 1.5, 6.0, 9.0), (0.0, 10.0, 20.0, 30.0, 45.0))]
    X = np.stack(cols, axis=1).astype(np.float32)
    edges = kmeans_bin_edges(X, 5)
    kb = KBinsDiscretizer(n_bins=5, strategy="kmeans", encode="ordinal").fit(X.astype(np.float64))
    for j in range(2):
        assert len(edges[j]) == 6 and np.all(np.diff(edges[j]) > 0)
        span = float(X[:, j].max() - X[:, j].min())
        assert np.abs(edges[j] - kb.bin_edges_[j]).max() <= 1e-2 * span
    const = np.stack([X[:, 0], np.full(X.shape[0], 3.0, np.float32)], axis=1)
    with warnings.catch_warnings(record=True) as rec:
        warnings.simplefilter("always")
        e = kmeans_bin_edges(const, [5, 3])
    assert np.array_equal(e[1], [-np.inf, np.inf]) and any("constant" in str(w.message) for w in rec)


@pytest.mark.gpu
def test_kmeans_sampling_gpu():
    from sklearn.metrics import adjusted_rand_score
    from cuml_b200.explainer import kmeans_sampling
    from oracle import blobs
    n, d, k = 6000, 8, 4
    X, centres, true = blobs.make_blobs(n, d, k)
    Xm = X.copy()
    Xm[::97, 3] = np.nan                                     # missing values are imputed, not propagated
    summary, names, labels = kmeans_sampling(Xm, k, round_values=True, detailed=True, random_state=0)
    assert summary.shape == (k, d) and names == [str(i) for i in range(d)] and labels.shape == (n,)
    assert np.isfinite(summary).all()
    assert labels.min() >= 0 and labels.max() < k and len(np.unique(labels)) == k
    col_mean3 = np.nanmean(Xm[:, 3].astype(np.float64))
    for j in range(d):                                       # every coordinate is a value that occurs in its column
        for i in range(k):
            if j == 3:                                       # ... or the imputed column mean (fp32 on the device)
                pool = np.where(np.isnan(Xm[:, j]), col_mean3, Xm[:, j].astype(np.float64))
                assert np.abs(pool - summary[i, j]).min() <= 1e-4
            else:
                assert (Xm[:, j] == summary[i, j]).any()
    if adjusted_rand_score(true, labels) >= 0.99:            # the usual outcome: one summary row per blob
        dist = np.sqrt(((summary[:, None, :] - centres[None, :, :].astype(np.float64)) ** 2).sum(-1))
        assert sorted(dist.argmin(1).tolist()) == list(range(k)) and dist.min(1).max() < 1.5
    plain = kmeans_sampling(X, k, round_values=False, random_state=0)
    assert plain.shape == (k, d)
