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


# ---- spectral clustering's label assignment (reference cluster/spectral_clustering.pyx:113,337-349) -----------
@pytest.mark.gpu
def test_spectral_label_assignment_gpu():
    from sklearn.metrics import adjusted_rand_score
    from cuml_b200.cluster import assign_labels_kmeans
    rng = np.random.default_rng(5)
    k, per = 4, 300
    # what a spectral embedding of k well-separated groups looks like: rows near k points of R^k, small spread
    anchors = np.eye(k, dtype=np.float32) * 0.5 + 0.1
    true = np.repeat(np.arange(k), per)
    emb = (anchors[true] + rng.normal(0, 0.01, size=(k * per, k))).astype(np.float32)
    perm = rng.permutation(len(emb))
    labels = assign_labels_kmeans(emb[perm], n_clusters=k, n_init=10, random_state=42)
    assert labels.dtype == np.int32 and labels.shape == (k * per,)
    assert adjusted_rand_score(true[perm], labels) == 1.0
    assert np.array_equal(labels, assign_labels_kmeans(emb[perm], n_clusters=k, n_init=10, random_state=42))


# ---- the scikit-learn-facing proxy (reference accel/_overrides/sklearn/cluster.py:12-21) --------------------------
def test_accel_proxy_parameter_translation_and_sklearn_plumbing():
    from sklearn.base import clone
    from sklearn.exceptions import NotFittedError
    from sklearn.utils.validation import check_is_fitted
    from cuml_b200.accel import KMeans, UnsupportedOnGPU
    km = KMeans(5, init="k-means++", n_init=3, max_iter=17, tol=1e-3, random_state=4, algorithm="elkan")
    p = km._engine_params()
    assert p["init"] == "scalable-k-means++" and (p["n_clusters"], p["n_init"], p["max_iter"], p["tol"]) == (5, 3, 17, 1e-3)
    assert KMeans(init="random")._engine_params()["init"] == "random"
    arr = np.zeros((8, 2))
    assert KMeans(init=arr)._engine_params()["init"] is arr
    with pytest.raises(UnsupportedOnGPU):
        KMeans(init=lambda X, k, rs: X[:k])._engine_params()     # no CPU fallback: unsupported means an error
    with pytest.raises(UnsupportedOnGPU):
        KMeans(init="pca")._engine_params()
    twin = clone(km)
    assert twin is not km and twin.get_params() == km.get_params()
    assert km.set_params(n_clusters=6).n_clusters == 6
    with pytest.raises(NotFittedError):
        check_is_fitted(km)
    with pytest.raises(NotFittedError):
        km.predict(np.zeros((2, 2), np.float32))
    # same constructor parameters, in the same order, as the scikit-learn class it stands in for
    import inspect
    from sklearn.cluster._kmeans import KMeans as SkKMeans
    assert list(inspect.signature(KMeans.__init__).parameters) == list(inspect.signature(SkKMeans.__init__).parameters)


@pytest.mark.gpu
def test_accel_proxy_under_sklearn_code_gpu():
    """unmodified scikit-learn code -- Pipeline, GridSearchCV, clone, pickle -- on the proxy"""
    import pickle
    import sklearn.cluster
    from sklearn.metrics import adjusted_rand_score
    from sklearn.model_selection import GridSearchCV
    from sklearn.pipeline import make_pipeline
    from sklearn.preprocessing import StandardScaler
    from sklearn.utils.validation import check_is_fitted
    from cuml_b200 import accel
    from oracle import blobs
    n, d, k = 4000, 8, 4
    X, centres, true = blobs.make_blobs(n, d, k)
    real = sklearn.cluster.KMeans
    try:
        assert accel.install() is accel.KMeans and sklearn.cluster.KMeans is accel.KMeans
        from sklearn.cluster import KMeans                      # user code, unchanged
        pipe = make_pipeline(StandardScaler(), KMeans(n_clusters=k, random_state=0, n_init=3))
        labels = pipe.fit_predict(X)
        km = pipe[-1]
        check_is_fitted(km)
        assert adjusted_rand_score(true, labels) >= 0.99
        assert km.cluster_centers_.shape == (k, d) and isinstance(km.cluster_centers_, np.ndarray)
        assert km.labels_.dtype == np.int32 and km.n_features_in_ == d and km.n_iter_ >= 1
        Xs = pipe[0].transform(X)
        assert np.array_equal(pipe.predict(X), km.predict(Xs)) and np.array_equal(km.predict(Xs), labels)
        D = pipe.transform(X[:50])
        assert D.shape == (50, k) and np.array_equal(D.argmin(1), labels[:50])
        assert abs(-km.score(Xs) - km.inertia_) / km.inertia_ < 1e-5
        assert km.get_feature_names_out().tolist() == [f"kmeans{i}" for i in range(k)]
        gs = GridSearchCV(KMeans(random_state=0, n_init=1), {"n_clusters": [2, k]}, cv=2).fit(X)   # clone + score
        assert gs.best_params_ == {"n_clusters": k}
        km2 = pickle.loads(pickle.dumps(km))
        assert np.array_equal(km2.predict(Xs), labels)
        sk = km.as_sklearn()                                     # hand-over to the real scikit-learn class
        assert type(sk) is real and (sk.predict(Xs) == labels).mean() >= 0.9999   # its own fp32 E-step: near-ties may differ
    finally:
        accel.uninstall()
    assert sklearn.cluster.KMeans is real
