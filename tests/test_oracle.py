"""Pins the CPU oracle (oracle/lloyd.py) against the reference's known answers and its CPU
execution path.  CPU only."""
import glob
import os

import numpy as np
import pytest

from oracle import blobs, lloyd, sklearn_ref


def test_sklearn_kat_weighted_toy():
    # sklearn/cluster/tests/test_k_means.py:62-80 (hand-computed); inertia 0.375 is for raw
    # weights; under the GPU path's sum(w)=n normalisation it is 0.1875.
    X = np.array([[0, 0], [0.5, 0], [0.5, 1], [1, 1]], dtype=np.float32)
    w = np.array([3, 1, 1, 3], dtype=np.float32)
    C0 = np.array([[0, 0], [1, 1]], dtype=np.float32)
    r = lloyd.fit(X, C0, max_iter=300, tol=1e-4, sample_weight=w, rule="sklearn")
    assert r["labels"].tolist() == [0, 0, 1, 1]
    np.testing.assert_allclose(r["inertia"], 0.375)
    np.testing.assert_allclose(r["centroids"], [[0.125, 0], [0.875, 1]])
    assert r["n_iter"] == 2
    r = lloyd.fit(X, C0, max_iter=300, tol=1e-4, sample_weight=w, rule="cuvs")
    assert r["labels"].tolist() == [0, 0, 1, 1]
    np.testing.assert_allclose(r["inertia"], 0.1875)
    np.testing.assert_allclose(r["centroids"], [[0.125, 0], [0.875, 1]])
    assert r["n_iter"] == 2


def test_doctest_kat():
    # python/cuml/cuml/cluster/kmeans.pyx:464-491: 4x2 toy -> labels [0,0,1,1],
    # centres [[1,1.5],[3.5,2.5]] (label order depends on the seeded init; compare as sets)
    X = np.array([[1.0, 1.0], [1.0, 2.0], [3.0, 2.0], [4.0, 3.0]], dtype=np.float32)
    r = lloyd.fit(X, X[[0, 3]].copy(), max_iter=300, tol=1e-4)
    assert r["labels"].tolist() == [0, 0, 1, 1]
    np.testing.assert_allclose(r["centroids"], [[1.0, 1.5], [3.5, 2.5]])


def test_cpp_example_kat():
    # cpp/examples/kmeans/kmeans_example.cpp:110-113,172-191: rows (1,1),(3,4),(1,2),(2,3), k=2,
    # tol 0.05 -> labels {0,1,0,1}, centroids {1,1.5,2.5,3.5}
    X = np.array([[1.0, 1.0], [3.0, 4.0], [1.0, 2.0], [2.0, 3.0]], dtype=np.float64)
    r = lloyd.fit(X, X[[0, 1]].copy(), max_iter=300, tol=0.05)
    assert r["labels"].tolist() == [0, 1, 0, 1]
    np.testing.assert_allclose(r["centroids"].ravel(), [1.0, 1.5, 2.5, 3.5], rtol=1e-15)
    lab, inertia = lloyd.predict(X, r["centroids"])
    np.testing.assert_allclose(inertia, 0.25 * 2 + 0.5 * 2)
    np.testing.assert_allclose(lloyd.transform(X, r["centroids"], sqrt=True)[0, 0], 0.5)


def test_empty_cluster_keeps_old_centroid():
    # xfail-list.yaml:679-682: GPU path does not relocate empty clusters
    X = np.array([[0, 0], [0.5, 0], [0.5, 1], [1, 1]], dtype=np.float64)
    C0 = np.array([[0.5, 0.5], [3, 3]])
    r = lloyd.fit(X, C0, max_iter=10, tol=1e-9)
    np.testing.assert_allclose(r["centroids"][1], [3, 3])
    np.testing.assert_allclose(r["centroids"][0], [0.5, 0.5])
    assert set(r["labels"].tolist()) == {0}


def test_first_min_tie_break():
    X = np.array([[0.0, 0.0]], dtype=np.float32)
    C = np.array([[1.0, 0.0], [-1.0, 0.0], [0.0, 1.0]], dtype=np.float32)
    lab, d = lloyd.e_step(X, C)
    assert lab[0] == 0 and d[0] == 1.0


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "sk_*.npz"))))
def test_golden_fixtures(path):
    g = np.load(path)
    w = g["sample_weight"] if g["sample_weight"].size else None
    n = g["X"].shape[0]
    r = lloyd.fit(g["X"], g["init"], max_iter=int(g["max_iter"]), tol=1e-12, sample_weight=w)
    agree = (r["labels"] == g["labels64"]).mean()
    assert agree == 1.0
    assert (r["labels"] == g["labels"]).mean() >= 0.9999
    scale = np.abs(g["centroids64"]).max()
    assert np.abs(r["centroids"] - g["centroids64"]).max() / scale < 1e-12
    assert np.abs(r["centroids"] - g["centroids"]).max() / scale < 1e-4
    # sklearn uses raw weights; the GPU rule normalises to sum(w)=n
    ref_inertia = float(g["inertia64"]) * (1.0 if w is None else n / float(w.astype(np.float64).sum()))
    assert abs(r["inertia"] - ref_inertia) / ref_inertia < 1e-10
    ref32 = float(g["inertia"]) * (1.0 if w is None else n / float(w.astype(np.float64).sum()))
    assert abs(r["inertia"] - ref32) / ref32 < 1e-5


def test_live_sklearn_regime1():
    # C1-shaped but small: the reference CPU path run live (SURVEY 8c regime 1)
    X, centres, _ = blobs.make_blobs(20000, 32, 16)
    init = blobs.parity_init(centres)
    sk = sklearn_ref.fit(X, init, max_iter=50, tol=0.0)
    r = lloyd.fit(X, init, max_iter=50, tol=1e-12)
    assert (r["labels"] == sk["labels"]).mean() >= 0.9999
    assert abs(r["inertia"] - sk["inertia"]) / sk["inertia"] < 1e-5
    assert np.abs(r["centroids"] - sk["centroids"]).max() / np.abs(sk["centroids"]).max() < 1e-4


def test_single_step_matches_sklearn_lloyd_iter():
    # one Lloyd step against sklearn's own lloyd_iter (max_iter=1): anywhere, incl. regime 2
    X, _, _ = blobs.make_blobs(5000, 16, 8)
    init = blobs.throughput_init(X, 8)
    sk = sklearn_ref.fit(X.astype(np.float64), init.astype(np.float64), max_iter=1, tol=0.0)
    _, _, _, C1, _, _ = lloyd.lloyd_step(X, init)
    # sklearn centres X before fitting: results agree to fp64 rounding
    np.testing.assert_allclose(C1, sk["centroids"], rtol=0, atol=1e-9)


def test_m_step_matches_bruteforce():
    rng = np.random.default_rng(0)
    X = rng.standard_normal((1000, 5)).astype(np.float32)
    lab = rng.integers(0, 7, 1000)
    w = rng.uniform(0.1, 3, 1000)
    S, W, _ = lloyd.m_step(X, lab, 8, w)
    for j in range(8):
        np.testing.assert_allclose(S[j], (X[lab == j].astype(np.float64) * w[lab == j, None]).sum(0), atol=1e-9)
        np.testing.assert_allclose(W[j], w[lab == j].sum(), atol=1e-12)


def test_label_disagreement_checker():
    X, centres, _ = blobs.make_blobs(3000, 8, 4)
    lab, _ = lloyd.e_step(X, centres)
    agree, bad = lloyd.label_disagreements_ok(X, centres, lab, 1e-6)
    assert agree == 1.0 and bad == 0
    lab2 = lab.copy()
    lab2[:5] = (lab2[:5] + 1) % 4
    agree, bad = lloyd.label_disagreements_ok(X, centres, lab2, 1e-6)
    assert bad == 5


def test_blobs_deterministic():
    X1, c1, l1 = blobs.make_blobs(1000, 8, 4)
    X2, c2, l2 = blobs.make_blobs(1000, 8, 4)
    assert np.array_equal(X1, X2) and np.array_equal(c1, c2) and np.array_equal(l1, l2)
    assert X1.dtype == np.float32 and X1.flags.c_contiguous


def test_transform_and_predict_match_sklearn():
    # the oracle's transform / predict against the reference CPU path's own methods (sklearn.cluster.KMeans):
    # sklearn.transform returns Euclidean distances (== ML::kmeans::transform under L2SqrtExpanded)
    from sklearn.cluster import KMeans as SK
    from oracle import blobs, lloyd
    X, centres, _ = blobs.make_blobs(3000, 12, 9, seed=77)
    init = blobs.parity_init(centres)
    sk = SK(n_clusters=9, init=init.astype(np.float64), n_init=1, max_iter=5, tol=0.0, algorithm="lloyd").fit(
        X.astype(np.float64))
    Cc = sk.cluster_centers_
    assert np.allclose(lloyd.transform(X, Cc, sqrt=True), sk.transform(X.astype(np.float64)), rtol=1e-9, atol=1e-9)
    assert np.allclose(lloyd.transform(X, Cc) , sk.transform(X.astype(np.float64)) ** 2, rtol=1e-9, atol=1e-8)
    lab, inertia = lloyd.predict(X, Cc)
    assert np.array_equal(lab, sk.predict(X.astype(np.float64)))
    assert abs(inertia + sk.score(X.astype(np.float64))) / inertia < 1e-10


def test_e_step_random_shapes_against_bruteforce():
    # the chunked fp64 E-step equals a direct (x - c)^2 argmin on random small shapes, including k = 1 and d = 1
    from oracle import lloyd
    rng = np.random.default_rng(5)
    for n, d, k in [(1, 1, 1), (17, 1, 3), (50, 7, 1), (300, 5, 11), (1000, 33, 64)]:
        X = rng.standard_normal((n, d)).astype(np.float32)
        Cc = rng.standard_normal((k, d)).astype(np.float32)
        lab, dmin = lloyd.e_step(X, Cc)
        d2 = ((X[:, None, :].astype(np.float64) - Cc[None].astype(np.float64)) ** 2).sum(2)
        assert np.array_equal(lab, d2.argmin(1))
        assert np.allclose(dmin, d2.min(1), rtol=1e-9, atol=1e-9)

