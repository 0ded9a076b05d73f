"""The Python side of the product, end to end on the CPU, against a stand-in for the C-ABI (tests/fake_capi.py).

The stand-in answers the library's compute calls from the oracle through the caller's raw buffers, so these tests
exercise exactly what the GPU tests exercise above the boundary -- argument order, dtype / index-width dispatch,
attribute plumbing, error strings -- and run the bodies of the estimator-level ``-m gpu`` tests unchanged."""
import numpy as np
import pytest

import fake_capi


@pytest.fixture()
def fake(monkeypatch):
    return fake_capi.install(monkeypatch)


def test_estimator_calls_the_boundary_with_the_reference_dispatch(fake):
    from cuml_b200.cluster import KMeans
    from oracle import blobs, lloyd
    X, centres, _ = blobs.make_blobs(3000, 12, 5)
    init = blobs.parity_init(centres)
    w = np.random.default_rng(0).uniform(0.5, 2.0, len(X)).astype(np.float32)
    km = KMeans(n_clusters=5, init=init, max_iter=7, tol=0.0).fit(X, sample_weight=w)
    ref = lloyd.fit(X, init, max_iter=7, tol=0.0, sample_weight=w)
    # ONE boundary call per fit: the fit hands back the labels / inertia of its own final pass (the reference needs a
    # second, redundant E-step for them, kmeans.pyx:803-812); int32 labels for small inputs (:277-281)
    assert fake.calls == ["cuml_b200_kmeans_fit_parts_labels_f32"]
    assert km.cluster_centers_.dtype == np.float32 and km.labels_.dtype == np.int32
    np.testing.assert_allclose(km.cluster_centers_, ref["centroids"], rtol=1e-6, atol=1e-6)
    assert np.array_equal(km.labels_, ref["labels"]) and km.n_iter_ == 7
    assert abs(km.inertia_ - ref["inertia"]) / ref["inertia"] < 1e-6
    assert np.array_equal(km.predict(X[:100]), ref["labels"][:100])
    np.testing.assert_allclose(km.transform(X[:50]), lloyd.transform(X[:50], km.cluster_centers_), rtol=1e-5)
    assert abs(-km.score(X, sample_weight=w) - ref["inertia"]) / ref["inertia"] < 1e-6
    assert fake.calls[-1] == "cuml_b200_kmeans_predict_f32_i32" and "cuml_b200_kmeans_transform_f32_i32" in fake.calls
    # fp64 input keeps fp64 through the boundary
    km64 = KMeans(n_clusters=5, init=init.astype(np.float64), max_iter=3, tol=0.0).fit(X.astype(np.float64))
    assert "cuml_b200_kmeans_fit_parts_labels_f64" in fake.calls and km64.cluster_centers_.dtype == np.float64
    # integer input is converted to fp32 (xfail-list.yaml:693-699)
    kmi = KMeans(n_clusters=2, init=np.array([[0, 0], [9, 9]]), max_iter=2, tol=0.0).fit(np.array([[0, 1], [1, 0], [9, 8], [8, 9]]))
    assert kmi.cluster_centers_.dtype == np.float32 and kmi.labels_.tolist() == [0, 0, 1, 1]


def test_partition_list_fit_single_rank(fake):
    from cuml_b200.cluster.kmeans_mg import KMeansMG
    from oracle import blobs, lloyd
    X, centres, _ = blobs.make_blobs(2000, 8, 4)
    init = blobs.parity_init(centres)
    km = KMeansMG(handle=fake_capi.FakeHandle(), n_clusters=4, init=init, max_iter=5, tol=0.0, random_state=1)
    km.fit([X[:700], X[700:700], X[700:]])                           # ragged, with an empty partition
    ref = lloyd.fit(X, init, max_iter=5, tol=0.0)
    assert fake.calls[0] == "cuml_b200_kmeans_fit_parts_f32"
    assert np.array_equal(km.labels_, ref["labels"])
    assert abs(km.inertia_ - ref["inertia"]) / ref["inertia"] < 1e-6
    assert abs(km.global_inertia_ - ref["inertia"]) / ref["inertia"] < 1e-6


def test_bodies_of_estimator_level_gpu_tests(fake):
    """the same assertions the GPU box checks, with the library replaced by the oracle: a failure here is a bug in
    the Python layer or in the test itself, found without spending GPU time"""
    import test_z_callers as test_callers
    import test_kmeans_gpu
    test_kmeans_gpu.test_fp64_fit_predict_transform_match_oracle()
    test_kmeans_gpu.test_predict_transform_score_match_oracle()
    test_kmeans_gpu.test_fit_stops_on_tolerance_like_oracle()
    test_kmeans_gpu.test_error_messages_match_reference()
    test_kmeans_gpu.test_doctest_kat_and_empty_cluster_rule()
    test_callers.test_kmeans_bin_edges_gpu()
    test_callers.test_kmeans_sampling_gpu()
    test_callers.test_spectral_label_assignment_gpu()
    test_callers.test_accel_proxy_under_sklearn_code_gpu()
    import glob
    import os
    for path in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "sk_*.npz"))):
        test_kmeans_gpu.test_fit_matches_reference_cpu_path_golden(path)


def test_distributed_estimator_world_size_one(fake, monkeypatch):
    import socket
    import torch.distributed as dist
    from cuml_b200.cluster import kmeans_mg
    from cuml_b200.distributed import KMeans as DistKMeans
    from cuml_b200.distributed import kmeans as dkm
    from oracle import blobs, lloyd
    monkeypatch.setattr(dkm, "comms_from_torch_distributed", lambda: fake_capi.FakeHandle())
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1)
    try:
        X, centres, _ = blobs.make_blobs(3000, 16, 6)
        w = np.random.default_rng(3).uniform(0.5, 2.0, size=len(X)).astype(np.float32)
        init = blobs.parity_init(centres)
        km = DistKMeans(n_clusters=6, init=init, max_iter=5, tol=0.0, random_state=None)
        km.fit([X[:1000], X[1000:]], sample_weight=[w[:1000], w[1000:]])
        ref = lloyd.fit(X, init, max_iter=5, tol=0.0, sample_weight=w)
        np.testing.assert_allclose(km.cluster_centers_, ref["centroids"], rtol=1e-5, atol=1e-5)
        assert abs(km.inertia_ - ref["inertia"]) / ref["inertia"] < 1e-5
        assert np.array_equal(np.asarray(km.labels_), ref["labels"])
        assert np.array_equal(np.asarray(km.predict(X)), ref["labels"])
        assert abs(-km.score(X, sample_weight=w) - ref["inertia"]) / ref["inertia"] < 1e-5
        assert np.asarray(km.transform(X[:4])).shape == (4, 6)
        km.close()
    finally:
        dist.destroy_process_group()
