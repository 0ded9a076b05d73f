"""World-size-2 CPU (gloo) tests of ``cuml_b200.distributed.KMeans`` -- the orchestration that replaces
``cuml.dask.cluster.KMeans`` (reference python/cuml/cuml/dask/cluster/kmeans.py:150-262,341-366).

What runs here is the product's orchestration code (global weight normalisation, the shared random_state,
the global row-count check, the collective preflight, the inertia / score sums) and the real ``KMeansMG``
parameter / row validation; only the two calls that enter libcuml_b200 are replaced by the oracle
(``_fit_mg_parts`` and the single-GPU predict), because there is no GPU in the CPU suite.  The same class runs
on GPUs in tests/test_kmeans_mg_gpu.py."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _stand_ins():
    """oracle-backed stand-ins for the two estimator classes (test infrastructure)"""
    import torch
    import torch.distributed as dist
    from cuml_b200.cluster.kmeans import KMeans
    from cuml_b200.cluster.kmeans_mg import KMeansMG
    from cuml_b200.distributed import KMeans as DistKMeans
    from oracle import lloyd

    class OracleMG(KMeansMG):
        seen_seeds = []

        def _fit_mg_parts(self, parts, sample_weight_parts=None):
            self._validate_fit_params()
            OracleMG.seen_seeds.append(self._c_params_seed())
            X = np.concatenate([np.asarray(p, dtype=np.float64) for p in parts])
            w = None if sample_weight_parts is None else np.concatenate(
                [np.asarray(p, dtype=np.float64) for p in sample_weight_parts])
            C = np.asarray(self.init, dtype=np.float64).copy()
            k, d = C.shape
            for _ in range(self.max_iter):
                labels, dmin = lloyd.e_step(X, C)
                S, W, _ = lloyd.m_step(X, labels, k, w)
                packed = torch.from_numpy(np.concatenate([S.ravel(), W]))
                dist.all_reduce(packed)
                p = packed.numpy()
                S, W = p[:k * d].reshape(k, d), p[k * d:]
                nz = W > 0
                C[nz] = S[nz] / W[nz, None]
            labels, dmin = lloyd.e_step(X, C)
            self._centers = torch.from_numpy(C)
            self._labels = torch.from_numpy(labels)
            self._in_kind = "numpy"
            self.inertia_ = float(dmin.sum() if w is None else (dmin * w).sum())   # weights NOT re-normalised
            self.n_iter_ = self.max_iter
            self.n_features_in_ = d
            return self

        def _c_params_seed(self):
            from cuml_b200.cluster.kmeans import check_random_seed
            return check_random_seed(self.random_state)

    class OracleSG(KMeans):
        def _predict_labels_inertia(self, X, sample_weight=None):
            return lloyd.predict(np.asarray(X), self._centers.numpy(), sample_weight, normalize=True)

        def transform(self, X):
            return lloyd.transform(np.asarray(X), self._centers.numpy())

    class Orchestrator(DistKMeans):
        _mg_class = OracleMG
        _sg_class = OracleSG

        def _get_handle(self):
            return None

    return Orchestrator, OracleMG


def _worker(rank, world, port, q, case):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = None
    try:
        from cuml_b200.cluster.kmeans_mg import shard_bounds
        from oracle import blobs
        Orchestrator, OracleMG = _stand_ins()
        n, d, k = 3001, 8, 5
        X, centres, _ = blobs.make_blobs(n, d, k)
        w = np.random.default_rng(3).uniform(0.5, 2.0, size=n)
        init = blobs.parity_init(centres)
        lo, hi = shard_bounds(n, rank, world)
        mid = lo + (hi - lo) // 3
        if case == "fit":
            km = Orchestrator(n_clusters=k, init=init, max_iter=4, tol=0.0, random_state=None)
            km.fit([X[lo:mid], X[mid:hi]], sample_weight=[w[lo:mid], w[mid:hi]])
            score = km.score([X[lo:mid], X[mid:hi]], sample_weight=[w[lo:mid], w[mid:hi]])
            pred = km.predict(X[lo:hi])
            tr = km.transform(X[lo:lo + 7])
            out = dict(centers=km.cluster_centers_, inertia=km.inertia_, labels=np.asarray(km.labels_),
                       seed=OracleMG.seen_seeds[-1], score=score, pred=np.asarray(pred), tr=tr,
                       n_iter=km.n_iter_)
        elif case == "too_few_rows":
            km = Orchestrator(n_clusters=4000, init="random", random_state=1)
            try:
                km.fit(X[lo:hi])
            except ValueError as e:
                out = dict(error=str(e))
        elif case == "preflight_rank1":
            # init='random' with k = 6 over 2 ranks: each rank must hold 3 rows; rank 1 holds only 2
            km = Orchestrator(n_clusters=6, init="random", random_state=1)
            try:
                km.fit(X[:100] if rank == 0 else X[100:102])
            except ValueError as e:
                out = dict(error=str(e))
        elif case == "mg_param":
            km = Orchestrator(n_clusters=k, init="k-means++", random_state=1)
            try:
                km.fit(X[lo:hi])
            except ValueError as e:
                out = dict(error=str(e))
        elif case == "weights_on_one_rank":
            km = Orchestrator(n_clusters=k, init=init, max_iter=1, tol=0.0, random_state=1)
            try:
                km.fit(X[lo:hi], sample_weight=w[lo:hi] if rank == 0 else None)
            except ValueError as e:
                out = dict(error=str(e))
    except Exception as e:   # a test bug must not leave the other rank waiting for the queue
        out = dict(crash=repr(e))
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def _run(case):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, case)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=180) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    outs = [o for _, o in outs]
    for o in outs:
        assert o is not None and "crash" not in o, o
    return outs


def test_distributed_fit_score_predict_gloo():
    from oracle import blobs, lloyd
    a, b = _run("fit")
    n, d, k = 3001, 8, 5
    X, centres, _ = blobs.make_blobs(n, d, k)
    w = np.random.default_rng(3).uniform(0.5, 2.0, size=n)
    ref = lloyd.fit(X, blobs.parity_init(centres), max_iter=4, tol=0.0, sample_weight=w)   # sum(w) == n globally
    assert np.array_equal(a["centers"], b["centers"])
    np.testing.assert_allclose(a["centers"], ref["centroids"], rtol=1e-12, atol=1e-12)
    # inertia_ is the sum of the rank-local inertias under the GLOBAL weight normalisation
    assert a["inertia"] == b["inertia"]
    assert abs(a["inertia"] - ref["inertia"]) / ref["inertia"] < 1e-12
    assert np.array_equal(np.concatenate([a["labels"], b["labels"]]), ref["labels"])
    assert np.array_equal(np.concatenate([a["pred"], b["pred"]]), ref["labels"])
    # random_state=None: rank 0's draw is used everywhere
    assert a["seed"] == b["seed"]
    assert a["n_iter"] == b["n_iter"] == 4
    # score: global normalisation, then every partition scored by the single-GPU model (which normalises its
    # own partition again) -- reference dask/cluster/kmeans.py:341-366
    from cuml_b200.cluster.kmeans_mg import shard_bounds
    wn = w * (n / w.sum())
    expect = 0.0
    for r in range(2):
        lo, hi = shard_bounds(n, r, 2)
        mid = lo + (hi - lo) // 3
        for s, e in ((lo, mid), (mid, hi)):
            expect += -lloyd.predict(X[s:e], ref["centroids"], wn[s:e], normalize=True)[1]
    assert a["score"] == b["score"]
    assert abs(a["score"] - expect) / abs(expect) < 1e-12
    np.testing.assert_allclose(a["tr"], lloyd.transform(X[:7], ref["centroids"]), rtol=1e-12)


def test_distributed_global_row_check_gloo():
    a, b = _run("too_few_rows")
    assert a["error"] == b["error"]
    assert a["error"].startswith("n_samples=3001 should be >= n_clusters=4000.")


def test_distributed_preflight_raises_on_every_rank_gloo():
    a, b = _run("preflight_rank1")
    assert "init='random' requires rank 1 to sample up to 3 initial centroid(s)" in b["error"]
    assert a["error"].startswith("[rank 1] ") and b["error"] in a["error"]
    a, b = _run("mg_param")
    assert "init='k-means++' is not supported for KMeansMG" in a["error"]
    assert "init='k-means++' is not supported for KMeansMG" in b["error"]
    a, b = _run("weights_on_one_rank")
    assert a["error"] == "sample_weight must be passed on every rank or on none"
    assert b["error"] == "[rank 0] " + a["error"]
