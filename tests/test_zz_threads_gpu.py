"""Concurrent use of the boundary from several host threads, one handle (CUDA stream) per thread -- the calling
convention of the reference's C++ API (wiki/cpp/DEVELOPER_GUIDE.md:11-20: algorithms are single-threaded per
raft::handle_t and callable concurrently with different handles; Python keeps a thread-local handle,
internals/base.py:20,37-40).  Collected last: written after round 1's GPU budget was spent."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_two_threads_two_streams_match_serial_fits():
    import torch
    from cuml_b200.cluster import KMeans
    from oracle import blobs
    shapes = [(60000, 64, 200), (90000, 16, 32)]           # CTA-pair kernel / row-packed single-CTA kernel
    data = []
    for n, d, k in shapes:
        X, centres, _ = blobs.make_blobs(n, d, k)
        data.append((X, blobs.parity_init(centres), k))

    def fit(X, init, k):
        km = KMeans(n_clusters=k, init=init, max_iter=8, tol=0.0, n_init=1).fit(X)
        return km.cluster_centers_, km.labels_, km.inertia_, km.predict(X[:5000])

    serial = [fit(*a) for a in data]
    out, errs = [None, None], []

    def worker(i):
        try:
            with torch.cuda.stream(torch.cuda.Stream()):      # own stream -> own thread-local handle
                for _ in range(3):
                    out[i] = fit(*data[i])
        except Exception as e:   # surfaced in the main thread
            errs.append(repr(e))

    # daemon threads: a worker that never returns fails the assertion below without keeping the interpreter alive
    threads = [threading.Thread(target=worker, args=(i,), daemon=True) for i in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
        assert not t.is_alive(), "worker thread did not finish"
    assert not errs, errs
    for got, ref in zip(out, serial):
        assert np.array_equal(got[0], ref[0])                 # deterministic kernels: bitwise the same model
        assert np.array_equal(got[1], ref[1]) and got[2] == ref[2]
        assert np.array_equal(got[3], ref[3])
