"""Generate golden fixtures by running the reference's CPU execution path (scikit-learn,
python/cuml/cuml/cluster/kmeans.pyx:604) on small seeded blobs.

Run from the repo root:  python tests/golden/make_golden.py
Outputs tests/golden/sk_*.npz (committed).  Each file holds X, init, sample_weight (or empty),
and sklearn's centroids / labels / inertia / n_iter for KMeans(init=array, n_init=1,
algorithm='lloyd', max_iter, tol), plus a float64 run of the same problem (inertia truth).
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import blobs, sklearn_ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [
    # name, n, d, k, weighted, max_iter, tol
    ("sk_small_d8_k5", 2000, 8, 5, False, 50, 0.0),
    ("sk_small_d32_k16", 4000, 32, 16, False, 50, 0.0),
    ("sk_weighted_d16_k7", 3000, 16, 7, True, 50, 0.0),
    ("sk_ragged_d20_k3", 1001, 20, 3, False, 50, 0.0),
]


def main():
    import sklearn
    for name, n, d, k, weighted, max_iter, tol in CASES:
        X, centres, _ = blobs.make_blobs(n, d, k, seed=blobs.DATA_SEED)
        init = blobs.parity_init(centres)
        w = None
        if weighted:
            w = np.random.default_rng(7).uniform(0.5, 2.0, size=n).astype(np.float32)
        r32 = sklearn_ref.fit(X, init, max_iter=max_iter, tol=tol, sample_weight=w)
        r64 = sklearn_ref.fit(X.astype(np.float64), init.astype(np.float64), max_iter=max_iter, tol=tol,
                              sample_weight=None if w is None else w.astype(np.float64))
        np.savez_compressed(
            os.path.join(HERE, name + ".npz"),
            X=X, init=init, sample_weight=np.zeros(0, np.float32) if w is None else w,
            centroids=r32["centroids"], labels=r32["labels"], inertia=r32["inertia"], n_iter=r32["n_iter"],
            centroids64=r64["centroids"], labels64=r64["labels"], inertia64=r64["inertia"],
            max_iter=max_iter, tol=tol, sklearn_version=sklearn.__version__)
        print(name, "n_iter", r32["n_iter"], "inertia", r32["inertia"], r64["inertia"])


if __name__ == "__main__":
    main()
