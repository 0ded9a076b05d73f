"""Deterministic synthetic blobs shared by the CUDA path and the oracle.

Follows BASELINE.md section 2 / SURVEY.md section 8(d): centres ~ U(-10, 10)^d (the box used by the
reference gbench, cpp/bench/sg/kmeans.cu:87-91), sigma = 1, equal-probability labels, fp32,
row-major.  Test infrastructure only (see oracle/__init__.py).
"""
from __future__ import annotations

import numpy as np

DATA_SEED = 1234
INIT_SEED = 42


def make_blobs(n, d, k, seed=DATA_SEED, sigma=1.0, dtype=np.float32, chunk=1 << 20):
    """Return (X [n,d], centres [k,d], true_labels [n])."""
    rng = np.random.default_rng(seed)
    centres = rng.uniform(-10.0, 10.0, size=(k, d))
    labels = rng.integers(0, k, size=n)
    X = np.empty((n, d), dtype=dtype)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        X[s:e] = (centres[labels[s:e]] + sigma * rng.standard_normal((e - s, d))).astype(dtype)
    return X, centres.astype(dtype), labels.astype(np.int64)


def throughput_init(X, k, seed=INIT_SEED):
    """k distinct data rows: a poor start, so every implementation runs all max_iter
    iterations.  NOT for centroid/label parity (SURVEY.md section 8c regime 2)."""
    rng = np.random.default_rng(seed)
    idx = rng.choice(X.shape[0], size=k, replace=False)
    return np.ascontiguousarray(X[np.sort(idx)])


def parity_init(centres, seed=INIT_SEED, jitter=0.5):
    """true centres + N(0, jitter^2): one centroid per blob, unique stable fixed point
    (SURVEY.md section 8c regime 1)."""
    rng = np.random.default_rng(seed)
    return (centres.astype(np.float64) + jitter * rng.standard_normal(centres.shape)).astype(centres.dtype)
