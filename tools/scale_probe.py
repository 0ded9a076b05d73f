"""Scale checks on one GPU: k-means|| + fit at 50M x 16 (C5 per-GPU shard is 25M), predict / transform timings."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from cuml_b200.cluster import KMeans

def blobs(n, d, k, seed=1):
    g = torch.Generator(device="cuda").manual_seed(seed)
    cent = torch.rand((k, d), device="cuda", generator=g) * 20 - 10
    lab = torch.randint(0, k, (n,), device="cuda", generator=g)
    X = torch.empty((n, d), device="cuda")
    for s in range(0, n, 1 << 22):
        e = min(n, s + (1 << 22))
        X[s:e] = cent[lab[s:e]] + torch.randn((e - s, d), device="cuda", generator=g)
    return X, lab, cent

for (n, d, k) in [(50_000_000, 16, 64), (10_000_000, 64, 256)]:
    X, lab, cent = blobs(n, d, k)
    for init in ("k-means||", "random"):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        km = KMeans(n_clusters=k, init=init, random_state=7, max_iter=20, tol=1e-4, n_init=1).fit(X)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        # purity-style check: fraction of rows whose cluster's majority true label matches
        pred = km.labels_.long()
        conf = torch.zeros((k, k), device="cuda", dtype=torch.int64)
        conf.index_put_((pred, lab), torch.ones_like(pred), accumulate=True)
        purity = conf.max(dim=1).values.sum().item() / n
        print(f"n={n} d={d} k={k} init={init}: fit {dt:.2f} s, n_iter {km.n_iter_}, inertia/n {km.inertia_/n:.3f} (ideal ~{d}), purity {purity:.4f}", flush=True)
    torch.cuda.synchronize(); t0 = time.perf_counter(); p = km.predict(X); torch.cuda.synchronize()
    print(f"  predict {time.perf_counter()-t0:.3f} s", flush=True)
    m = min(n, 2_000_000)
    torch.cuda.synchronize(); t0 = time.perf_counter(); T = km.transform(X[:m]); torch.cuda.synchronize()
    print(f"  transform {m} rows {time.perf_counter()-t0:.3f} s", flush=True)
    del X, lab, T, p
    torch.cuda.empty_cache()
