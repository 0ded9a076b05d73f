"""First-contact GPU diagnostics: each stage runs in its own subprocess with a timeout so a
trapped kernel cannot poison later stages.  Writes gpurun_out/gpu_check.json.

    python tools/gpu_check.py [stage ...]
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")


def _setup():
    import ctypes as C
    import numpy as np
    import torch
    from cuml_b200 import _lib
    lib = _lib.load()
    h = _lib.Handle()
    return C, np, torch, _lib, lib, h


def stage_simt():
    """SIMT E-step + M-step + fit vs the oracle on small blobs."""
    C, np, torch, _lib, lib, h = _setup()
    from oracle import blobs, lloyd
    res = {}
    for (n, d, k) in [(2000, 8, 5), (5000, 32, 16), (3001, 20, 3), (4000, 64, 40), (1500, 7, 9)]:
        X, centres, _ = blobs.make_blobs(n, d, k)
        init = blobs.parity_init(centres)
        Xd = torch.from_numpy(X).cuda()
        Cd = torch.from_numpy(init).cuda()
        labels = torch.zeros(n, dtype=torch.int32, device="cuda")
        packed = torch.zeros(k * d + k + 1, dtype=torch.float64, device="cuda")
        shift = torch.zeros(1, dtype=torch.float64, device="cuda")
        _lib.check(lib.cuml_b200_kmeans_lloyd_step_f32(h.ptr, Xd.data_ptr(), n, d, None, k, Cd.data_ptr(),
                                                       labels.data_ptr(), packed.data_ptr(), shift.data_ptr(), 1))
        h.sync()
        lab_o, S, W, C_new, inertia, shift2 = lloyd.lloyd_step(X, init)
        lab = labels.cpu().numpy()
        pk = packed.cpu().numpy()
        res[f"{n}x{d}k{k}"] = dict(
            label_agree=float((lab == lab_o).mean()),
            S_err=float(np.abs(pk[:k * d].reshape(k, d) - S).max() / np.abs(S).max()),
            W_err=float(np.abs(pk[k * d:k * d + k] - W).max()),
            inertia_rel=float(abs(pk[-1] - inertia) / inertia),
            C_err=float(np.abs(Cd.cpu().numpy() - C_new).max() / np.abs(C_new).max()),
            shift_rel=float(abs(shift.item() - shift2) / max(shift2, 1e-30)))
    return res


def stage_tc_layout():
    """tensor-core engine: structured inputs that expose operand-layout mistakes."""
    C, np, torch, _lib, lib, h = _setup()
    res = {}
    for (n, d, k) in [(128, 32, 32), (256, 32, 32), (300, 64, 64), (1000, 128, 256), (513, 16, 48), (700, 96, 300)]:
        rng = np.random.default_rng(0)
        X = rng.standard_normal((n, d)).astype(np.float32)
        Cc = rng.standard_normal((k, d)).astype(np.float32)
        Xd, Cd = torch.from_numpy(X).cuda(), torch.from_numpy(Cc).cuda()
        labels = torch.full((n,), -1, dtype=torch.int32, device="cuda")
        kp = C.c_int64()
        _lib.check(lib.cuml_b200_kmeans_debug_dots_f32(h.ptr, Xd.data_ptr(), n, d, k, Cd.data_ptr(), labels.data_ptr(),
                                                       None, C.byref(kp)))
        dots = torch.zeros((n, kp.value), dtype=torch.float32, device="cuda")
        _lib.check(lib.cuml_b200_kmeans_debug_dots_f32(h.ptr, Xd.data_ptr(), n, d, k, Cd.data_ptr(), labels.data_ptr(),
                                                       dots.data_ptr(), C.byref(kp)))
        ref = X.astype(np.float64) @ Cc.astype(np.float64).T
        got = dots.cpu().numpy()[:, :k].astype(np.float64)
        err = np.abs(got - ref)
        part = 0.5 * (Cc.astype(np.float64) ** 2).sum(1)[None, :] - ref
        lab_ref = part.argmin(1)
        lab = labels.cpu().numpy()
        res[f"{n}x{d}k{k}"] = dict(k_pad=int(kp.value), max_abs_err=float(err.max()),
                                   rel_err=float(err.max() / np.abs(ref).max()),
                                   mean_abs_err=float(err.mean()), label_agree=float((lab == lab_ref).mean()),
                                   pad_cols_zero=bool((dots.cpu().numpy()[:, k:] == 0).all()))
        if err.max() > 1e-2:
            # locate the damage
            bad = np.argwhere(err > 1e-2)
            res[f"{n}x{d}k{k}"]["first_bad"] = bad[:8].tolist()
            res[f"{n}x{d}k{k}"]["bad_rows"] = int(len(np.unique(bad[:, 0])))
            res[f"{n}x{d}k{k}"]["bad_cols"] = int(len(np.unique(bad[:, 1])))
            res[f"{n}x{d}k{k}"]["sample_got"] = got[:2, :4].tolist()
            res[f"{n}x{d}k{k}"]["sample_ref"] = ref[:2, :4].tolist()
    return res


def stage_tc_parity():
    """tensor-core engine vs the fp64 oracle on blobs (labels, sums, inertia), several shapes."""
    C, np, torch, _lib, lib, h = _setup()
    from oracle import blobs, lloyd
    res = {}
    for (n, d, k) in [(100000, 32, 16), (50000, 64, 256), (40000, 128, 1024), (60000, 16, 64), (30001, 48, 100)]:
        X, centres, _ = blobs.make_blobs(n, d, k)
        init = blobs.parity_init(centres)
        Xd = torch.from_numpy(X).cuda()
        out = {}
        for eng, name in ((2, "tc"), (1, "simt")):
            Cd = torch.from_numpy(init).cuda()
            labels = torch.zeros(n, dtype=torch.int32, device="cuda")
            packed = torch.zeros(k * d + k + 1, dtype=torch.float64, device="cuda")
            shift = torch.zeros(1, dtype=torch.float64, device="cuda")
            _lib.check(lib.cuml_b200_kmeans_lloyd_step_f32(h.ptr, Xd.data_ptr(), n, d, None, k, Cd.data_ptr(),
                                                           labels.data_ptr(), packed.data_ptr(), shift.data_ptr(), eng))
            h.sync()
            lab_o, S, W, C_new, inertia, shift2 = lloyd.lloyd_step(X, init)
            lab = labels.cpu().numpy()
            pk = packed.cpu().numpy()
            agree, n_bad = lloyd.label_disagreements_ok(X, init, lab, 2.0 ** -20)
            out[name] = dict(label_agree=agree, inexcusable=n_bad,
                             inertia_rel=float(abs(pk[-1] - inertia) / inertia),
                             C_err=float(np.abs(Cd.cpu().numpy() - C_new).max() / np.abs(C_new).max()))
        res[f"{n}x{d}k{k}"] = out
    return res


def stage_perf():
    """quick timing of the E-step and the full Lloyd step at a few shapes (CUDA events)."""
    C, np, torch, _lib, lib, h = _setup()
    res = {}
    for (n, d, k) in [(2_000_000, 64, 256), (1_000_000, 128, 1024), (4_000_000, 16, 64), (2_000_000, 32, 16)]:
        g = torch.Generator(device="cuda").manual_seed(1)
        cent = (torch.rand((k, d), device="cuda", generator=g) * 20 - 10)
        lab = torch.randint(0, k, (n,), device="cuda", generator=g)
        X = cent[lab] + torch.randn((n, d), device="cuda", generator=g)
        Cd = X[torch.randperm(n, device="cuda", generator=g)[:k]].clone()
        labels = torch.zeros(n, dtype=torch.int32, device="cuda")
        out = {}
        for eng, name in ((2, "tc"), (1, "simt")):
            if name == "simt" and n * k * d > 3e11:
                continue
            for what in ("assign", "step"):
                def run():
                    if what == "assign":
                        _lib.check(lib.cuml_b200_kmeans_assign_f32(h.ptr, X.data_ptr(), n, d, k, Cd.data_ptr(),
                                                                   labels.data_ptr(), eng))
                    else:
                        _lib.check(lib.cuml_b200_kmeans_lloyd_step_f32(h.ptr, X.data_ptr(), n, d, None, k,
                                                                       Cd.data_ptr(), labels.data_ptr(), None, None, eng))
                for _ in range(2):
                    run()
                h.sync()
                t0 = time.perf_counter()
                reps = 5
                for _ in range(reps):
                    run()
                h.sync()
                ms = (time.perf_counter() - t0) / reps * 1e3
                out[f"{name}_{what}_ms"] = ms
                if what == "assign":
                    out[f"{name}_assign_tflops_algo"] = 2.0 * n * k * d / (ms * 1e-3) / 1e12
                    out[f"{name}_assign_GBs"] = 4.0 * n * d / (ms * 1e-3) / 1e9
        res[f"{n}x{d}k{k}"] = out
    return res


def stage_estimator():
    """Python estimator end to end incl. seeding paths."""
    C, np, torch, _lib, lib, h = _setup()
    from cuml_b200.cluster import KMeans
    from oracle import blobs, lloyd
    from sklearn.metrics import adjusted_rand_score
    res = {}
    X, centres, true = blobs.make_blobs(20000, 32, 16)
    init = blobs.parity_init(centres)
    km = KMeans(n_clusters=16, init=init, max_iter=50, tol=0.0, n_init=1).fit(X)
    o = lloyd.fit(X, init, max_iter=50, tol=0.0)
    res["array"] = dict(inertia_rel=abs(km.inertia_ - o["inertia"]) / o["inertia"], n_iter=km.n_iter_,
                        label_agree=float((km.labels_ == o["labels"]).mean()),
                        C_err=float(np.abs(km.cluster_centers_ - o["centroids"]).max() / np.abs(o["centroids"]).max()))
    for init_name in ("k-means||", "k-means++", "random"):
        t0 = time.perf_counter()
        km = KMeans(n_clusters=16, init=init_name, random_state=3, n_init=2 if init_name != "random" else 10).fit(X)
        res[init_name] = dict(ari=float(adjusted_rand_score(true, km.labels_)), inertia=km.inertia_, n_iter=km.n_iter_,
                              seconds=time.perf_counter() - t0)
    T = km.transform(X[:100])
    res["transform_err"] = float(np.abs(T - lloyd.transform(X[:100], km.cluster_centers_)).max())
    res["score"] = float(km.score(X))
    w = np.random.default_rng(0).uniform(0.5, 2, len(X)).astype(np.float32)
    km = KMeans(n_clusters=16, init=init, max_iter=50, tol=1e-6).fit(X, sample_weight=w)
    o = lloyd.fit(X, init, max_iter=50, tol=1e-6, sample_weight=w)
    res["weighted"] = dict(inertia_rel=abs(km.inertia_ - o["inertia"]) / o["inertia"], n_iter=[km.n_iter_, o["n_iter"]])
    X64 = X.astype(np.float64)
    km = KMeans(n_clusters=16, init=init.astype(np.float64), max_iter=20, tol=1e-9).fit(X64)
    o = lloyd.fit(X64, init, max_iter=20, tol=1e-9)
    res["fp64"] = dict(inertia_rel=abs(km.inertia_ - o["inertia"]) / o["inertia"],
                       C_err=float(np.abs(km.cluster_centers_ - o["centroids"]).max()))
    return res


STAGES = dict(simt=stage_simt, tc_layout=stage_tc_layout, tc_parity=stage_tc_parity, perf=stage_perf,
              estimator=stage_estimator)


def main():
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 2 and sys.argv[1] == "--run":
        name = sys.argv[2]
        r = STAGES[name]()
        print("RESULT " + json.dumps(r))
        return
    names = sys.argv[1:] or list(STAGES)
    allres = {}
    for name in names:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--run", name], capture_output=True,
                               text=True, timeout=600)
            line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
            allres[name] = dict(rc=p.returncode, seconds=time.time() - t0,
                                result=json.loads(line[-1][7:]) if line else None,
                                stderr=p.stderr[-3000:], stdout="" if line else p.stdout[-3000:])
        except subprocess.TimeoutExpired as e:
            allres[name] = dict(rc="timeout", seconds=time.time() - t0, stdout=str(e.stdout)[-2000:],
                                stderr=str(e.stderr)[-2000:])
        with open(os.path.join(OUT, "gpu_check.json"), "w") as f:
            json.dump(allres, f, indent=1)
        print(name, json.dumps(allres[name], indent=1)[:6000])


if __name__ == "__main__":
    main()
