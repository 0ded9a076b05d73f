"""CPU-only checks of the boundary: the C-ABI library builds for sm_100a, loads, exports every
symbol include/cuml_b200/kmeans_c.h declares, and fails loudly (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from cuml_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    return _lib.load()


def test_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "cuml_b200", "kmeans_c.h")).read()
    declared = set(re.findall(r"\b(cuml_b200_[a-z0-9_]+)\s*\(", header))
    declared -= {"cuml_b200_kmeans_params", "cuml_b200_handle"}
    assert declared, "no declarations parsed"
    for sym in sorted(declared):
        assert hasattr(lib, sym), sym
    assert declared == set(_lib.EXPORTED_SYMBOLS)


def test_params_default_match_reference(lib):
    # reference cpp/include/cuml/cluster/kmeans_params.hpp:17-32
    p = _lib.default_params()
    assert (p.metric, p.n_clusters, p.init, p.max_iter) == (0, 8, 0, 300)
    assert p.tol == 1e-4 and p.n_init == 1 and p.oversampling_factor == 2.0
    assert p.batch_samples == 1 << 15 and p.batch_centroids == 0
    assert p.init_size == 0 and p.device_buffer_samples == 0 and p.rng_seed == 0


def test_params_struct_layout_matches_header():
    # field order / C layout (natural alignment): 4x int32, double, int32 (+pad), 2x uint64, ...
    assert _lib.KMeansParams.metric.offset == 0
    assert _lib.KMeansParams.tol.offset == 16
    assert _lib.KMeansParams.rng_seed.offset == 32
    assert _lib.KMeansParams.oversampling_factor.offset == 56
    assert _lib.KMeansParams.init_size.offset == 72
    assert C.sizeof(_lib.KMeansParams) == 88


def test_tc_support_predicate(lib):
    assert lib.cuml_b200_kmeans_tc_supported(64, 256) == 1
    assert lib.cuml_b200_kmeans_tc_supported(128, 1024) == 1
    assert lib.cuml_b200_kmeans_tc_supported(30, 8) == 0   # n_features % 4 != 0 -> CUDA-core kernel


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    st = lib.cuml_b200_handle_create(C.byref(h), None, None, 0, 1)
    assert st == 2  # CUML_B200_CUDA_ERROR
    assert b"CUDA" in lib.cuml_b200_last_error()
    from cuml_b200.cluster import KMeans
    with pytest.raises(Exception):
        KMeans(n_clusters=2).fit(np.zeros((4, 2), np.float32))


def test_product_does_not_import_oracle():
    # the product path must never route through oracle/ (test infrastructure only)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "cuml_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                # scikit-learn's KMeans may only be *named*: the interop target of as_sklearn() and the public
                # name the accel proxy swaps (cuml_b200/accel/cluster.py install / uninstall) -- never fitted
                for allowed in ('_cpu_class_path = "sklearn.cluster.KMeans"',
                                "from sklearn.cluster._kmeans import KMeans as SkKMeans",
                                "import sklearn.cluster\n", "_saved[\"KMeans\"] = sklearn.cluster.KMeans\n",
                                "sklearn.cluster.KMeans = KMeans\n", "sklearn.cluster.KMeans = _saved.pop(\"KMeans\")\n",
                                "``sklearn.cluster.KMeans``"):
                    src = src.replace(allowed, "")
                assert "sklearn.cluster" not in src, f


def test_estimator_param_mapping():
    from cuml_b200.cluster import KMeans, KMeansMG
    p = KMeans(n_clusters=5, init="k-means++", random_state=7)._c_params()
    assert p.oversampling_factor == 0.0 and p.init == _lib.INIT_KMEANS_PLUS_PLUS and p.n_init == 1
    p = KMeans(init="random", random_state=7)._c_params()
    assert p.init == _lib.INIT_RANDOM and p.n_init == 10
    p = KMeans(init=np.zeros((8, 2)), random_state=7)._c_params()
    assert p.init == _lib.INIT_ARRAY and p.n_init == 10
    p = KMeans(init="k-means||", random_state=7, n_init=3)._c_params()
    assert p.init == _lib.INIT_KMEANS_PLUS_PLUS and p.oversampling_factor == 2.0 and p.n_init == 3
    with pytest.raises(ValueError, match="random_state"):
        KMeansMG(handle=None)._c_params()
    with pytest.raises(ValueError, match="k-means\\+\\+"):
        KMeansMG(handle=None, init="k-means++", random_state=1)._validate_fit_params()
    with pytest.raises(ValueError, match="oversampling_factor=0"):
        KMeansMG(handle=None, oversampling_factor=0, random_state=1)._validate_fit_params()
    with pytest.raises(ValueError, match="positive integer"):
        KMeans(n_clusters=0)._validate_fit_params()
    km = KMeans(n_clusters=3)
    assert km.get_params()["n_clusters"] == 3
    km.set_params(n_clusters=4, tol=0.0)
    assert km.n_clusters == 4 and km.tol == 0.0


def test_mg_random_init_split_rule():
    # reference python/cuml/cuml/cluster/kmeans_mg.py:63-81
    from cuml_b200.cluster.kmeans_mg import random_init_rows_required, shard_bounds, KMeansMG
    assert [random_init_rows_required(10, r, 4) for r in range(4)] == [4, 2, 2, 2]
    assert [random_init_rows_required(3, r, 8) for r in range(8)] == [1, 1, 1, 0, 0, 0, 0, 0]
    assert sum(random_init_rows_required(256, r, 8) for r in range(8)) == 256
    est = KMeansMG(handle=None, init="random", n_clusters=10, random_state=0)
    with pytest.raises(ValueError, match="init='random' requires rank 0"):
        est.validate(np.zeros((3, 2)), 0, 4)
    est.validate(np.zeros((4, 2)), 0, 4)
    b = [shard_bounds(10, r, 4) for r in range(4)]
    assert b == [(0, 3), (3, 6), (6, 8), (8, 10)]


def test_out_of_core_dispatch_predicate():
    # the estimator streams only host numpy fp32/fp64 matrices with more rows than device_buffer_samples
    import numpy as np
    from cuml_b200.cluster import KMeans
    from cuml_b200.cluster.kmeans_mg import KMeansMG
    X = np.zeros((20, 3), np.float32)
    assert KMeans(device_buffer_samples=10)._streams_from_host(X)
    assert not KMeans(device_buffer_samples=0)._streams_from_host(X)
    assert not KMeans(device_buffer_samples=20)._streams_from_host(X)          # fits the buffer: staged whole
    assert not KMeans(device_buffer_samples=10)._streams_from_host(X.astype(np.int32))
    assert not KMeans(device_buffer_samples=10)._streams_from_host(X.tolist())
    assert KMeans(device_buffer_samples=10)._streams_from_host(X.astype(np.float64))
    assert KMeansMG._multi_gpu and not KMeans._multi_gpu



def test_chunked_host_predict_logic(monkeypatch):
    # the chunk loop of the estimator's host predict (reference _kmeans_predict_host_chunked, kmeans.pyx:356-434)
    # with the C call replaced by a numpy nearest-centre: labels are stitched in order, weights are normalised once
    # over the whole input and the chunk inertias add up
    import numpy as np
    import torch
    from cuml_b200.cluster import kmeans as km_mod
    from cuml_b200.cluster import KMeans

    rng = np.random.default_rng(0)
    X = rng.standard_normal((1000, 5)).astype(np.float32)
    Cc = rng.standard_normal((7, 5)).astype(np.float32)
    w = rng.uniform(0.5, 2.0, 1000)
    calls = []

    def fake_predict(self, handle, params, Xb, wb, centers, normalize_weights=True):
        assert normalize_weights is False
        x, c = Xb.numpy().astype(np.float64), centers.numpy().astype(np.float64)
        d2 = ((x[:, None, :] - c[None, :, :]) ** 2).sum(2)
        lab = d2.argmin(1)
        ww = np.ones(len(x)) if wb is None else wb.numpy().astype(np.float64)
        calls.append(len(x))
        return torch.from_numpy(lab.astype(np.int32)), float((ww * d2[np.arange(len(x)), lab]).sum())

    class FakeHandle:
        def sync(self):
            pass

    monkeypatch.setattr(KMeans, "_c_predict", fake_predict)
    monkeypatch.setattr(KMeans, "_c_params", lambda self: None)
    monkeypatch.setattr(km_mod, "get_handle", lambda: FakeHandle())
    est = KMeans(n_clusters=7, device_buffer_samples=300)
    est._centers = torch.from_numpy(Cc)
    est._in_kind = "numpy"
    labels, inertia = est._predict_labels_inertia(X, sample_weight=w)
    d2 = ((X[:, None, :].astype(np.float64) - Cc[None].astype(np.float64)) ** 2).sum(2)
    assert calls == [300, 300, 300, 100]
    assert np.array_equal(np.asarray(labels), d2.argmin(1))
    wn = w * (len(X) / w.sum())
    assert abs(inertia - (wn * d2.min(1)).sum()) / inertia < 1e-6



@pytest.mark.parametrize("which,own", [("estep", "fused_l2_argmin_sm100"), ("mstep", "centroid_update_tma")])
def test_kernel_planners_hold_their_invariants_on_a_shape_grid(tmp_path, which, own):
    # tests/cpp/plan_check_*.cu include the kernel source (the planners are file-local) and walk ~8k (d, k) shapes on
    # the host: shared-memory budget, ring depths, thread counts.  Links against the library's other objects.
    import glob
    import shutil
    import subprocess
    from cuml_b200 import build
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("no nvcc")
    build.build()
    objs = [o for o in glob.glob(os.path.join(build.OBJDIR, "*.o")) if os.path.basename(o) != own + ".o"]
    assert objs, "library objects missing"
    exe = str(tmp_path / ("plan_check_" + which))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O1", "-std=c++17", "--expt-relaxed-constexpr",
           "-I" + os.path.join(ROOT, "include"), "-I" + build.CSRC, "-o", exe,
           os.path.join(ROOT, "tests", "cpp", "plan_check_%s.cu" % which), *objs, "-lcuda", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "bad 0" in r.stdout, r.stdout[-2000:]


def test_cpp_surface_has_the_reference_signatures(tmp_path):
    # all 14 ML::kmeans overloads with the reference's exact parameter types + KMeansParams fields / defaults
    # (tests/cpp/surface_signatures.cpp does not compile otherwise); runs without a GPU: nothing is called
    import shutil
    import subprocess
    from cuml_b200 import build
    gxx = shutil.which("g++")
    if gxx is None or not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"):
        pytest.skip("no g++ / CUDA headers")
    libdir = os.path.dirname(build.lib_path())
    exe = str(tmp_path / "surface_signatures")
    cmd = [gxx, "-std=c++17", "-I" + os.path.join(ROOT, "include"), "-I/usr/local/cuda/include",
           os.path.join(ROOT, "tests", "cpp", "surface_signatures.cpp"), "-L" + libdir, "-lcuml_b200",
           "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + libdir, "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "surface ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("src", ["kmeans_example.cpp", "kmeans_bench.cpp", "kmeans_mg_test.cpp"])
def test_cpp_surface_compiles_and_links(tmp_path, src):
    # the C++ ML::kmeans::* mirror (include/cuml/cluster/kmeans.hpp) against the built library: every forwarder the
    # examples use resolves to an exported C-ABI symbol (no GPU needed to compile and link)
    import shutil
    import subprocess
    from cuml_b200 import build
    gxx = shutil.which("g++")
    if gxx is None or not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"):
        pytest.skip("no g++ / CUDA headers")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(build.lib_path())
    exe = str(tmp_path / src.replace(".cpp", ""))
    cmd = [gxx, "-std=c++17", "-I" + os.path.join(root, "include"), "-I/usr/local/cuda/include",
           os.path.join(root, "examples", src), "-L" + libdir, "-lcuml_b200", "-L/usr/local/cuda/lib64", "-lcudart",
           "-Wl,-rpath," + libdir, "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
