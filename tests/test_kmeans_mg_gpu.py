"""Multi-rank tests of the row-sharded fit (reference contract: the multi-GPU model equals the single-GPU one,
cpp/tests/mg/kmeans_test.cu:116-137,167-193; python/cuml/tests/dask/test_dask_kmeans.py:54-126).

Topologies: "shared" = two ranks (two processes) on device 0 over the library's peer-memory communicator -- runs on a
one-GPU box; "peer" / "nccl" = one rank per device over the peer-memory communicator / NCCL, needs >= 2 GPUs."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


TOPOLOGIES = ["shared", "peer", "nccl"]


def _need(topology, world):
    import torch
    have = torch.cuda.device_count()
    if topology != "shared" and have < world:
        pytest.skip(f"topology {topology!r} needs {world} GPUs, {have} visible")


def _device_of(rank, topology):
    return 0 if topology == "shared" else rank


def _backend_of(topology):
    return "peer" if topology == "shared" else topology


def _worker(rank, world, port, q, init_kind, topology="nccl"):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(_device_of(rank, topology))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cuml_b200.cluster.kmeans_mg import KMeansMG, comms_from_torch_distributed, shard_bounds
    from oracle import blobs
    n, d, k = 40000, 32, 16
    X, centres, _ = blobs.make_blobs(n, d, k)
    lo, hi = shard_bounds(n, rank, world)
    h = comms_from_torch_distributed(backend=_backend_of(topology))
    assert h.comm_kind == _backend_of(topology)
    init = blobs.parity_init(centres) if init_kind == "array" else init_kind
    km = KMeansMG(handle=h, n_clusters=k, init=init, max_iter=10 if init_kind == "array" else 50,
                  tol=0.0 if init_kind == "array" else 1e-6, random_state=5,
                  n_init=1 if init_kind == "array" else 5)
    # two ragged local partitions per rank
    mid = lo + (hi - lo) // 3
    km.fit([X[lo:mid], X[mid:hi]])
    q.put((rank, km.cluster_centers_, km.inertia_, km.global_inertia_, km.labels_, km.n_iter_))
    dist.barrier()
    h.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("topology", TOPOLOGIES)
@pytest.mark.parametrize("init_kind", ["array", "k-means||", "random"])
def test_two_rank_fit_matches_single_gpu(init_kind, topology):
    _fit_ranks_match_single_gpu(2, init_kind, topology)


@pytest.mark.parametrize("world,topology", [(4, "peer"), (4, "nccl"), (8, "peer"), (8, "nccl")])
@pytest.mark.parametrize("init_kind", ["array", "k-means||"])
def test_many_rank_fit_matches_single_gpu(init_kind, world, topology):
    # the reference's MG gtest runs on however many ranks the communicator has (kmeans_test.cu:116-137); one rank per GPU
    _fit_ranks_match_single_gpu(world, init_kind, topology)


def _fit_ranks_match_single_gpu(world, init_kind, topology):
    _need(topology, world)
    import torch.multiprocessing as mp
    from cuml_b200.cluster import KMeans
    from oracle import blobs, lloyd
    from sklearn.metrics import adjusted_rand_score
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, init_kind, topology), daemon=True)
             for r in range(world)]
    for p in procs:
        p.start()
    try:
        outs = sorted([q.get(timeout=240) for _ in procs], key=lambda t: t[0])
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    finally:
        for p in procs:          # a rank that died or hangs must not outlive the test
            if p.is_alive():
                p.kill()
    n, d, k = 40000, 32, 16
    X, centres, true = blobs.make_blobs(n, d, k)
    for o in outs[1:]:
        assert np.array_equal(outs[0][1], o[1])    # identical centroids on every rank
    labels = np.concatenate([o[4] for o in outs])
    total_inertia = sum(o[2] for o in outs)
    assert abs(total_inertia - outs[0][3]) / outs[0][3] < 1e-5
    if init_kind == "array":
        ref = lloyd.fit(X, blobs.parity_init(centres), max_iter=10, tol=0.0)
        assert np.abs(outs[0][1] - ref["centroids"]).max() / np.abs(ref["centroids"]).max() <= 1e-4
        assert abs(total_inertia - ref["inertia"]) / ref["inertia"] <= 1e-5
        assert (labels == ref["labels"]).mean() >= 0.9999
    elif init_kind == "random":
        # random rows rarely seed one centroid per blob (k = 16): pin consistency, not optimality
        assert adjusted_rand_score(true, labels) >= 0.6
    else:
        assert adjusted_rand_score(true, labels) >= 0.99


# ---- cuml_b200.distributed.KMeans: the orchestration that stands in for cuml.dask.cluster.KMeans -----------------
def _dist_worker(rank, world, port, q, topology="nccl"):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["CUML_B200_COMM"] = _backend_of(topology)
    torch.cuda.set_device(_device_of(rank, topology))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = None
    try:
        from cuml_b200.cluster.kmeans_mg import shard_bounds
        from cuml_b200.distributed import KMeans as DistKMeans
        from oracle import blobs
        n, d, k = 30000, 32, 8
        X, centres, _ = blobs.make_blobs(n, d, k)
        w = np.random.default_rng(3).uniform(0.5, 2.0, size=n).astype(np.float32)
        lo, hi = shard_bounds(n, rank, world)
        mid = lo + (hi - lo) // 3
        km = DistKMeans(n_clusters=k, init=blobs.parity_init(centres), max_iter=6, tol=0.0, random_state=None)
        km.fit([X[lo:mid], X[mid:hi]], sample_weight=[w[lo:mid], w[mid:hi]])
        score = km.score(X[lo:hi], sample_weight=w[lo:hi])
        out = dict(centers=km.cluster_centers_, inertia=km.inertia_, labels=np.asarray(km.labels_),
                   pred=np.asarray(km.predict(X[lo:hi])), score=score, tr=np.asarray(km.transform(X[lo:lo + 5])),
                   n_iter=km.n_iter_)
        km.close()
    except Exception as e:   # report instead of leaving the parent waiting on the queue
        import traceback
        out = dict(crash=traceback.format_exc() + repr(e))
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,topology", [(1, "nccl"), (2, "shared"), (2, "peer"), (2, "nccl"), (8, "peer"), (8, "nccl")])
def test_distributed_estimator(world, topology):
    _need(topology, world)
    import torch.multiprocessing as mp
    from cuml_b200.cluster.kmeans_mg import shard_bounds
    from oracle import blobs, lloyd
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dist_worker, args=(r, world, port, q, topology), daemon=True) for r in range(world)]
    for p in procs:
        p.start()
    try:
        outs = sorted([q.get(timeout=240) for _ in procs], key=lambda t: t[0])
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    finally:
        for p in procs:          # a rank that died or hangs must not outlive the test
            if p.is_alive():
                p.kill()
    outs = [o for _, o in outs]
    for o in outs:
        assert o is not None and "crash" not in o, o
    n, d, k = 30000, 32, 8
    X, centres, _ = blobs.make_blobs(n, d, k)
    w = np.random.default_rng(3).uniform(0.5, 2.0, size=n).astype(np.float32)
    ref = lloyd.fit(X, blobs.parity_init(centres), max_iter=6, tol=0.0, sample_weight=w)
    for o in outs:
        assert np.array_equal(o["centers"], outs[0]["centers"])     # identical on every rank
        assert o["inertia"] == outs[0]["inertia"] and o["score"] == outs[0]["score"]
        assert o["n_iter"] == 6
    assert np.abs(outs[0]["centers"] - ref["centroids"]).max() / np.abs(ref["centroids"]).max() <= 1e-4
    assert abs(outs[0]["inertia"] - ref["inertia"]) / ref["inertia"] <= 1e-5
    labels = np.concatenate([o["labels"] for o in outs])
    assert (labels == ref["labels"]).mean() >= 0.9999
    assert (np.concatenate([o["pred"] for o in outs]) == ref["labels"]).mean() >= 0.9999
    # score: weights normalised globally, then each rank's rows scored (and re-normalised) by the local model
    wn = w.astype(np.float64) * (n / w.astype(np.float64).sum())
    expect = 0.0
    for r in range(world):
        lo, hi = shard_bounds(n, r, world)
        expect += -lloyd.predict(X[lo:hi], ref["centroids"], wn[lo:hi], normalize=True)[1]
    assert abs(outs[0]["score"] - expect) / abs(expect) <= 1e-5
    To = lloyd.transform(X[:5], ref["centroids"])
    assert np.abs(outs[0]["tr"] - To).max() / To.max() < 1e-5


# ---- the reference's multi-GPU gtest inputs through the C++ surface (examples/kmeans_mg_test.cpp) ---------------------
@pytest.mark.parametrize("ranks,mode", [(2, "shared"), (2, "peer"), (2, "nccl"), (8, "peer"), (8, "nccl")])
def test_cpp_mg_gtest_inputs_rank_processes(tmp_path, ranks, mode):
    # cpp/tests/mg/kmeans_test.cu:50-195: eight inputs, float and double, weighted and not, ARI >= 0.99 on every rank's
    # shard -- here with 2 or 8 rank processes ("shared": both on device 0 over the peer-memory communicator)
    import shutil
    import subprocess
    _need(mode, ranks)
    from cuml_b200 import build
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "kmeans_mg_test")
    libdir = os.path.dirname(build.lib_path())
    cmd = [gxx, "-O1", "-std=c++17", "-I" + os.path.join(root, "include"), "-I/usr/local/cuda/include",
           os.path.join(root, "examples", "kmeans_mg_test.cpp"), "-L" + libdir, "-lcuml_b200",
           "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + libdir, "-o", exe]
    subprocess.run(cmd, check=True, capture_output=True)
    r = subprocess.run([exe, str(ranks), mode], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
