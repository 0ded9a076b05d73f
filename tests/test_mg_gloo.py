"""World-size-2 CPU (gloo) test of the row-sharded Lloyd protocol: each rank
reduces its own row shard, ONE sum-allreduce of the packed [S | W | inertia] buffer per
iteration, identical finalize on every rank == the unsharded iteration.  The per-rank
arithmetic here is the oracle (test infrastructure); on GPUs the same protocol runs inside
libcuml_b200 with NCCL (tests/test_kmeans_mg_gpu.py)."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cuml_b200.cluster.kmeans_mg import shard_bounds
    from oracle import blobs, lloyd
    n, d, k = 6001, 16, 7
    X, centres, _ = blobs.make_blobs(n, d, k)
    C = blobs.throughput_init(X, k).astype(np.float64)
    lo, hi = shard_bounds(n, rank, world)
    Xl = X[lo:hi]
    for _ in range(5):
        labels, dmin = lloyd.e_step(Xl, C)
        S, W, _ = lloyd.m_step(Xl, labels, k)
        packed = torch.from_numpy(np.concatenate([S.ravel(), W, [dmin.sum()]]))
        dist.all_reduce(packed, op=dist.ReduceOp.SUM)
        p = packed.numpy()
        S, W = p[:k * d].reshape(k, d), p[k * d:k * d + k]
        Cn = C.copy()
        nz = W > 0
        Cn[nz] = S[nz] / W[nz, None]
        C = Cn
    q.put((rank, C, float(p[-1])))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_lloyd_equals_unsharded_gloo():
    import torch.multiprocessing as mp
    from oracle import blobs, lloyd
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n, d, k = 6001, 16, 7
    X, _, _ = blobs.make_blobs(n, d, k)
    C = blobs.throughput_init(X, k).astype(np.float64)
    inertia = None
    for _ in range(5):
        _, _, _, C, inertia, _ = lloyd.lloyd_step(X, C)
    assert np.array_equal(outs[0][1], outs[1][1])          # bitwise identical on both ranks
    np.testing.assert_allclose(outs[0][1], C, rtol=1e-12, atol=1e-12)
    assert abs(outs[0][2] - inertia) / inertia < 1e-12
