"""``KMeansMG`` -- per-rank multi-GPU estimator, mirror of
``cuml.cluster.kmeans_mg.KMeansMG`` (reference python/cuml/cuml/cluster/kmeans_mg.py:9-81) and of
the partition-list fit ``KMeans._fit_mg_parts`` (reference kmeans.pyx:823-976).

One process per GPU.  The handle carries the NCCL communicator; ``comms_from_torch_distributed``
plays the role of ``raft_dask.common.comms.Comms`` (reference dask/cluster/kmeans.py:189-190):
rank 0 creates the NCCL unique id and ``torch.distributed`` ships it to the other ranks.
"""
from __future__ import annotations

import ctypes as C

from .. import _lib
from .kmeans import KMeans, _as_device_matrix, _torch


def comms_from_torch_distributed(stream=None, backend=None):
    """Create a Handle whose communicator spans the default torch.distributed group.

    backend: "nccl" (an NCCL communicator built from a unique id rank 0 creates) or "peer" (the library's own
    peer-memory collectives: every rank maps the others' exchange windows through CUDA IPC; one node only).  Default:
    the CUML_B200_COMM environment variable, else "peer" when two ranks share a device (NCCL refuses that topology),
    else "nccl".  torch.distributed only ships the rendezvous blobs."""
    import os
    import socket
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream or 1   # 0x1 = cudaStreamLegacy
    h = _lib.Handle(stream=stream, n_ranks=world, rank=rank)
    if world <= 1:
        return h
    where = [None] * world
    dist.all_gather_object(where, (socket.gethostname(), str(torch.cuda.get_device_properties(torch.cuda.current_device()).uuid)))
    one_node = len({w[0] for w in where}) == 1
    shared_device = len(set(where)) < world
    if backend is None:
        backend = os.environ.get("CUML_B200_COMM") or ("peer" if shared_device else "nccl")
    if backend == "peer":
        if not one_node:
            raise ValueError("the peer-memory communicator needs all ranks on one node")
        mine = h.peer_window_create(world)
        handles = [None] * world
        dist.all_gather_object(handles, mine)
        h.peer_window_attach(handles, rank, world)
        dist.barrier()          # every rank has mapped every window before the first exchange
    elif backend == "nccl":
        if shared_device:
            raise ValueError("NCCL cannot place two ranks on one device; use backend='peer'")
        box = [_lib.Handle.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        h.init_comm(box[0], rank, world)
    else:
        raise ValueError(f"unknown communicator backend {backend!r}")
    return h


def random_init_rows_required(n_clusters, rank, n_workers):
    """rows rank `rank` must hold for init='random' (reference kmeans_mg.py:63-81)."""
    n_sampling = min(n_workers, n_clusters)
    if rank >= n_sampling:
        return 0
    req = n_clusters // n_sampling
    if rank == 0:
        req += n_clusters % n_sampling
    return req


def shard_bounds(n_rows, rank, n_workers):
    """contiguous row block [lo, hi) of rank `rank`."""
    base, rem = divmod(n_rows, n_workers)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class KMeansMG(KMeans):
    """A multi-GPU implementation of KMeans: every rank calls ``fit`` on its local rows."""

    _multi_gpu = True

    def __init__(self, *, handle, **kwargs):
        self.handle = handle
        super().__init__(**kwargs)

    def fit(self, X, sample_weight=None):
        if isinstance(X, (list, tuple)):
            return self._fit_mg_parts(X, sample_weight_parts=sample_weight)
        return self._fit_mg_parts([X], sample_weight_parts=None if sample_weight is None else [sample_weight])

    def _validate_fit_params(self):
        super()._validate_fit_params()
        if isinstance(self.init, str):
            if self.init == "k-means++":
                raise ValueError("init='k-means++' is not supported for KMeansMG. "
                                 "Use init='k-means||' or init='scalable-k-means++'.")
            if self.init not in {"scalable-k-means++", "k-means||", "random"}:
                raise ValueError(f"init={self.init!r} is not supported for KMeansMG.")
        if self.oversampling_factor == 0:
            raise ValueError("oversampling_factor=0 is not supported for KMeansMG.")

    def validate(self, X, rank, n_workers):
        n_rows = sum(len(p) for p in X) if isinstance(X, (list, tuple)) else len(X)
        if not isinstance(self.init, str) or self.init != "random":
            return
        required = random_init_rows_required(self.n_clusters, rank, n_workers)
        if rank < min(n_workers, self.n_clusters) and n_rows < required:
            raise ValueError(
                f"init='random' requires rank {rank} to sample up to {required} initial centroid(s), but this "
                f"rank only has {n_rows} row(s). Repartition the data so each rank has enough rows for "
                f"initialization, reduce n_clusters, or provide explicit initial centers.")

    def _fit_mg_parts(self, parts, sample_weight_parts=None):
        torch = _torch()
        self._validate_fit_params()
        if len(parts) == 0:
            raise ValueError("`parts` must be a non-empty sequence of partitions")
        if sample_weight_parts is not None and len(sample_weight_parts) != len(parts):
            raise ValueError("sample_weight partitions must match the data partitions")
        ins = [_as_device_matrix(p) for p in parts]
        dtype = ins[0].t.dtype
        d = ins[0].t.shape[1]
        for i in ins:
            if i.t.dtype != dtype:
                raise ValueError("all partitions must share a dtype")
            if i.t.shape[1] != d:
                raise ValueError("all partitions must share n_features")
        ws = None
        if sample_weight_parts is not None:
            ws = [_as_device_matrix(w, dtype=dtype, name="sample_weight", ndim=1).t for w in sample_weight_parts]
        self.n_features_in_ = d
        centers = self._prepare_centers(ins[0].t)
        params = self._c_params()
        lib = _lib.load()
        f32 = dtype == torch.float32
        n_parts = len(ins)
        xp = (C.c_void_p * n_parts)(*[i.t.data_ptr() for i in ins])
        np_rows = (C.c_int64 * n_parts)(*[i.t.shape[0] for i in ins])
        wp = (C.c_void_p * n_parts)(*[w.data_ptr() for w in ws]) if ws is not None else None
        inertia = (C.c_float if f32 else C.c_double)()
        n_iter = C.c_int64()
        fn = lib.cuml_b200_kmeans_fit_parts_f32 if f32 else lib.cuml_b200_kmeans_fit_parts_f64
        _lib.check(fn(self.handle.ptr, C.byref(params), xp, np_rows, n_parts, d, wp, centers.data_ptr(),
                      C.byref(inertia), C.byref(n_iter)))
        # per-partition local predict, weights NOT re-normalised (reference kmeans.pyx:938-956)
        labels, local_inertia = [], 0.0
        for idx, i in enumerate(ins):
            if i.t.shape[0] == 0:
                labels.append(torch.zeros(0, dtype=torch.int32, device=i.t.device))
                continue
            lab, inn = self._c_predict(self.handle, params, i.t, ws[idx] if ws is not None else None, centers,
                                       normalize_weights=False)
            labels.append(lab)
            local_inertia += inn
        self.handle.sync()
        self._centers = centers
        self._labels = torch.cat(labels) if len(labels) > 1 else labels[0]
        self._in_kind = ins[0].kind
        self.inertia_ = local_inertia          # rank-local; the client sums (dask/cluster/kmeans.py:237-243)
        self.global_inertia_ = float(inertia.value)
        self.n_iter_ = int(n_iter.value)
        return self
