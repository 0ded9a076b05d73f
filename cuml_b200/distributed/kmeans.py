"""Multi-GPU ``KMeans`` for one-process-per-GPU programs -- the role of ``cuml.dask.cluster.KMeans``
(reference python/cuml/cuml/dask/cluster/kmeans.py:36-373) without Dask.

The reference's client object submits one task per Dask worker; here every rank of a
``torch.distributed`` group calls the same method on its own partitions (SPMD), and the steps the
reference runs on the scheduler become small collectives over that group:

* global weight normalisation (``_check_normalize_sample_weight``, dask/cluster/kmeans.py:141-147),
* one ``random_state`` for all ranks (``:175-177``),
* the global row-count check (``:182-192``) and the all-worker preflight (``_func_preflight_fit``,
  ``:97-113``) -- a rank-local failure is raised on *every* rank, so nobody is left waiting in NCCL,
* fit through ``KMeansMG`` on the handle that carries the NCCL communicator (``_func_fit``, ``:115-134``),
* ``inertia_`` = sum of the rank-local inertias (``:237-243``); ``labels_`` stays sharded: each rank keeps the
  labels of its own rows (the reference keeps them as a distributed Dask array, ``:245-262``),
* ``predict`` / ``transform`` are embarrassingly parallel on a single-GPU model holding the shared centres
  (``DelayedPredictionMixin``); ``score`` sums the per-rank scores (``:341-366``).

The arithmetic is entirely in ``libcuml_b200`` (through ``KMeansMG`` / ``KMeans``); this module only
orchestrates.  Collectives on host scalars use the group's backend device (CUDA tensors under NCCL, CPU
tensors under gloo).
"""
from __future__ import annotations

from numbers import Integral

import numpy as np

from ..cluster.kmeans import KMeans as _SingleGPUKMeans
from ..cluster.kmeans import check_random_seed
from ..cluster.kmeans_mg import KMeansMG, comms_from_torch_distributed


def _validate_n_clusters(n_clusters):
    # reference dask/cluster/kmeans.py:29-33
    if not isinstance(n_clusters, Integral) or n_clusters <= 0:
        raise ValueError(f"n_clusters={n_clusters} should be a positive integer.")


def _n_rows(parts):
    return int(sum(int(p.shape[0]) for p in parts))


def _weight_sum(w):
    if hasattr(w, "detach"):           # torch tensor (host or device)
        return float(w.detach().double().sum().item())
    return float(np.asarray(w, dtype=np.float64).sum())


def _scaled(w, scale):
    # a scaled copy (the reference scales its Dask collection lazily; the caller's array is not modified)
    if hasattr(w, "detach"):
        return w * scale
    return np.asarray(w) * scale


class KMeans:
    """Multi-GPU KMeans: collective ``fit`` over a ``torch.distributed`` group, one process per GPU.

    Parameters are the reference's (``n_clusters, max_iter, tol, verbose, random_state, init,
    oversampling_factor, max_samples_per_batch`` ..., dask/cluster/kmeans.py:48-93); ``group`` replaces the
    Dask ``client`` and ``handle`` may carry an existing communicator (otherwise one is created on first fit
    and owned by this object).
    """

    # the estimator classes are class attributes so that the CPU (gloo) tests of the orchestration can
    # substitute stand-ins; the product always runs the two below
    _mg_class = KMeansMG
    _sg_class = _SingleGPUKMeans

    def __init__(self, *, group=None, handle=None, verbose=False, n_clusters=8, **kwargs):
        self.group = group
        self.handle = handle
        self._owns_handle = False
        self.kwargs = dict(kwargs, verbose=verbose, n_clusters=n_clusters)
        self._local_model = None

    # ---- group helpers ------------------------------------------------------------------------------
    def _dist(self):
        import torch.distributed as dist
        if not dist.is_available() or not dist.is_initialized():
            raise RuntimeError("cuml_b200.distributed.KMeans needs an initialised torch.distributed process group "
                               "(one process per GPU)")
        return dist

    def _rank_world(self):
        dist = self._dist()
        return dist.get_rank(self.group), dist.get_world_size(self.group)

    def _coll_device(self):
        import torch
        dist = self._dist()
        if "nccl" in str(dist.get_backend(self.group)):
            return torch.device("cuda", torch.cuda.current_device())
        return torch.device("cpu")

    def _allreduce_sum(self, values):
        """element-wise sum over the ranks of a short list of host scalars (fp64)"""
        import torch
        dist = self._dist()
        t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=self._coll_device())
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return [float(v) for v in t.cpu().tolist()]

    def _raise_collectively(self, local_error):
        """every rank passes its local exception (or None); if any rank failed, all ranks raise"""
        dist = self._dist()
        rank, world = self._rank_world()
        box = [None] * world
        msg = None if local_error is None else (type(local_error).__name__, str(local_error))
        dist.all_gather_object(box, msg, group=self.group)
        for r, m in enumerate(box):
            if m is not None:
                if r == rank:
                    raise local_error
                exc = ValueError if m[0] == "ValueError" else RuntimeError
                raise exc(f"[rank {r}] {m[1]}")

    def _get_handle(self):
        if self.handle is None:
            self.handle = comms_from_torch_distributed()
            self._owns_handle = True
        return self.handle

    def close(self):
        if self._owns_handle and self.handle is not None:
            self.handle.close()
        self.handle = None
        self._owns_handle = False

    # ---- fit --------------------------------------------------------------------------------------------
    @staticmethod
    def _as_parts(X):
        return list(X) if isinstance(X, (list, tuple)) else [X]

    def _check_weight_parts(self, parts, w_parts):
        """collective: weights must be given on every rank or on none, one per data partition"""
        dist = self._dist()
        _, world = self._rank_world()
        flags = [None] * world
        dist.all_gather_object(flags, w_parts is not None, group=self.group)
        err = None
        if len(set(flags)) != 1:
            err = ValueError("sample_weight must be passed on every rank or on none")
        elif w_parts is not None and len(w_parts) != len(parts):
            err = ValueError("sample_weight partitions must match the data partitions")
        self._raise_collectively(err)

    def _normalized_weights(self, w_parts, n_local):
        """reference ``_check_normalize_sample_weight`` over the *global* array: sum(w) == n_samples"""
        if w_parts is None:
            return None
        ws_local = sum(_weight_sum(w) for w in w_parts)
        n_global, ws_global = self._allreduce_sum([n_local, ws_local])
        if not ws_global > 0.0:
            raise ValueError("sample weights must have a positive sum")
        scale = n_global / ws_global
        return [_scaled(w, scale) for w in w_parts]

    def fit(self, X, sample_weight=None):
        """Collective: every rank passes its local partition(s) of the rows (and of the weights)."""
        rank, world = self._rank_world()
        parts = self._as_parts(X)
        w_parts = None if sample_weight is None else self._as_parts(sample_weight)
        self._check_weight_parts(parts, w_parts)
        n_local = _n_rows(parts)
        w_parts = self._normalized_weights(w_parts, n_local)

        # one random_state for all ranks (reference :175-177): rank 0 decides
        kwargs = dict(self.kwargs)
        box = [check_random_seed(kwargs.get("random_state")) if rank == 0 else None]
        self._dist().broadcast_object_list(box, src=self._dist().get_global_rank(self.group, 0)
                                           if self.group is not None else 0, group=self.group)
        kwargs["random_state"] = box[0]

        # predictable global failures first (reference :179-192); identical on all ranks by construction
        n_clusters = kwargs["n_clusters"]
        _validate_n_clusters(n_clusters)
        total_rows = int(round(self._allreduce_sum([n_local])[0]))
        if total_rows < n_clusters:
            raise ValueError(
                f"n_samples={total_rows} should be >= n_clusters={n_clusters}. "
                f"There are fewer data points across all workers than the "
                f"number of requested clusters. Please reduce n_clusters or "
                f"increase the number of data points.")

        # all-worker preflight (reference _func_preflight_fit): MG-specific parameter and rank-local row checks
        err = None
        model = None
        try:
            model = self._mg_class(handle=self._get_handle(), **kwargs)
            model._validate_fit_params()
            model.validate(parts, rank, world)
        except (ValueError, RuntimeError) as e:
            err = e
        self._raise_collectively(err)

        model.fit(parts, sample_weight=w_parts)

        # the shared centres feed a local single-GPU model for predict / transform / score
        self._set_internal_model(model, kwargs)
        self.inertia_ = self._allreduce_sum([model.inertia_])[0]
        self.labels_ = model.labels_             # this rank's rows, in partition order
        self.n_iter_ = model.n_iter_
        self.n_features_in_ = model.n_features_in_
        return self

    def _set_internal_model(self, mg_model, kwargs):
        local = self._sg_class(**kwargs)
        for name in ("_centers", "_labels", "_in_kind", "inertia_", "n_iter_", "n_features_in_"):
            if hasattr(mg_model, name):
                setattr(local, name, getattr(mg_model, name))
        self._local_model = local

    def _check_is_fitted(self):
        if self._local_model is None:
            raise RuntimeError("This KMeans instance is not fitted yet. Call 'fit' first.")

    @property
    def cluster_centers_(self):
        self._check_is_fitted()
        return self._local_model.cluster_centers_

    # ---- inference: rank-local, no collective except the final sum of ``score`` ------------------------------
    def fit_predict(self, X, sample_weight=None):
        self.fit(X, sample_weight=sample_weight)
        return self.labels_

    def _per_part(self, X, fn):
        out = [fn(p) for p in self._as_parts(X)]
        return out if isinstance(X, (list, tuple)) else out[0]

    def predict(self, X):
        self._check_is_fitted()
        return self._per_part(X, self._local_model.predict)

    def transform(self, X):
        self._check_is_fitted()
        return self._per_part(X, self._local_model.transform)

    def fit_transform(self, X, sample_weight=None):
        return self.fit(X, sample_weight=sample_weight).transform(X)

    def score(self, X, sample_weight=None):
        """Collective.  As in the reference (:341-366) the weights are normalised over the global array first and
        every partition is then scored by the single-GPU model."""
        self._check_is_fitted()
        parts = self._as_parts(X)
        w_parts = None if sample_weight is None else self._as_parts(sample_weight)
        self._check_weight_parts(parts, w_parts)
        w_parts = self._normalized_weights(w_parts, _n_rows(parts))
        local = 0.0
        for i, p in enumerate(parts):
            if int(p.shape[0]) == 0:
                continue
            local += self._local_model.score(p, sample_weight=None if w_parts is None else w_parts[i])
        return self._allreduce_sum([local])[0]

    def get_params(self, deep=True):
        return dict(self.kwargs)

    def set_params(self, **params):
        self.kwargs.update(params)
        return self

    def _get_param_names(self):
        return list(self.kwargs.keys())
