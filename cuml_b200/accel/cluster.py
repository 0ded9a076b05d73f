"""``KMeans`` with scikit-learn's constructor, methods and fitted attributes, computed by the B200 engine -- the
role of the reference's ``cuml.accel`` proxy for this one estimator (reference
python/cuml/cuml/accel/_overrides/sklearn/cluster.py:12-21 over accel/estimator_proxy.py:170-735), so that
unmodified scikit-learn code (``clone``, ``Pipeline``, ``GridSearchCV``, ``check_is_fitted``) drives the engine.

Hyper-parameters are translated exactly as the reference translates them (``KMeans._params_from_cpu``,
reference kmeans.pyx:621-643): ``init='k-means++'`` becomes the scalable k-means++ of the GPU path,
``'random'`` and array inits pass through, a callable init is unsupported; ``algorithm`` and ``copy_x`` have no
GPU meaning and are only kept so that ``get_params`` / ``clone`` round-trip.  Fitted attributes come back as host
numpy arrays (reference ``_attrs_to_cpu``, kmeans.pyx:660-682).

One deliberate difference: where the reference's proxy falls back to scikit-learn on the CPU for unsupported
hyper-parameters, this one raises ``UnsupportedOnGPU`` -- the product has no CPU path.
"""
from __future__ import annotations

import numpy as np
from sklearn.base import BaseEstimator, ClusterMixin, TransformerMixin

from ..cluster.kmeans import KMeans as _EngineKMeans


class UnsupportedOnGPU(ValueError):
    """hyper-parameters the engine does not implement (reference internals/interop.py ``UnsupportedOnGPU``)"""


class KMeans(ClusterMixin, TransformerMixin, BaseEstimator):
    """Drop-in for ``sklearn.cluster.KMeans`` (same signature as scikit-learn 1.x)."""

    _engine_class = _EngineKMeans      # class attribute so that the CPU suite can substitute the C-ABI stand-in

    def __init__(self, n_clusters=8, *, init="k-means++", n_init="auto", max_iter=300, tol=1e-4, verbose=0,
                 random_state=None, copy_x=True, algorithm="lloyd"):
        self.n_clusters = n_clusters
        self.init = init
        self.n_init = n_init
        self.max_iter = max_iter
        self.tol = tol
        self.verbose = verbose
        self.random_state = random_state
        self.copy_x = copy_x
        self.algorithm = algorithm

    # ---- hyper-parameter translation (reference kmeans.pyx:621-643) -------------------------------------------
    def _engine_params(self):
        init = self.init
        if callable(init):
            raise UnsupportedOnGPU(f"`init={init!r}` is not supported")
        if isinstance(init, str):
            if init == "k-means++":
                init = "scalable-k-means++"
            elif init != "random":
                raise UnsupportedOnGPU(f"`init={init!r}` is not supported")
        return dict(n_clusters=self.n_clusters, init=init, n_init=self.n_init, max_iter=self.max_iter, tol=self.tol,
                    random_state=self.random_state, verbose=bool(self.verbose), output_type="numpy")

    def _fitted_engine(self):
        eng = getattr(self, "_engine", None)
        if eng is None:
            from sklearn.exceptions import NotFittedError
            raise NotFittedError(f"This {type(self).__name__} instance is not fitted yet. Call 'fit' with "
                                 f"appropriate arguments before using this estimator.")
        return eng

    # ---- estimator API -------------------------------------------------------------------------------------------
    def fit(self, X, y=None, sample_weight=None):
        eng = self._engine_class(**self._engine_params())
        eng.fit(X, sample_weight=sample_weight)
        self._engine = eng
        # host copies, as the reference hands them to scikit-learn code (kmeans.pyx:660-682)
        self.cluster_centers_ = np.asarray(eng.cluster_centers_)
        self.labels_ = np.asarray(eng.labels_)
        self.inertia_ = float(eng.inertia_)
        self.n_iter_ = int(eng.n_iter_)
        self.n_features_in_ = int(eng.n_features_in_)
        cols = getattr(X, "columns", None)
        if cols is not None and all(isinstance(c, str) for c in cols):   # DataFrame input, as scikit-learn records it
            self.feature_names_in_ = np.asarray(cols, dtype=object)
        return self

    def fit_predict(self, X, y=None, sample_weight=None):
        return self.fit(X, sample_weight=sample_weight).labels_

    def predict(self, X):
        return np.asarray(self._fitted_engine().predict(X))

    def transform(self, X):
        return np.asarray(self._fitted_engine().transform(X))

    def fit_transform(self, X, y=None, sample_weight=None):
        return self.fit(X, sample_weight=sample_weight).transform(X)

    def score(self, X, y=None, sample_weight=None):
        return float(self._fitted_engine().score(X, sample_weight=sample_weight))

    def get_feature_names_out(self, input_features=None):
        self._fitted_engine()
        return np.asarray([f"kmeans{i}" for i in range(int(self.n_clusters))], dtype=object)

    @property
    def _n_features_out(self):
        return int(self.n_clusters)

    def __sklearn_is_fitted__(self):
        return getattr(self, "_engine", None) is not None

    def as_sklearn(self):
        """a fitted ``sklearn.cluster.KMeans`` carrying this model (reference ``Base.as_sklearn``)"""
        return self._fitted_engine().as_sklearn()


_saved = {}


def install():
    """Make ``sklearn.cluster.KMeans`` resolve to the proxy (what ``cuml.accel.install()`` does for this estimator,
    reference accel/core.py; an attribute swap instead of the reference's import hooks).  Returns the proxy class."""
    import sklearn.cluster
    if "KMeans" not in _saved:
        _saved["KMeans"] = sklearn.cluster.KMeans
        sklearn.cluster.KMeans = KMeans
    return KMeans


def uninstall():
    import sklearn.cluster
    if "KMeans" in _saved:
        sklearn.cluster.KMeans = _saved.pop("KMeans")
