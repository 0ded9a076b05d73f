"""One-process-per-GPU orchestration of the k-means path (the ``cuml.dask.cluster`` role)."""
from .kmeans import KMeans  # noqa: F401
