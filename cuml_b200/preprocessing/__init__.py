"""Callers of the k-means path inside the reference's preprocessing package."""
from .discretization import kmeans_bin_edges  # noqa: F401
