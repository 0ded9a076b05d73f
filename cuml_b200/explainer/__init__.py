"""Callers of the k-means path inside the reference's explainer package."""
from .sampling import kmeans_sampling  # noqa: F401
