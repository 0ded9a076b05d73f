"""scikit-learn-facing proxy of the k-means path."""
from .cluster import KMeans, UnsupportedOnGPU, install, uninstall  # noqa: F401
