from .kmeans import KMeans  # noqa: F401
from .kmeans_mg import KMeansMG  # noqa: F401
