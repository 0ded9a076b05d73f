from .kmeans import KMeans  # noqa: F401
from .kmeans_mg import KMeansMG  # noqa: F401
from .spectral_labels import assign_labels_kmeans  # noqa: F401
