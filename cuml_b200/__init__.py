"""cuml_b200: a B200-native k-means engine behind cuML's operator surface.

Only the k-means hot path exists here: ``cuml_b200.cluster.KMeans`` /
``KMeansMG`` mirror ``cuml.cluster.KMeans`` / ``cuml.cluster.kmeans_mg.KMeansMG`` and call the
C-ABI library ``cuml_b200/lib/libcuml_b200.so`` (hand-written CUDA for sm_100a).  There is no CPU
fallback: without the library or without a CUDA device every compute call raises.
"""
__version__ = "0.1.0"
