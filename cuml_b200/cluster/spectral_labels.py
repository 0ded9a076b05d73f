"""Label assignment of spectral clustering -- the last step of the reference's ``SpectralClustering.fit`` /
``spectral_clustering`` (reference python/cuml/cuml/cluster/spectral_clustering.pyx:113 "Kmeans is used for
assigning labels", config ``n_clusters`` / ``n_init`` / ``seed`` ``:337-349``), a caller of the k-means path
(SURVEY.md 8f-4).

Only that call is mirrored: the k-NN graph, the normalised Laplacian and the Lanczos eigensolver that produce the
embedding are outside SURVEY.md section 8.  Given the ``(n_samples, n_components)`` embedding the reference runs
k-means with ``n_init`` seeded restarts and returns the labels of the best run (lowest inertia) as int32.
"""
from __future__ import annotations


def assign_labels_kmeans(embedding, n_clusters=8, n_init=10, random_state=None, _estimator=None):
    """int32 labels of the rows of ``embedding`` (array-like, (n_samples, n_components)): k-means with
    ``n_init`` restarts from the seeded scalable k-means++ init, best inertia wins.  numpy in -> numpy out;
    torch / ``__cuda_array_interface__`` in -> torch CUDA tensor out."""
    from .kmeans import KMeans
    km = (_estimator or KMeans)(n_clusters=n_clusters, n_init=n_init, random_state=random_state)
    km.fit(embedding)
    return km.labels_
