"""Drop-in subset of ``mtflearn.clustering`` for the consumers of the Zernike features ("next" row f4)."""
from ._clustering_functions import gmm_lbs, kmeans_lbs, sort_lbs

__all__ = ["kmeans_lbs", "gmm_lbs", "sort_lbs"]
