"""Drop-in subset of ``mtflearn.features`` for the Zernike hot path (features/__init__.py:1-35
of the reference): same names, same call signatures, CUDA kernels underneath."""
from ._zps import ZPs, release_plans
from ._zmoments import zmoments
from ._indexing import (construct_complex_matrix, construct_real_matrix, construct_rot_maps_matrix,
                        nm2j, nm2j_complex)
from ._keypoint import KeyPoints, clear_border
from ._local_max import local_max
from ._dimension_reduction import pca
from ._series import series_features

__all__ = ["ZPs", "zmoments", "construct_rot_maps_matrix", "construct_complex_matrix", "construct_real_matrix",
           "KeyPoints", "clear_border", "local_max", "pca", "series_features", "release_plans", "nm2j", "nm2j_complex"]
