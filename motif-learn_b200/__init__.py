"""motif_learn_b200 -- B200-native (sm_100a) Zernike hot path of jiadongdan/motif-learn.

Covers ONE path of the reference: ``ZPs`` (patch-moment transform + dense moment map) and
``zmoments`` (real/complex moments, normalisation, n-fold symmetry scores), plus the
``KeyPoints`` patch gather that feeds it.  Same Python API as ``mtflearn.features``; hand-
written CUDA kernels behind a C ABI (include/zernike_b200.h); no CPU fallback.
"""
__version__ = "0.1.0"

from . import features            # noqa: F401
from . import clustering          # noqa: F401
from . import denoise             # noqa: F401
from .features import ZPs, zmoments, KeyPoints  # noqa: F401

__all__ = ["features", "clustering", "denoise", "ZPs", "zmoments", "KeyPoints", "__version__"]
