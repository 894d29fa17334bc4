"""``pca`` -- principal components of the feature matrix on the GPU.

Host-side mirror of ``mtflearn.features.pca`` (mtflearn/features/_dimension_reduction.py:3-6):
``sklearn.decomposition.PCA(n_components).fit_transform(X)``.  For ``n_samples >= 10 n_features``
scikit-learn (>= 1.5) takes the covariance route, which is what runs here: the two passes over the
``(N, M)`` matrix -- ``X^T X`` with column sums, and the scores ``(X - mean) V^T`` -- are float64 CUDA
kernels (``zb200_gram_f32`` / ``zb200_pca_scores_f32``); the ``M x M`` symmetric eigenproblem and the
``svd_flip(u_based_decision=False)`` sign rule are host glue, as they are inside scikit-learn.
numpy in -> float64 numpy out; CUDA tensor in -> float64 CUDA tensor out.  The matrix is consumed as
float32 (the dtype the feature kernels produce).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib
from ._device import is_torch, np_ptr


def pca_fit(X, n_components=2):
    """(mean[M], components[n_components, M], explained_variance[n_components]) of the rows of X."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    if X.ndim != 2:
        raise ValueError("Expected 2D array, got %dD array instead" % X.ndim)
    dev = X if is_torch(X) else torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32))
    dev = dev.to(device="cuda", dtype=torch.float32).contiguous()
    n, m = int(dev.shape[0]), int(dev.shape[1])
    if not 1 <= n_components <= min(n, m):
        raise ValueError(f"n_components={n_components} must be between 1 and min(n_samples, n_features)={min(n, m)}")
    gram = torch.empty((m, m), dtype=torch.float64, device=dev.device)
    colsum = torch.empty((m,), dtype=torch.float64, device=dev.device)
    stream = C.c_void_p(_lib.current_stream_ptr())
    _lib.check(lib.zb200_gram_f32(int(dev.data_ptr()), n, m, int(gram.data_ptr()), int(colsum.data_ptr()), stream), "gram")
    g = gram.cpu().numpy()
    mean = colsum.cpu().numpy() / n
    cov = (g - n * np.outer(mean, mean)) / (n - 1)
    cov = (cov + cov.T) * 0.5
    evals, evecs = np.linalg.eigh(cov)
    evals, evecs = evals[::-1], evecs[:, ::-1]
    evals = np.where(evals < 0.0, 0.0, evals)
    vt = np.ascontiguousarray(evecs.T[:n_components])
    # svd_flip(u=None, v, u_based_decision=False): the entry of largest magnitude of every component is positive
    idx = np.argmax(np.abs(vt), axis=1)
    signs = np.sign(vt[np.arange(vt.shape[0]), idx])
    signs[signs == 0] = 1.0
    vt *= signs[:, None]
    return dev, mean, vt, evals[:n_components]


def pca(X, n_components=2, reconstruct=False):
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev, mean, vt, _ = pca_fit(X, n_components)
    n, m = int(dev.shape[0]), int(dev.shape[1])
    out = torch.empty((n, vt.shape[0]), dtype=torch.float64, device=dev.device)
    mean = np.ascontiguousarray(mean, dtype=np.float64)
    _lib.check(lib.zb200_pca_scores_f32(int(dev.data_ptr()), n, m, np_ptr(mean), np_ptr(vt), int(vt.shape[0]),
                                        int(out.data_ptr()), C.c_void_p(_lib.current_stream_ptr())), "pca_scores")
    return out if (is_torch(X) and X.is_cuda) else out.cpu().numpy()
