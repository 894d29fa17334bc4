"""``kmeans_lbs`` / ``gmm_lbs`` -- cluster labels of the feature matrix on the GPU.

Host-side mirror of ``mtflearn.clustering`` (mtflearn/clustering/_clustering_functions.py:8-43): same names, same
arguments (including the reference's ``ramdom_state`` spelling), same label convention (clusters renumbered by
decreasing size with the reference's own ``np.unique`` / ``np.argsort`` lines).  The reference delegates to
scikit-learn -- ``KMeans(n_clusters=n, random_state=0)`` and ``GaussianMixture(n, 'full', random_state=0)`` -- whose
published algorithms are restated here with every pass over the ``(N, d)`` matrix on the device (float64 arithmetic
on the float32 features, ``csrc/zb200_cluster.cu``) and the small parameter updates on the host:

* seeding: greedy k-means++ (``sklearn/cluster/_kmeans.py:_kmeans_plusplus``) with ``2 + int(log k)`` local trials,
  drawn from the SAME ``numpy.random.RandomState`` stream in the same order (one ``choice``, then one ``uniform`` per
  further centre), so that a seeded run picks the centres scikit-learn picks;
* Lloyd iterations (``_kmeans_single_lloyd``): stop on unchanged labels or when the squared centre shift falls below
  ``tol * mean(var(X))``, then one more assignment if the stop was not strict;
* EM with full covariances (``sklearn/mixture/_base.py:fit_predict``, ``_gaussian_mixture.py``): k-means
  initialisation from the same generator, ``reg_covar=1e-6``, ``tol=1e-3`` on the mean log-likelihood, ``max_iter=100``,
  labels = argmax of the weighted log-probabilities of the final parameters.

Labels are integers; parity with the reference is exact on data whose clusters are separated (tests/golden/
clustering.npz), while rounding-level differences in distances may move samples that sit on a cluster boundary.
numpy in -> numpy int64 labels out; a CUDA tensor is consumed in place (float32).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib
from ..features._device import is_torch, np_ptr


def _device_matrix(X):
    torch = _lib.require_cuda()
    if X.ndim != 2:
        raise ValueError("Expected 2D array, got %dD array instead" % X.ndim)
    dev = X if is_torch(X) else torch.from_numpy(np.ascontiguousarray(X, dtype=np.float32))
    return dev.to(device="cuda", dtype=torch.float32).contiguous()


def _stream():
    return C.c_void_p(_lib.current_stream_ptr())


def _moments(dev):
    """Column means and population variances from the float64 Gram kernel (one pass over X)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    n, d = int(dev.shape[0]), int(dev.shape[1])
    gram = torch.empty((d, d), dtype=torch.float64, device=dev.device)
    colsum = torch.empty((d,), dtype=torch.float64, device=dev.device)
    _lib.check(lib.zb200_gram_f32(int(dev.data_ptr()), n, d, int(gram.data_ptr()), int(colsum.data_ptr()), _stream()), "gram")
    mean = colsum.cpu().numpy() / n
    var = np.maximum(np.diagonal(gram.cpu().numpy()) / n - mean * mean, 0.0)
    return mean, var


def _row(dev, index, mean):
    return dev[int(index)].double().cpu().numpy() - mean


def _kmeans_plusplus(dev, mean, n_clusters, rs):
    """Greedy k-means++ on the centred data; returns centres (k, d) in centred coordinates."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    n, d = int(dev.shape[0]), int(dev.shape[1])
    n_local_trials = 2 + int(np.log(n_clusters))
    centres = np.empty((n_clusters, d), dtype=np.float64)
    first = rs.choice(n, p=np.full(n, 1.0) / float(n))
    centres[0] = _row(dev, first, mean)
    closest = torch.empty((1, n), dtype=torch.float64, device=dev.device)
    pot = np.zeros(1)
    _lib.check(lib.zb200_kmeans_mindist_f32(int(dev.data_ptr()), n, d, np_ptr(mean), np_ptr(centres[0:1].copy()), 1, None,
                                            int(closest.data_ptr()), np_ptr(pot), _stream()), "kmeans_mindist")
    closest, current_pot = closest[0], float(pot[0])
    for c in range(1, n_clusters):
        rand_vals = rs.uniform(size=n_local_trials) * current_pot
        cum = np.cumsum(closest.cpu().numpy())
        ids = np.searchsorted(cum, rand_vals)
        np.clip(ids, None, n - 1, out=ids)
        cand = np.ascontiguousarray(np.stack([_row(dev, i, mean) for i in ids]))
        out = torch.empty((len(ids), n), dtype=torch.float64, device=dev.device)
        pots = np.zeros(len(ids))
        _lib.check(lib.zb200_kmeans_mindist_f32(int(dev.data_ptr()), n, d, np_ptr(mean), np_ptr(cand), len(ids),
                                                int(closest.data_ptr()), int(out.data_ptr()), np_ptr(pots), _stream()),
                   "kmeans_mindist")
        best = int(np.argmin(pots))
        current_pot = float(pots[best])
        closest = out[best].contiguous()
        centres[c] = cand[best]
    return centres


def _kmeans_fit(dev, n_clusters, rs, max_iter=300, tol=1e-4):
    """(labels CUDA int32 tensor, centres (k, d) absolute, mean) -- KMeans(n_clusters, n_init=1, random_state=rs)."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    n, d = int(dev.shape[0]), int(dev.shape[1])
    if n_clusters is None or n_clusters < 1 or n_clusters > n:
        raise ValueError(f"n_clusters={n_clusters} must be between 1 and n_samples={n}")
    mean, var = _moments(dev)
    tol_abs = float(np.mean(var) * tol)
    centres = _kmeans_plusplus(dev, mean, n_clusters, rs)
    labels = torch.full((n,), -1, dtype=torch.int32, device=dev.device)
    sums = np.zeros((n_clusters, d))
    counts = np.zeros(n_clusters)
    changed = C.c_int64(0)
    strict = False
    for _ in range(max_iter):
        _lib.check(lib.zb200_kmeans_step_f32(int(dev.data_ptr()), n, d, np_ptr(mean), np_ptr(centres), n_clusters,
                                             int(labels.data_ptr()), 1, np_ptr(sums), np_ptr(counts), C.byref(changed),
                                             _stream()), "kmeans_step")
        if np.any(counts == 0):
            raise RuntimeError("k-means produced an empty cluster; scikit-learn's relocation of empty clusters is not implemented")
        new = sums / counts[:, None]
        shift_tot = float(((new - centres) ** 2).sum())
        centres = np.ascontiguousarray(new)
        if changed.value == 0:
            strict = True
            break
        if shift_tot <= tol_abs:
            break
    if not strict:                                     # labels consistent with the final centres
        _lib.check(lib.zb200_kmeans_step_f32(int(dev.data_ptr()), n, d, np_ptr(mean), np_ptr(centres), n_clusters,
                                             int(labels.data_ptr()), 0, None, None, C.byref(changed), _stream()), "kmeans_step")
    return labels, centres + mean, mean


def _reorder_by_count(lbs):
    # the reference's own lines (_clustering_functions.py:18-22 and :28-32)
    unique, counts = np.unique(lbs, return_counts=True)
    lbs_order = np.argsort(counts)[::-1]
    order_dict = dict(zip(lbs_order, unique))
    return np.vectorize(order_dict.get)(lbs)


def kmeans_lbs(X, n=None, random_state=0):
    dev = _device_matrix(X)
    rs = random_state if isinstance(random_state, np.random.RandomState) else np.random.RandomState(random_state)
    labels, _, _ = _kmeans_fit(dev, n, rs)
    return _reorder_by_count(labels.cpu().numpy().astype(np.int64))


def _gaussian_parameters(nk, sx, sxx, mean, reg_covar):
    """means (absolute), covariances, precision Cholesky factors, log-determinants from the centred accumulators."""
    from scipy import linalg
    k, d = sx.shape
    nk = nk + 10 * np.finfo(np.float64).eps
    mu_c = sx / nk[:, None]
    cov = sxx / nk[:, None, None] - mu_c[:, :, None] * mu_c[:, None, :]
    cov = 0.5 * (cov + np.transpose(cov, (0, 2, 1)))
    cov[:, np.arange(d), np.arange(d)] += reg_covar
    prec = np.empty_like(cov)
    for c in range(k):
        try:
            chol = linalg.cholesky(cov[c], lower=True)
        except linalg.LinAlgError:
            raise ValueError("Fitting the mixture model failed because some components have ill-defined empirical "
                             "covariance (for instance caused by singleton or collapsed samples). Try to decrease the "
                             "number of components, increase reg_covar, or scale the input data.")
        prec[c] = linalg.solve_triangular(chol, np.eye(d), lower=True).T
    log_det = np.log(np.diagonal(prec, axis1=1, axis2=2)).sum(axis=1)
    return nk, mu_c + mean, np.ascontiguousarray(prec), log_det


def gmm_lbs(X, n, type='full', ramdom_state=0):
    if type != 'full':
        raise NotImplementedError("only covariance_type='full' (the reference's default) runs on the GPU")
    torch = _lib.require_cuda()
    lib = _lib.load()
    dev = _device_matrix(X)
    n_samples, d = int(dev.shape[0]), int(dev.shape[1])
    rs = ramdom_state if isinstance(ramdom_state, np.random.RandomState) else np.random.RandomState(ramdom_state)
    reg_covar, tol, max_iter = 1e-6, 1e-3, 100
    labels, _, mean = _kmeans_fit(dev, n, rs)          # init_params='kmeans': one-hot responsibilities
    nk, sx, sxx = np.zeros(n), np.zeros((n, d)), np.zeros((n, d, d))

    def m_step(log_resp, lab):
        _lib.check(lib.zb200_gmm_mstep_f32(int(dev.data_ptr()), n_samples, d, np_ptr(mean),
                                           None if log_resp is None else int(log_resp.data_ptr()),
                                           None if lab is None else int(lab.data_ptr()), n, np_ptr(nk), np_ptr(sx), np_ptr(sxx),
                                           _stream()), "gmm_mstep")
        return _gaussian_parameters(nk.copy(), sx, sxx, mean, reg_covar)

    nk_, means, prec, log_det = m_step(None, labels)
    weights = nk_ / n_samples
    log_resp = torch.empty((n_samples, n), dtype=torch.float64, device=dev.device)
    lower = C.c_double(0.0)
    lower_bound = -np.inf
    for _ in range(max_iter):
        prev = lower_bound
        logw = np.ascontiguousarray(np.log(weights))
        _lib.check(lib.zb200_gmm_estep_f32(int(dev.data_ptr()), n_samples, d, np_ptr(logw), np_ptr(np.ascontiguousarray(log_det)),
                                           np_ptr(np.ascontiguousarray(means)), np_ptr(prec), n, int(log_resp.data_ptr()), None,
                                           C.byref(lower), _stream()), "gmm_estep")
        nk_, means, prec, log_det = m_step(log_resp, None)
        weights = nk_ / nk_.sum()
        lower_bound = lower.value
        if abs(lower_bound - prev) < tol:
            break
    out = torch.empty((n_samples,), dtype=torch.int32, device=dev.device)
    logw = np.ascontiguousarray(np.log(weights))
    _lib.check(lib.zb200_gmm_estep_f32(int(dev.data_ptr()), n_samples, d, np_ptr(logw), np_ptr(np.ascontiguousarray(log_det)),
                                       np_ptr(np.ascontiguousarray(means)), np_ptr(prec), n, None, int(out.data_ptr()), None,
                                       _stream()), "gmm_estep")
    return _reorder_by_count(out.cpu().numpy().astype(np.int64))


def sort_lbs(lbs):
    unique_lbs, counts = np.unique(lbs, return_counts=True)
    idx = np.argsort(counts)[::-1]
    unique_lbs = unique_lbs[idx]
    lbs_order = range(len(unique_lbs))
    order_dict = dict(zip(unique_lbs, lbs_order))
    lbs_ = np.vectorize(order_dict.get)(lbs)
    return lbs_
