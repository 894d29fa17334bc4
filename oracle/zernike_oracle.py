"""CPU oracle for the Zernike hot path -- TEST INFRASTRUCTURE, not product code.

A float64 numpy/scipy restatement of the reference algorithm of
``mtflearn.features`` (jiadongdan/motif-learn 0.1.2) for the path named in
BASELINE.json's north_star.  Every function cites the reference file:line it
follows.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
cpu-baseline / ``--impl reference`` leg may import this module; the product
package (``motif_learn_b200``) never does and has no CPU fallback.

Parity status: the reference's own tests pin only ``nm2j`` integers, validation
messages and ``select`` ordering (tests/features/test_zmoments.py:5-88) -- they
hold no numeric vector for ZPs/transform/to_complex/rot_maps.  The oracle is
therefore pinned against (i) those reference tests, restated in
``tests/test_oracle_golden.py``, and (ii) outputs of the REAL reference run in
the build container through ``oracle/ref_shim.py``; the vectors are committed as
``tests/golden/*.npz`` together with the generating script
``oracle/make_goldens.py``.

Third-party arithmetic the reference delegates to (not under /root/reference,
versions unpinned in pyproject.toml:14-22; this image: numpy 2.3.5, scipy
1.18.1): ``numpy.dot`` (OpenBLAS dgemm), ``scipy.signal.fftconvolve``
(pocketfft), ``scipy.special.factorial``, ``numpy.linalg.norm``.  The oracle
calls the same library entry points for the same steps so that the
``cpu_baseline`` it provides is the reference's own CPU algorithm.

"Next" rows: ``render_atoms`` and ``pca`` are pinned against the live reference
(tests/golden/render.npz, pca.npz).  ``local_max`` is pinned only in part: its
suppression step is the reference's own ``filter_peaks_by_distance`` run live
(tests/golden/peaks.npz), but the candidate step restates scikit-image's
``peak_local_max`` from its published algorithm -- scikit-image is absent from
this image and unpinned in the reference: PARITY UNPINNED for that step.
"""
from __future__ import annotations

import math

import numpy as np

__all__ = [
    "nm2j", "nm2j_complex", "mode_table", "zernike_basis", "zernike_basis_exact",
    "project_patches", "moment_map_fft", "moment_map_direct", "valid_mask",
    "complex_matrix", "to_complex", "to_real", "normalize", "select_indices",
    "rotate", "rot_weights", "rot_maps", "mirror_map", "clear_border",
    "extract_patches", "render_atoms", "peak_local_max_md1", "filter_peaks_by_distance", "local_max", "pca",
]


# --------------------------------------------------------------------------- #
# index maps                                                                   #
# --------------------------------------------------------------------------- #
def nm2j(n, m):
    """Single index of a real mode; follows _zmoments.py:3-69 (j=((n+2)n+m)//2)."""
    n = np.asarray(n)
    m = np.asarray(m)
    if n.shape != m.shape:
        raise ValueError("`n` and `m` must have the same shape.")
    if not np.all(np.isclose(n % 1, 0)):
        raise ValueError("Radial order `n` must be integer-valued.")
    if not np.all(np.isclose(m % 1, 0)):
        raise ValueError("Azimuthal frequency `m` must be integer-valued.")
    n = n.astype(int)
    m = m.astype(int)
    if np.any(n < 0):
        raise ValueError("Radial order `n` must be non-negative.")
    if np.any(np.abs(m) > n):
        raise ValueError("Azimuthal frequency `m` must satisfy |m| ≤ n.")
    if np.any((n - np.abs(m)) % 2 != 0):
        raise ValueError("`n - |m|` must be even.")
    j = ((n + 2) * n + m) // 2
    return j.item() if j.shape == () else j


def nm2j_complex(n, m):
    """Index of a complex (m>=0) mode; follows _zmoments.py:71-91."""
    n = np.atleast_1d(n)
    m = np.atleast_1d(m)
    if not np.all(n >= 0):
        raise ValueError("Radial order n must be non-negative.")
    if not np.all(m >= 0):
        raise ValueError("Azimuthal frequency m must be non-negative.")
    if not np.all(np.abs(m) <= n):
        raise ValueError("Azimuthal frequency m must satisfy |m| ≤ n.")
    if not np.all((n - np.abs(m)) % 2 == 0):
        raise ValueError("n - |m| must be even.")
    raw = n ** 2 + 2 * n + 2 * m
    out = np.where(n % 2 == 0, raw // 4, (raw - 1) // 4)
    return out.item() if out.size == 1 else out


def mode_table(n_max: int):
    """(n, m) in ZPs order: n ascending, m=-n,-n+2,..,n (_zps.py:77-80)."""
    ns, ms = [], []
    for n in range(n_max + 1):
        for m in range(-n, n + 1, 2):
            ns.append(n)
            ms.append(m)
    return np.array(ns), np.array(ms)


# --------------------------------------------------------------------------- #
# basis                                                                        #
# --------------------------------------------------------------------------- #
def _grid(size: int):
    """rho, theta on linspace(-1,1,size)^2 -- _zps.py:68-72 (xv=column, yv=row)."""
    ax = np.linspace(-1, 1, size)
    xv, yv = np.meshgrid(ax, ax)
    return np.sqrt(xv ** 2 + yv ** 2), np.arctan2(yv, xv)


def _radial_factorial(n: int, am: int, rho: np.ndarray) -> np.ndarray:
    """Power-sum radial polynomial with float factorials -- _zps.py:52-64."""
    from scipy.special import factorial
    acc = np.zeros_like(rho)
    for s in range((n - am) // 2 + 1):
        c = ((-1) ** s * factorial(n - s)) / (
            factorial(s) * factorial((n + am) // 2 - s) * factorial((n - am) // 2 - s))
        acc += c * rho ** (n - 2 * s)
    return acc


def zernike_basis(n_max: int, size: int):
    """(n, m, V[M,size,size]) exactly as the reference builds it (_zps.py:66-90).

    V = [rho<=1] * R_n^|m|(rho) * sqrt(2(n+1)/(1+[m==0])) * (sin(|m|theta) if m<0
    else cos(m theta)).  Same float-factorial power sum, so it agrees with the
    reference to the last bits (and inherits its n_max>~24 inaccuracy).
    """
    rho, theta = _grid(size)
    inside = rho <= 1
    n_arr, m_arr = mode_table(n_max)
    planes = np.empty((len(n_arr), size, size), dtype=np.float64)
    for j, (n, m) in enumerate(zip(n_arr.tolist(), m_arr.tolist())):
        rad = _radial_factorial(n, abs(m), rho) * np.sqrt(2 * (n + 1) / (1 + (m == 0)))
        rad = np.where(inside, rad, 0)
        planes[j] = rad * (np.sin(-m * theta) if m < 0 else np.cos(m * theta))
    return n_arr, m_arr, planes


def zernike_basis_exact(n_max: int, size: int, points):
    """V[:, y, x] at the given (row, col) points with the radial polynomial
    evaluated EXACTLY (integer coefficients, rational arithmetic on the float64
    rho, one final rounding).  Not the reference algorithm: an independent
    yardstick that shows which of {reference power sum, GPU recurrence} is closer
    to the true polynomial.  Slow -- sampled points only."""
    from fractions import Fraction
    rho, theta = _grid(size)
    n_arr, m_arr = mode_table(n_max)
    out = np.zeros((len(n_arr), len(points)), dtype=np.float64)
    for col, (r, c) in enumerate(points):
        if not rho[r, c] <= 1:
            continue
        fr = Fraction(float(rho[r, c]))
        for j, (n, m) in enumerate(zip(n_arr.tolist(), m_arr.tolist())):
            am = abs(m)
            half = (n - am) // 2
            tot = Fraction(0)
            for s in range(half + 1):
                coef = ((-1) ** s * math.factorial(n - s)
                        // (math.factorial(s) * math.factorial((n + am) // 2 - s) * math.factorial(half - s)))
                tot += coef * fr ** (n - 2 * s)
            rad = float(tot) * math.sqrt(2 * (n + 1) / (1 + (m == 0)))
            out[j, col] = rad * (math.sin(am * theta[r, c]) if m < 0 else math.cos(am * theta[r, c]))
    return n_arr, m_arr, out


# --------------------------------------------------------------------------- #
# transforms                                                                   #
# --------------------------------------------------------------------------- #
def _area(size: int) -> float:
    return np.pi * size ** 2 / 4          # _zps.py:154,177


def project_patches(patches: np.ndarray, basis: np.ndarray) -> np.ndarray:
    """Z[N,M] = X[N,k^2] @ V[M,k^2]^T / area -- _zps.py:146-157 (numpy.dot)."""
    k = basis.shape[-1]
    if patches.ndim != 3 or patches.shape[1] != k or patches.shape[2] != k:
        raise ValueError(
            f"For batch processing, image size ({patches.shape[1]}x{patches.shape[2]}) must match "
            f"polynomial size ({k}x{k})")
    flat_v = basis.reshape(-1, k * k)
    flat_x = patches.reshape(patches.shape[0], k * k)
    return np.dot(flat_x, flat_v.T) / _area(k)


def moment_map_fft(image: np.ndarray, basis: np.ndarray, n_arr: np.ndarray,
                   chunk: int | None = None) -> np.ndarray:
    """Dense map Z[M,H,W] via scipy.signal.fftconvolve('same') then (-1)^n/area
    -- _zps.py:159-193.  ``chunk`` processes that many modes per FFT call (the
    per-mode transforms are independent) to bound memory."""
    from scipy.signal import fftconvolve
    k = basis.shape[-1]
    h, w = image.shape
    if h < k or w < k:
        raise ValueError(
            f"For FFT convolution, image size ({h}x{w}) must be at least "
            f"as large as polynomial size ({k}x{k})")
    m_modes = basis.shape[0]
    sign = np.where(np.asarray(n_arr) % 2 == 0, 1.0, -1.0)[:, None, None]
    if chunk is None or chunk >= m_modes:
        conv = fftconvolve(np.broadcast_to(image, (m_modes, h, w)), basis, mode="same", axes=[1, 2])
        return sign * conv / _area(k)
    out = np.empty((m_modes, h, w), dtype=np.result_type(image.dtype, basis.dtype))
    for s in range(0, m_modes, chunk):
        e = min(m_modes, s + chunk)
        conv = fftconvolve(np.broadcast_to(image, (e - s, h, w)), basis[s:e], mode="same", axes=[1, 2])
        out[s:e] = sign[s:e] * conv / _area(k)
    return out


def moment_map_direct(image: np.ndarray, basis: np.ndarray) -> np.ndarray:
    """Direct-form definition of the same map:
    Z[j,y,x] = 1/area * sum_{a,b} img0[y-k//2+a, x-k//2+b] V[j,a,b], img0 zero
    extended.  O(HWk^2M): small cases only; an FFT-free cross-check."""
    k = basis.shape[-1]
    h, w = image.shape
    half = k // 2
    pad = np.zeros((h + k, w + k), dtype=np.float64)
    pad[half:half + h, half:half + w] = image
    win = np.lib.stride_tricks.sliding_window_view(pad, (k, k))[:h, :w]   # (H,W,k,k)
    return np.einsum("yxab,jab->jyx", win, basis, optimize=True) / _area(k)


def valid_mask(shape_hw, patch_size: int) -> np.ndarray:
    """_zmoments.py:279-294 -- reproduced as-is incl. the even-k off-by-one."""
    mask = np.ones(shape_hw, dtype=bool)
    before = (patch_size - 1) // 2
    after = patch_size - 1 - before
    mask[:before, :] = False
    mask[-after:, :] = False
    mask[:, :before] = False
    mask[:, -after:] = False
    return mask


# --------------------------------------------------------------------------- #
# zmoments algebra                                                             #
# --------------------------------------------------------------------------- #
def complex_matrix(n_arr, m_arr) -> np.ndarray:
    """0/1/i matrix C with Zc = C @ Z -- _zmoments.py:111-132.
    Row = nm2j_complex(n,|m|) rank, entry 1 for m>=0 and 1j for m<0."""
    order = np.lexsort((m_arr, n_arr))
    n_s, m_s = np.asarray(n_arr)[order], np.asarray(m_arr)[order]
    cj = np.atleast_1d(nm2j_complex(n_s, np.abs(m_s)))
    rows = {v: r for r, v in enumerate(np.unique(cj))}
    mat = np.zeros((len(rows), len(n_s)), dtype=complex)
    for col, (tag, mm) in enumerate(zip(cj.tolist(), m_s.tolist())):
        mat[rows[tag], col] = 1.0 if mm >= 0 else 1j
    return mat


def to_complex(data: np.ndarray, n_arr, m_arr):
    """(Zc, n_c, m_c) -- _zmoments.py:300-316.  Zc[n,m]=Z[n,+m]+i Z[n,-m]."""
    cm = complex_matrix(n_arr, m_arr)
    if data.ndim == 2:
        zc = np.dot(cm, data.T).T
    else:
        zc = np.tensordot(cm, data, axes=([1], [0]))
    pick = np.where(cm == 1j, 0, cm)
    m_c = pick.dot(np.abs(m_arr)).real.astype(int)
    n_c = pick.dot(np.abs(n_arr)).real.astype(int)
    return zc, n_c, m_c


def to_real(zc: np.ndarray, n_c, m_c):
    """(Z, n, m) inverse of to_complex -- _zmoments.py:134-196,318-341."""
    n_r, m_r = [], []
    for nn, mm in zip(np.asarray(n_c).tolist(), np.asarray(m_c).tolist()):
        if mm == 0:
            n_r.append(nn); m_r.append(0)
        else:
            n_r += [nn, nn]; m_r += [mm, -mm]
    n_r, m_r = np.array(n_r), np.array(m_r)
    order = np.lexsort((m_r, n_r))
    n_r, m_r = n_r[order], m_r[order]
    cm = complex_matrix(n_r, m_r)                       # (Mc, M)
    inv = np.where(cm == 1, 1.0, 0) + np.where(cm == 1j, -1j, 0)
    inv = inv.T                                         # (M, Mc)
    if zc.ndim == 2:
        z = np.dot(zc, inv.T).real
    else:
        z = np.tensordot(inv, zc, axes=([1], [0])).real
    return z, n_r, m_r


def normalize(data: np.ndarray, order=None) -> np.ndarray:
    """Divide by the p-norm over the mode axis -- _zmoments.py:344-356."""
    axis = 1 if data.ndim == 2 else 0
    return data / np.linalg.norm(data, ord=order, axis=axis, keepdims=True)


def select_indices(m_arr, m_select, invert: bool = False) -> np.ndarray:
    """Column indices kept by select / unselect -- _zmoments.py:359-374."""
    m_abs = np.abs(np.asarray(m_arr))
    want = np.unique(np.abs(np.atleast_1d(m_select).ravel()))
    if invert:
        want = np.array([v for v in np.unique(m_abs) if v not in want])
    return np.where(np.isin(m_abs, want))[0]


def rotate(data: np.ndarray, n_arr, m_arr, theta_deg: float):
    """Complex moments times exp(-i m theta) -- _zmoments.py:377-418."""
    zc, n_c, m_c = (data, n_arr, m_arr) if np.iscomplexobj(data) else to_complex(data, n_arr, m_arr)
    fac = np.exp(-1j * np.deg2rad(theta_deg) * m_c)
    if zc.ndim == 3:
        fac = fac[:, None, None]
    return zc * fac, n_c, m_c


def rot_weights(n_folds, m_arr) -> np.ndarray:
    """W[f, j] -- _zmoments.py:199-235: +1 if |m|%f==0 and |m|>1, 0 if |m| in
    {0,1}, else -1/(f-1) (0 when f<=1)."""
    folds = np.atleast_1d(n_folds).ravel()
    m_abs = np.abs(np.atleast_1d(m_arr).ravel())
    w = np.zeros((len(folds), len(m_abs)))
    for r, f in enumerate(folds):
        hit = (m_abs % f == 0) & (m_abs > 1)
        skip = (m_abs == 0) | (m_abs == 1)
        w[r, hit] = 1
        w[r, ~(hit | skip)] = -1.0 / (f - 1) if f > 1 else 0
    return w


def rot_maps(data: np.ndarray, n_arr, m_arr, n_folds, p=2, m_unselect=None) -> np.ndarray:
    """n-fold symmetry scores -- _zmoments.py:420-462."""
    if m_unselect is None:
        m_unselect = (0, 1)
    elif 0 not in m_unselect:
        raise ValueError("m=0 must be included in m_unselect.")
    keep = select_indices(m_arr, m_unselect, invert=True)
    sub = data[:, keep] if data.ndim == 2 else data[keep]
    if p is not None:
        sub = normalize(sub, order=p)
    sq = sub ** 2
    w = rot_weights(n_folds, np.asarray(m_arr)[keep])
    return np.dot(sq, w.T) if data.ndim == 2 else np.tensordot(w, sq, axes=([1], [0]))


def mirror_map(data: np.ndarray, n_arr, m_arr, theta=None, p=2, m_unselect=(0, 1)) -> np.ndarray:
    """max_theta sum Re(Zc^2 e^{-i m theta}) -- _zmoments.py:464-493 ("next" row a17)."""
    if theta is None:
        theta = np.linspace(0, 2 * np.pi, 360, endpoint=False)
    keep = select_indices(m_arr, m_unselect, invert=True)
    sub = data[:, keep] if data.ndim == 2 else data[keep]
    n_k, m_k = np.asarray(n_arr)[keep], np.asarray(m_arr)[keep]
    if p is not None:
        sub = normalize(sub, order=p)
    zc, _, m_c = to_complex(sub, n_k, m_k)
    a, b = zc.real, zc.imag
    ang = np.outer(np.asarray(theta), m_c)                # (T, Mc)
    mat = np.hstack([np.cos(ang), np.sin(ang)])
    if data.ndim == 2:
        return np.dot(np.hstack([a * a - b * b, 2 * a * b]), mat.T).max(axis=1)
    return np.tensordot(mat, np.vstack([a * a - b * b, 2 * a * b]), axes=([1], [0])).max(axis=0)


# --------------------------------------------------------------------------- #
# patch gather                                                                 #
# --------------------------------------------------------------------------- #
def clear_border(pts: np.ndarray, shape_hw, size: int) -> np.ndarray:
    """Strict-inequality border filter -- _keypoint.py:44-48 (pts are (x,y))."""
    x, y = pts[:, 0], pts[:, 1]
    lo = size // 2 + 1
    keep = (x > lo) & (x < shape_hw[1] - size // 2 - 1) & (y > lo) & (y < shape_hw[0] - size // 2 - 1)
    return pts[keep]


def extract_patches(img: np.ndarray, pts: np.ndarray, size: int, flat: bool = False) -> np.ndarray:
    """patches[i] = img[y-k//2 : y-k//2+k, x-k//2 : x-k//2+k], (x,y)=rint(pts[i])
    (half-to-even) -- _keypoint.py:60-78."""
    half = size // 2
    centres = np.rint(pts).astype(int)
    out = np.empty((len(centres), size, size), dtype=img.dtype)
    for i, (x, y) in enumerate(centres):
        out[i] = img[y - half:y - half + size, x - half:x - half + size]
    return out.reshape(len(centres), size * size) if flat else out


# --------------------------------------------------------------------------- #
# synthetic frames ("next" row f1)                                              #
# --------------------------------------------------------------------------- #
def render_atoms(shape_hw, pts: np.ndarray, amps, sigma: float, r_factor: float = 3.0) -> np.ndarray:
    """Sum of tapered Gaussians drawn like the reference's synthetic frames --
    mtflearn/datasets/_tapered_gaussian.py:3-97 (used by HoneyCombLattice.to_image,
    _honeycomb_lattice.py:169-226): value A*exp(-r^2/(2 sigma^2))*(1-3t^2+2t^3), t=r/R, R=r_factor*sigma,
    zero beyond R; evaluated in float64 and accumulated atom by atom into a float32 frame."""
    h, w = shape_hw
    img = np.zeros((h, w), dtype=np.float32)
    amps = np.broadcast_to(np.asarray(amps, dtype=float), (len(pts),))
    cut = r_factor * float(sigma)
    for (x0, y0), a in zip(np.asarray(pts, dtype=float), amps):
        xa, xb = max(int(np.floor(x0 - cut)), 0), min(int(np.ceil(x0 + cut)) + 1, w)
        ya, yb = max(int(np.floor(y0 - cut)), 0), min(int(np.ceil(y0 + cut)) + 1, h)
        if xa >= xb or ya >= yb:
            continue
        xx, yy = np.meshgrid(np.arange(xa, xb), np.arange(ya, yb), indexing="xy")
        r = np.sqrt((xx - x0) ** 2 + (yy - y0) ** 2)
        t = r / cut
        val = np.where(r <= cut, a * np.exp(-0.5 * r ** 2 / float(sigma) ** 2) * (1.0 - 3.0 * t ** 2 + 2.0 * t ** 3), 0.0)
        img[ya:yb, xa:xb] += val
    return img


# --------------------------------------------------------------------------- #
# peak detection ("next" row f2)                                                #
# --------------------------------------------------------------------------- #
def peak_local_max_md1(image: np.ndarray, threshold_abs=None) -> np.ndarray:
    """Restatement of ``skimage.feature.peak_local_max(image, min_distance=1, threshold_abs=t)`` as the
    reference calls it (mtflearn/features/_local_max_v2.py:61).  scikit-image is a third-party dependency
    that is NOT installed in this image and is unpinned in the reference's pyproject.toml, so this follows
    its published algorithm (skimage/feature/peak.py, 0.19-0.25: ``_get_threshold``, ``_get_peak_mask``,
    ``_exclude_border``, ``_get_high_intensity_peaks``):
      threshold = threshold_abs if given else image.min();
      mask = (image == maximum_filter(image, 3x3 footprint, mode='nearest')), all-False for a trivial
      (constant) image, & (image > threshold); the 1-pixel border is cleared; coordinates sorted by
      ``argsort(-intensity, kind='stable')``; ``ensure_spacing(spacing=1, p_norm=inf)`` only rejects points
      strictly closer than 1 pixel, i.e. nothing.  Returns (N, 2) (row, col)."""
    import scipy.ndimage as ndi
    image = np.asarray(image)
    threshold = image.min() if threshold_abs is None else threshold_abs
    out = image == ndi.maximum_filter(image, footprint=np.ones((3, 3), dtype=bool), mode="nearest")
    if np.all(out):
        out[:] = False
    out &= image > threshold
    out[:1] = out[-1:] = False
    out[:, :1] = out[:, -1:] = False
    coord = np.nonzero(out)
    order = np.argsort(-image[coord], kind="stable")
    return np.transpose(coord)[order]


def filter_peaks_by_distance(image: np.ndarray, peaks_xy: np.ndarray, min_distance: float) -> np.ndarray:
    """mtflearn/features/_local_max_v2.py:6-43 -- visit peaks by descending intensity; a visited peak that
    is still kept suppresses every other peak within Euclidean distance <= min_distance.  Equal
    intensities are visited in the incoming (raster) order here (the reference: numpy's unstable argsort)."""
    from scipy.spatial import cKDTree
    peaks_xy = np.asarray(peaks_xy)
    if len(peaks_xy) == 0:
        return peaks_xy.reshape(0, 2)
    vals = np.array([image[y, x] for x, y in peaks_xy])
    peaks_xy = peaks_xy[np.argsort(-vals, kind="stable")]
    tree = cKDTree(peaks_xy)
    keep = np.ones(len(peaks_xy), dtype=bool)
    for i, p in enumerate(peaks_xy):
        if keep[i]:
            for j in tree.query_ball_point(p, r=min_distance):
                if j != i:
                    keep[j] = False
    return peaks_xy[keep]


def local_max(image: np.ndarray, min_distance: float, threshold=None) -> np.ndarray:
    """mtflearn/features/_local_max_v2.py:46-66: (N, 2) peaks as (x, y), brightest first."""
    peaks = peak_local_max_md1(image, threshold)
    return filter_peaks_by_distance(image, peaks[:, ::-1], min_distance)


# --------------------------------------------------------------------------- #
# PCA of the feature matrix ("next" row f4)                                     #
# --------------------------------------------------------------------------- #
def pca(X: np.ndarray, n_components: int = 2) -> np.ndarray:
    """mtflearn/features/_dimension_reduction.py:3-6 -- ``PCA(n_components).fit_transform(X)``.  scikit-learn
    is a third-party dependency (unpinned in the reference's pyproject.toml; 1.9.0 in this image); this is its
    covariance route (``svd_solver='covariance_eigh'``, chosen by 'auto' for n_samples >= 10 n_features):
    C = (X^T X - n mu mu^T)/(n-1), eigh, components flipped so that their largest-magnitude entry is positive
    (``svd_flip(u_based_decision=False)``), scores = (X - mu) V^T.  Pinned against the live reference in
    tests/golden/pca.npz."""
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    mean = X.sum(axis=0) / n
    cov = (X.T @ X - n * np.outer(mean, mean)) / (n - 1)
    evals, evecs = np.linalg.eigh((cov + cov.T) * 0.5)
    vt = evecs[:, ::-1].T[:n_components].copy()
    idx = np.argmax(np.abs(vt), axis=1)
    signs = np.sign(vt[np.arange(len(vt)), idx])
    signs[signs == 0] = 1.0
    vt *= signs[:, None]
    return (X - mean) @ vt.T


# ---------------------------------------------------------------------------------------------------------------
# "next" row f1, second part: lattice coordinates and the TMD simulator
# ---------------------------------------------------------------------------------------------------------------
def honeycomb_coords(size: int, l: float = 12.0, angle: float = 0.0, random_shift: bool = True, seed=None,
                     jitter: float = 0.0):
    """All generated A / B sites of ``HoneyCombLattice`` (mtflearn/datasets/_honeycomb_lattice.py:69-140): origin
    shift u1, u2 = rng.random(2) (``:75-79``), sites n1, n2 in [-N, N] with N = ceil(size/l)+3 (``:85, :112-117``),
    rotation, centring, jitter drawn A first then B (``:123-135``).  Returns (coords_A, coords_B, rng)."""
    rng = np.random.default_rng(seed)
    u1, u2 = rng.random(2) if random_shift else (0.0, 0.0)
    a1 = np.array([1.5 * l, np.sqrt(3.0) * l / 2.0])
    a2 = np.array([1.5 * l, -np.sqrt(3.0) * l / 2.0])
    d_a, d_b = np.array([0.0, 0.0]), np.array([float(l), 0.0])
    n_idx = int(np.ceil(int(size) / float(l))) + 3
    offset = u1 * a1 + u2 * a2
    idx = np.arange(-n_idx, n_idx + 1, dtype=np.float64)
    big_r = (idx[:, None, None] * a1 + idx[None, :, None] * a2).reshape(-1, 2)
    th = np.deg2rad(float(angle))
    rot = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    centre = np.array([int(size) / 2.0, int(size) / 2.0])
    out = []
    for d in (d_a, d_b):
        out.append((big_r + d + offset) @ rot.T + centre)
    if jitter > 0.0:
        out[0] = out[0] + rng.normal(0.0, jitter, out[0].shape)
        out[1] = out[1] + rng.normal(0.0, jitter, out[1].shape)
    return out[0], out[1], rng


def tmd_blur(shape_hw, atoms: np.ndarray, scales: np.ndarray, sigma: float, amp: float) -> np.ndarray:
    """One species of ``TMDImageSimulator.simulate`` (mtflearn/datasets/_tmd_simulator.py:151-186):
    ``place_atoms_delta`` (round to pixels, drop outside, float32 ``np.add.at``), ``gaussian_kernel`` of size
    ``int(6 sigma) | 1`` and ``scipy.signal.fftconvolve(delta, kernel, mode='same')`` -- the same library call."""
    from scipy.signal import fftconvolve
    h, w = shape_hw
    delta = np.zeros((h, w), dtype=np.float32)
    xi = np.round(atoms[:, 0]).astype(int)
    yi = np.round(atoms[:, 1]).astype(int)
    ok = (xi >= 0) & (xi < w) & (yi >= 0) & (yi < h)
    np.add.at(delta, (yi[ok], xi[ok]), scales[ok])
    ksize = int(6 * sigma) | 1
    r = np.arange(-ksize // 2 + 1, ksize // 2 + 1)
    gx, gy = np.meshgrid(r, r)
    kernel = amp * np.exp(-(gx ** 2 + gy ** 2) / (2 * sigma ** 2))
    return fftconvolve(delta, kernel, mode="same"), delta


# ---------------------------------------------------------------------------------------------------------------
# "next" row f4, second part: cluster labels (the reference delegates to scikit-learn; same library calls here)
# ---------------------------------------------------------------------------------------------------------------
def _reorder_by_count(lbs: np.ndarray) -> np.ndarray:
    """mtflearn/clustering/_clustering_functions.py:18-22: clusters renumbered by decreasing size."""
    unique, counts = np.unique(lbs, return_counts=True)
    lbs_order = np.argsort(counts)[::-1]
    order_dict = dict(zip(lbs_order, unique))
    return np.vectorize(order_dict.get)(lbs)


def kmeans_lbs(X: np.ndarray, n: int, random_state=0) -> np.ndarray:
    """``kmeans_lbs`` (mtflearn/clustering/_clustering_functions.py:8-23): sklearn KMeans(n_clusters=n, random_state)."""
    from sklearn.cluster import KMeans
    return _reorder_by_count(KMeans(n_clusters=n, random_state=random_state).fit(X).labels_)


def gmm_lbs(X: np.ndarray, n: int, random_state=0) -> np.ndarray:
    """``gmm_lbs`` (``:25-34``): sklearn GaussianMixture(n, covariance_type='full', random_state).fit(X).predict(X)."""
    from sklearn.mixture import GaussianMixture
    return _reorder_by_count(GaussianMixture(n, covariance_type="full", random_state=random_state).fit(X).predict(X))


# ---------------------------------------------------------------------------------------------------------------
# "next" row f4, third part: patch-SVD denoiser
# ---------------------------------------------------------------------------------------------------------------
def patch_starts(extent: int, patch: int, step: int) -> np.ndarray:
    """``_patch_start_indices`` (mtflearn/denoise/_denoise_svd.py:10-20): every ``step`` pixels plus the last start."""
    last = extent - patch
    idx = np.arange(0, last, step)
    if idx.size == 0 or idx[-1] != last:
        idx = np.append(idx, last)
    return idx


def denoise_extract(img: np.ndarray, patch: int, step: int) -> np.ndarray:
    """``extract_patches`` (``:22-48``): patches at the start grid, row-major."""
    ys, xs = patch_starts(img.shape[0], patch, step), patch_starts(img.shape[1], patch, step)
    return np.stack([img[y:y + patch, x:x + patch] for y in ys for x in xs])


def denoise_reconstruct(patches: np.ndarray, shape_hw, step: int) -> np.ndarray:
    """``reconstruct_patches`` (``:51-70``): overlap-add in grid order, divided by the overlap count."""
    h, w = shape_hw
    k = patches.shape[1]
    ys, xs = patch_starts(h, k, step), patch_starts(w, patches.shape[2], step)
    img, cnt = np.zeros((h, w)), np.zeros((h, w))
    it = iter(patches)
    for y in ys:
        for x in xs:
            img[y:y + k, x:x + patches.shape[2]] += next(it)
            cnt[y:y + k, x:x + patches.shape[2]] += 1.0
    return img / cnt


def denoise_svd_exact(img: np.ndarray, patch: int, n_components: int, step: int):
    """``denoise_svd`` (``:79-120``) with the exact truncated SVD in place of sklearn's randomized one (which the
    reference calls with random_state=None; 7 power iterations bring it within 1e-7 of this on the golden frame)."""
    p = denoise_extract(img, patch, step)
    u, s, vt = np.linalg.svd(p.reshape(p.shape[0], -1), full_matrices=False)
    low = (u[:, :n_components] * s[:n_components]) @ vt[:n_components]
    return denoise_reconstruct(low.reshape(-1, patch, patch), img.shape, step), s[:n_components]
