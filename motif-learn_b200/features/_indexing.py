"""Host-side index bookkeeping of Zernike modes (integers only, no moment arithmetic).

Mirrors the free functions exported by ``mtflearn.features`` with the same names,
argument meaning and error texts (the texts are pinned by the reference's
tests/features/test_zmoments.py:28-58):

* ``nm2j``                      -- mtflearn/features/_zmoments.py:3-69
* ``nm2j_complex``              -- _zmoments.py:71-91
* ``construct_complex_matrix``  -- _zmoments.py:111-132
* ``construct_real_matrix``     -- _zmoments.py:134-196
* ``construct_rot_maps_matrix`` -- _zmoments.py:199-235

plus the pairing tables the CUDA kernels consume instead of dense 0/1/i matrices.
"""
from __future__ import annotations

import numpy as np


def check_array1d(data) -> np.ndarray:
    """Scalar or array-like -> flat 1-D ndarray (_zmoments.py:94-108)."""
    return np.atleast_1d(data).ravel()


def nm2j(n, m):
    n_a, m_a = np.asarray(n), np.asarray(m)
    if n_a.shape != m_a.shape:
        raise ValueError("`n` and `m` must have the same shape.")
    if not np.all(np.isclose(n_a % 1, 0)):
        raise ValueError("Radial order `n` must be integer-valued.")
    if not np.all(np.isclose(m_a % 1, 0)):
        raise ValueError("Azimuthal frequency `m` must be integer-valued.")
    n_i, m_i = n_a.astype(int), m_a.astype(int)
    if np.any(n_i < 0):
        raise ValueError("Radial order `n` must be non-negative.")
    if np.any(np.abs(m_i) > n_i):
        raise ValueError("Azimuthal frequency `m` must satisfy |m| ≤ n.")
    if np.any((n_i - np.abs(m_i)) % 2 != 0):
        raise ValueError("`n - |m|` must be even.")
    j = (n_i * (n_i + 2) + m_i) // 2
    return j.item() if j.ndim == 0 else j


def nm2j_complex(n, m):
    n_a, m_a = np.atleast_1d(n), np.atleast_1d(m)
    if not np.all(n_a >= 0):
        raise ValueError("Radial order n must be non-negative.")
    if not np.all(m_a >= 0):
        raise ValueError("Azimuthal frequency m must be non-negative.")
    if not np.all(np.abs(m_a) <= n_a):
        raise ValueError("Azimuthal frequency m must satisfy |m| ≤ n.")
    if not np.all((n_a - np.abs(m_a)) % 2 == 0):
        raise ValueError("n - |m| must be even.")
    base = n_a ** 2 + 2 * n_a + 2 * m_a
    idx = np.where(n_a % 2 == 0, base // 4, (base - 1) // 4)
    return idx.item() if idx.size == 1 else idx


def mode_order(n, m) -> np.ndarray:
    """Stable permutation that sorts modes by (n, m) -- the zmoments ctor order
    (_zmoments.py:269-271)."""
    return np.lexsort((np.asarray(m), np.asarray(n)))


def complex_pairing(n, m):
    """For modes already sorted by (n, m): tables describing Zc = C @ Z without the matrix.

    Returns (pos, neg, n_c, m_c): for complex row c, ``pos[c]`` is the column holding the
    m>=0 member (coefficient 1) and ``neg[c]`` the m<0 member (coefficient 1j); -1 when
    that member is absent.  ``n_c``/``m_c`` follow the reference rule of taking them from
    the coefficient-1 member (_zmoments.py:309-312)."""
    n, m = np.asarray(n).astype(int), np.asarray(m).astype(int)
    tag = np.atleast_1d(nm2j_complex(n, np.abs(m)))
    rows = np.unique(tag)
    where = {v: r for r, v in enumerate(rows.tolist())}
    pos = np.full(len(rows), -1, dtype=np.int32)
    neg = np.full(len(rows), -1, dtype=np.int32)
    for col, (t, mm) in enumerate(zip(tag.tolist(), m.tolist())):
        if mm >= 0:
            pos[where[t]] = col
        else:
            neg[where[t]] = col
    n_c = np.where(pos >= 0, np.abs(n)[np.maximum(pos, 0)], 0).astype(int)
    m_c = np.where(pos >= 0, np.abs(m)[np.maximum(pos, 0)], 0).astype(int)
    return pos, neg, n_c, m_c


def construct_complex_matrix(n, m) -> np.ndarray:
    """Dense 0/1/i matrix with Zc = C @ Z (kept for API parity; kernels use the pairing)."""
    n, m = np.asarray(n), np.asarray(m)
    order = mode_order(n, m)
    pos, neg, _, _ = complex_pairing(n[order], m[order])
    mat = np.zeros((len(pos), len(n)), dtype=complex)
    r = np.arange(len(pos))
    mat[r[pos >= 0], pos[pos >= 0]] = 1
    mat[r[neg >= 0], neg[neg >= 0]] = 1j
    return mat


def real_pairing(n_c, m_c):
    """Inverse tables: real modes (n, +-m) sorted by (n, m); ``src[j]`` is the complex row
    feeding real mode j and ``take_imag[j]`` says whether its imaginary part is used."""
    n_c, m_c = np.asarray(n_c).astype(int), np.asarray(m_c).astype(int)
    n_r, m_r, src, imag = [], [], [], []
    for row, (nn, mm) in enumerate(zip(n_c.tolist(), m_c.tolist())):
        n_r.append(nn); m_r.append(mm); src.append(row); imag.append(0)
        if mm != 0:
            n_r.append(nn); m_r.append(-mm); src.append(row); imag.append(1)
    n_r, m_r = np.array(n_r, dtype=int), np.array(m_r, dtype=int)
    order = mode_order(n_r, m_r)
    return (np.asarray(src, dtype=np.int32)[order], np.asarray(imag, dtype=np.uint8)[order],
            n_r[order], m_r[order])


def construct_real_matrix(n, m):
    """(inv_matrix, n_real, m_real) with Z = inv_matrix @ Zc (real part)."""
    src, imag, n_r, m_r = real_pairing(n, m)
    inv = np.zeros((len(src), len(np.asarray(n))), dtype=complex)
    inv[np.arange(len(src)), src] = np.where(imag == 1, -1j, 1.0)
    return inv, n_r, m_r


def construct_rot_maps_matrix(n_folds, m) -> np.ndarray:
    """W[f, j]: +1 where |m_j| is a multiple of the fold (and > 1), 0 for |m_j| in {0, 1},
    -1/(fold-1) elsewhere (0 for fold <= 1)."""
    folds = check_array1d(n_folds)
    m_abs = np.abs(check_array1d(m))
    mat = np.zeros((len(folds), len(m_abs)))
    low = m_abs <= 1
    for r, fold in enumerate(folds):
        match = (m_abs % fold == 0) & ~low
        mat[r, match] = 1
        mat[r, ~(match | low)] = -1.0 / (fold - 1) if fold > 1 else 0
    return mat


def select_index(m, m_select, invert: bool = False) -> np.ndarray:
    """Indices kept by ``select`` (or ``unselect`` when invert) in original order."""
    m_abs = np.abs(np.asarray(m))
    wanted = np.unique(np.abs(check_array1d(m_select)))
    if invert:
        wanted = np.array([v for v in np.unique(m_abs) if v not in wanted])
    return np.where(np.isin(m_abs, wanted))[0]
