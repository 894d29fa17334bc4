"""Synthetic STEM-like inputs for tests and benchmarks (host-side, numpy, vectorised).

The reference draws its test images with ``mtflearn.datasets.HoneyCombLattice`` -- a Python
per-atom loop that takes 3.5 s per 2048x2048 frame and cannot travel to the GPU box.  This
module generates inputs of the same family (two-sublattice honeycomb of tapered Gaussian
atoms, bond length in pixels, optional rotation / jitter / vacancies / dopants / noise) in
one vectorised pass.  It is an input generator, not a port: frames are not bit-identical to
the reference's, and parity is always judged against the oracle on the SAME array.
"""
from __future__ import annotations

import numpy as np


def honeycomb_points(shape, bond: float = 12.0, angle: float = 0.0, seed: int | None = 0,
                     jitter: float = 0.0, margin: float = 0.0):
    """Atom centres (x, y) of a honeycomb lattice covering ``shape`` = (H, W).

    Returns (pts, sublattice) with sublattice 0/1; ``margin`` keeps atoms whose centre lies up
    to that many pixels outside the frame (so their tails still get rendered)."""
    h, w = shape
    rng = np.random.default_rng(seed)
    a1 = np.array([1.5 * bond, np.sqrt(3.0) * bond / 2.0])
    a2 = np.array([1.5 * bond, -np.sqrt(3.0) * bond / 2.0])
    span = int(np.ceil(max(h, w) / bond)) + 3
    idx = np.arange(-span, span + 1)
    i1, i2 = np.meshgrid(idx, idx, indexing="ij")
    cells = i1.reshape(-1, 1) * a1 + i2.reshape(-1, 1) * a2
    u = rng.random(2)
    cells = cells + u[0] * a1 + u[1] * a2
    pts = np.concatenate([cells, cells + np.array([bond, 0.0])])
    sub = np.concatenate([np.zeros(len(cells), dtype=np.int8), np.ones(len(cells), dtype=np.int8)])
    th = np.deg2rad(angle)
    rot = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    pts = pts @ rot.T + np.array([w / 2.0, h / 2.0])
    if jitter > 0:
        pts = pts + rng.normal(0.0, jitter, pts.shape)
    keep = ((pts[:, 0] >= -margin) & (pts[:, 0] <= w - 1 + margin)
            & (pts[:, 1] >= -margin) & (pts[:, 1] <= h - 1 + margin))
    return pts[keep], sub[keep]


def render_atoms(shape, pts, amps, sigma: float) -> np.ndarray:
    """Sum of Gaussians (smooth-step tapered to zero at 3 sigma) into a float32 frame."""
    h, w = shape
    img = np.zeros(h * w, dtype=np.float64)
    reach = int(np.ceil(3.0 * sigma)) + 1
    off = np.arange(-reach, reach + 1)
    oy, ox = np.meshgrid(off, off, indexing="ij")
    cutoff = 3.0 * sigma
    for s in range(0, len(pts), 32768):
        p = pts[s:s + 32768]
        a = amps[s:s + 32768]
        cx = np.floor(p[:, 0]).astype(np.int64)
        cy = np.floor(p[:, 1]).astype(np.int64)
        xs = cx[:, None, None] + ox[None]
        ys = cy[:, None, None] + oy[None]
        r = np.hypot(xs - p[:, 0, None, None], ys - p[:, 1, None, None])
        t = np.clip(r / cutoff, 0.0, 1.0)
        val = a[:, None, None] * np.exp(-0.5 * (r / sigma) ** 2) * (1.0 - 3.0 * t ** 2 + 2.0 * t ** 3)
        ok = (xs >= 0) & (xs < w) & (ys >= 0) & (ys < h) & (r <= cutoff)
        np.add.at(img, (ys[ok] * w + xs[ok]), val[ok])
    return img.reshape(h, w).astype(np.float32)


def honeycomb_image(size, bond: float = 12.0, seed: int | None = 0, angle: float = 0.0,
                    sigma: float | None = None, amp_a: float = 1.0, amp_b: float = 0.5,
                    jitter: float = 0.0, vacancy_frac: float = 0.0, dopant_frac: float = 0.0,
                    noise: float = 0.0):
    """(img float32 (H,W), pts float64 (P,2) as (x,y)) -- a synthetic honeycomb STEM frame.

    ``size`` is an int (square) or (H, W).  ``vacancy_frac`` of the atoms are removed and
    ``dopant_frac`` scaled by 0.8 ("with defects", BASELINE config 4); ``noise`` adds seeded
    Gaussian noise of that sigma.  ``pts`` are the ground-truth atom centres inside the frame
    (defect sites included), the stand-in for detected peaks."""
    shape = (size, size) if np.isscalar(size) else tuple(size)
    sigma = bond / 4.0 if sigma is None else sigma
    pts, sub = honeycomb_points(shape, bond, angle, seed, jitter, margin=3.0 * sigma + 1)
    amps = np.where(sub == 0, amp_a, amp_b).astype(np.float64)
    rng = np.random.default_rng(None if seed is None else seed + 7919)
    if vacancy_frac > 0:
        amps[rng.random(len(amps)) < vacancy_frac] = 0.0
    if dopant_frac > 0:
        amps[rng.random(len(amps)) < dopant_frac] *= 0.8
    img = render_atoms(shape, pts, amps, sigma)
    if noise > 0:
        img = (img + rng.normal(0.0, noise, img.shape)).astype(np.float32)
    h, w = shape
    inside = (pts[:, 0] >= 0) & (pts[:, 0] < w) & (pts[:, 1] >= 0) & (pts[:, 1] < h)
    return img, pts[inside]


def nfold_patches(size: int = 64, n_fold: int = 3, count: int = 10, centre: bool = True) -> np.ndarray:
    """(count, size, size) float32 test patches: ``n_fold`` Gaussian blobs on a ring (plus an
    optional centre blob), each patch rotated by a different angle -- the n-fold test family
    used to sanity-check rot_maps (a 3-fold patch must score ~1 on fold 3)."""
    yy, xx = np.mgrid[:size, :size].astype(np.float64)
    mid, sig, ring = size / 2.0, size / 10.0, size / 3.0
    out = np.zeros((count, size, size), dtype=np.float64)
    for i, rot in enumerate(np.linspace(0.0, 2 * np.pi, count, endpoint=False)):
        if centre:
            out[i] += np.exp(-((xx - mid) ** 2 + (yy - mid) ** 2) / (2 * sig ** 2))
        for f in range(n_fold):
            ang = rot + 2 * np.pi * f / n_fold
            bx, by = mid + ring * np.cos(ang), mid + ring * np.sin(ang)
            out[i] += np.exp(-((xx - bx) ** 2 + (yy - by) ** 2) / (2 * sig ** 2))
        out[i] /= out[i].max()
    return out.astype(np.float32)


def render_atoms_gpu(shape, pts, amps, sigma: float, r_factor: float = 3.0, out=None):
    """GPU counterpart of the reference's ``add_tapered_gaussian`` (mtflearn/datasets/
    _tapered_gaussian.py:3-97): draws the atoms ``pts`` (rows of (x, y), may lie outside the frame)
    with per-atom or scalar amplitude into a float32 CUDA frame of ``shape`` = (H, W).  Pass
    ``out`` (a CUDA float32 tensor) to add to an existing frame, e.g. the second sub-lattice."""
    import ctypes as C
    from . import _lib
    torch = _lib.require_cuda()
    lib = _lib.load()
    h, w = shape
    pts_np = np.ascontiguousarray(np.asarray(pts, dtype=np.float64).reshape(-1, 2))
    d_pts = torch.from_numpy(pts_np).cuda()
    scalar = np.ndim(amps) == 0
    d_amps = None if scalar else torch.from_numpy(np.ascontiguousarray(np.asarray(amps, dtype=np.float64))).cuda()
    if not scalar and d_amps.numel() != pts_np.shape[0]:
        raise ValueError("If amplitude is array-like, its length must match number of points")
    img = out if out is not None else torch.empty((h, w), dtype=torch.float32, device="cuda")
    _lib.check(lib.zb200_render_atoms_f32(int(d_pts.data_ptr()), None if scalar else int(d_amps.data_ptr()),
                                          float(amps) if scalar else 0.0, pts_np.shape[0], float(sigma), float(r_factor),
                                          int(h), int(w), int(img.data_ptr()), 0 if out is None else 1,
                                          C.c_void_p(_lib.current_stream_ptr())), "render_atoms")
    return img


def honeycomb_frame_gpu(size, bond: float = 12.0, seed: int | None = 0, angle: float = 0.0,
                        sigma: float | None = None, amp_a: float = 1.0, amp_b: float = 0.5, jitter: float = 0.0,
                        vacancy_frac: float = 0.0, dopant_frac: float = 0.0, noise: float = 0.0):
    """``honeycomb_image`` with the drawing done on the GPU (row f1 renderer): the lattice sites come from the
    same host routine (integer / trigonometric bookkeeping, milliseconds), the tapered Gaussians are drawn by
    ``zb200_render_atoms_f32`` and the optional Gaussian noise by the device RNG (seeded).  Returns
    (frame CUDA float32 (H, W), pts float64 numpy (P, 2)) -- a 2048^2 frame takes a few ms instead of seconds,
    which is what makes BASELINE configs 4 and 5 (a 4096^2 frame, a 256-frame series) practical to generate."""
    from . import _lib
    torch = _lib.require_cuda()
    shape = (size, size) if np.isscalar(size) else tuple(size)
    sigma = bond / 4.0 if sigma is None else sigma
    pts, sub = honeycomb_points(shape, bond, angle, seed, jitter, margin=3.0 * sigma + 1)
    amps = np.where(sub == 0, amp_a, amp_b).astype(np.float64)
    rng = np.random.default_rng(None if seed is None else seed + 7919)
    if vacancy_frac > 0:
        amps[rng.random(len(amps)) < vacancy_frac] = 0.0
    if dopant_frac > 0:
        amps[rng.random(len(amps)) < dopant_frac] *= 0.8
    img = render_atoms_gpu(shape, pts, amps, sigma)
    if noise > 0:
        gen = torch.Generator(device="cuda")
        gen.manual_seed(0 if seed is None else int(seed) + 104729)
        img.add_(torch.randn(img.shape, generator=gen, device="cuda", dtype=torch.float32), alpha=float(noise))
    h, w = shape
    inside = (pts[:, 0] >= 0) & (pts[:, 0] < w) & (pts[:, 1] >= 0) & (pts[:, 1] < h)
    return img, pts[inside]
