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


# =================================================================================================================
# GPU mirrors of the reference's synthetic-data classes ("next" row f1 of SURVEY.md section 8)
# =================================================================================================================
class HoneyCombLattice:
    """``mtflearn.datasets.HoneyCombLattice`` (mtflearn/datasets/_honeycomb_lattice.py:6-226) with the coordinate
    generation and the drawing on the GPU: same constructor, same attributes, ``get_points()`` and ``to_image()``
    with the reference's arguments.  The reference builds the sites with a Python double loop and draws them with a
    per-atom numpy loop (3.5 s per 2048^2 frame); here ``zb200_lattice_coords_f64`` and ``zb200_render_atoms_f32`` do
    both in a few milliseconds.  Randomness (origin shift, jitter) is drawn on the host from the same
    ``numpy.random.default_rng(seed)`` stream in the same order, so a seeded lattice is the reference's lattice."""

    def __init__(self, size: int = 512, l: float = 12.0, a=None, angle: float = 0.0, random_shift: bool = True,
                 seed=None, jitter: float = 0.0):
        self.size = int(size)
        self.l = float(l)
        inferred_a = self.l * np.sqrt(3.0)
        if a is not None:
            a = float(a)
            if not np.isclose(a, inferred_a, rtol=1e-5, atol=1e-6):
                raise ValueError(
                    f"Inconsistent 'a' and 'l': got a={a}, l={self.l}, "
                    f"but for ideal graphene expect a≈sqrt(3)*l≈{inferred_a:.6f}."
                )
            self.a = a
        else:
            self.a = inferred_a
        self.angle_deg = float(angle)
        self.angle = np.deg2rad(self.angle_deg)
        self.a1 = np.array([1.5 * self.l, np.sqrt(3.0) * self.l / 2.0], dtype=np.float64)
        self.a2 = np.array([1.5 * self.l, -np.sqrt(3.0) * self.l / 2.0], dtype=np.float64)
        self.dA = np.array([0.0, 0.0], dtype=np.float64)
        self.dB = np.array([self.l, 0.0], dtype=np.float64)
        self.rng = np.random.default_rng(seed)
        self.random_shift = random_shift
        if random_shift:
            self.shift_u1, self.shift_u2 = self.rng.random(2)
        else:
            self.shift_u1 = 0.0
            self.shift_u2 = 0.0
        self.jitter = float(jitter)
        self.N = int(np.ceil(self.size / self.l)) + 3
        self._coords_A = None          # CUDA float64 tensors ((2N+1)^2, 2)
        self._coords_B = None

    def set_angle(self, angle: float) -> None:
        self.angle_deg = float(angle)
        self.angle = np.deg2rad(self.angle_deg)
        self._coords_A = None
        self._coords_B = None

    def _generate_coordinates(self) -> None:
        import ctypes as C
        from . import _lib
        torch = _lib.require_cuda()
        lib = _lib.load()
        count = (2 * self.N + 1) ** 2
        offset = self.shift_u1 * self.a1 + self.shift_u2 * self.a2
        ja = jb = None
        if self.jitter > 0.0:                         # the reference draws A first, then B, from the same generator
            ja = torch.from_numpy(self.rng.normal(0.0, self.jitter, (count, 2))).cuda()
            jb = torch.from_numpy(self.rng.normal(0.0, self.jitter, (count, 2))).cuda()
        ca = torch.empty((count, 2), dtype=torch.float64, device="cuda")
        cb = torch.empty_like(ca)
        vec = lambda v: np.ascontiguousarray(v, dtype=np.float64).ctypes.data_as(C.c_void_p)      # noqa: E731
        a1, a2, dA, dB, off = (np.ascontiguousarray(v, dtype=np.float64) for v in (self.a1, self.a2, self.dA, self.dB, offset))
        _lib.check(lib.zb200_lattice_coords_f64(self.N, vec(a1), vec(a2), vec(dA), vec(dB), vec(off), float(self.angle),
                                                self.size / 2.0, None if ja is None else int(ja.data_ptr()),
                                                None if jb is None else int(jb.data_ptr()), int(ca.data_ptr()),
                                                int(cb.data_ptr()), C.c_void_p(_lib.current_stream_ptr())), "lattice_coords")
        self._coords_A, self._coords_B = ca, cb

    def _ensure(self):
        if self._coords_A is None or self._coords_B is None:
            self._generate_coordinates()

    def coordinates(self):
        """All generated sites (A, B) as host arrays -- the reference's ``_coords_A`` / ``_coords_B``."""
        self._ensure()
        return self._coords_A.cpu().numpy(), self._coords_B.cpu().numpy()

    def get_points(self):
        """A and B sublattice coordinates clipped to [0, size) (host arrays, like the reference)."""
        ca, cb = self.coordinates()
        size = self.size
        keep = lambda c: c[(c[:, 0] >= 0.0) & (c[:, 0] < size) & (c[:, 1] >= 0.0) & (c[:, 1] < size)]      # noqa: E731
        return keep(ca), keep(cb)

    def to_image(self, sigma=None, intensity_A: float = 1.0, intensity_B: float = 0.5, normalize: bool = False,
                 as_tensor: bool = False):
        """Rasterise the lattice (tapered Gaussians, cut-off 3 sigma) into a (size, size) float32 frame.
        ``as_tensor=True`` keeps the frame in HBM."""
        import ctypes as C
        from . import _lib
        torch = _lib.require_cuda()
        lib = _lib.load()
        self._ensure()
        if sigma is None:
            sigma = self.l / 4.0
        img = torch.empty((self.size, self.size), dtype=torch.float32, device="cuda")
        stream = C.c_void_p(_lib.current_stream_ptr())
        # sites whose support misses the frame contribute nothing; the kernel bins atoms per tile, so the whole
        # over-generated list can be passed (the reference filters it first for its per-atom Python loop)
        for coords, amp, acc in ((self._coords_A, intensity_A, 0), (self._coords_B, intensity_B, 1)):
            _lib.check(lib.zb200_render_atoms_f32(int(coords.data_ptr()), None, float(amp), int(coords.shape[0]), float(sigma),
                                                  3.0, self.size, self.size, int(img.data_ptr()), acc, stream), "render_atoms")
        if normalize:
            vmax = float(img.max())
            if vmax > 0.0:
                img /= vmax
        return img if as_tensor else img.cpu().numpy()


class TMDImageSimulator:
    """``mtflearn.datasets.TMDImageSimulator`` (mtflearn/datasets/_tmd_simulator.py:7-210): hexagonal TMD lattice with
    vacancies and dopants, rendered as delta functions blurred by a per-species Gaussian.  Same constructor, defect
    API (``add_single_vacancy, add_double_vacancy, add_random_vacancies, add_random_dopants``), attributes
    (``pts, pts_all, labels, lbs, filtered_indices``) and ``simulate(return_masks=False)``.  The lattice bookkeeping
    is vectorised integer / float64 numpy on the host (the reference's pure-Python loops and its O(n^2) visibility
    test take about a minute at 4096^2); placement and blur run on the GPU (``zb200_render_stamps_f32``: every atom
    stamps its kernel directly, which is what the reference's ``fftconvolve(delta, kernel, 'same')`` evaluates up to
    FFT round-off).  Quirks kept: defect indices are GLOBAL atom indices but are applied to the per-species arrays
    (``_tmd_simulator.py:167-174``); species are drawn in sorted label order (the reference iterates a ``set``)."""

    def __init__(self, size=(512, 512), a=30, theta=0.0, species_params=None, basis=None, use_fft=True):
        self.size = size
        self.a = a
        self.theta = theta
        self.use_fft = use_fft
        self.species_params = species_params or {'TM': {'sigma': 2.0, 'A': 1.0}, 'X': {'sigma': 1.5, 'A': 0.6}}
        self.basis = basis or [(0.0, 0.0, 'TM'), (1 / 3, 1 / 3, 'X')]
        self.vacancies = []
        self.dopants = []
        self._cached_coords = None
        self.pts_all = None
        self.pts = None
        self.labels = None
        self.lbs = None
        self.filtered_indices = {}

    def rotation_matrix(self):
        theta = np.deg2rad(self.theta)
        return np.array([[np.cos(theta), -np.sin(theta)], [np.sin(theta), np.cos(theta)]])

    def generate_lattice(self):
        a1 = self.a * np.array([1, 0])
        a2 = self.a * np.array([0.5, np.sqrt(3) / 2])
        R = self.rotation_matrix()
        height, width = self.size
        nx = int(width // self.a * 2) + 4
        ny = int(height // (self.a * np.sqrt(3) / 2) * 2) + 4
        i = np.arange(-nx, nx, dtype=np.float64)[:, None, None]
        j = np.arange(-ny, ny, dtype=np.float64)[None, :, None]
        base = i * a1 + j * a2                                        # (2nx, 2ny, 2), the reference's loop order
        per_basis = []
        for dx, dy, label in self.basis:
            pos = base + dx * a1 + dy * a2
            rot = np.stack([R[0, 0] * pos[..., 0] + R[0, 1] * pos[..., 1],
                            R[1, 0] * pos[..., 0] + R[1, 1] * pos[..., 1]], axis=-1)
            per_basis.append(rot.reshape(-1, 2))
        n_cells, n_basis = per_basis[0].shape[0], len(self.basis)
        labels_basis = [b[2] for b in self.basis]
        self.pts_all = np.stack(per_basis, axis=1).reshape(-1, 2)      # cell-major, basis atom innermost
        all_labels = np.tile(np.array(labels_basis), n_cells)
        coords = {}
        for label in sorted(set(labels_basis)):
            coords[label] = np.concatenate([per_basis[b].reshape(n_cells, 1, 2) for b in range(n_basis) if labels_basis[b] == label],
                                           axis=1).reshape(-1, 2)
        self._cached_coords = coords
        mask = ((self.pts_all[:, 0] >= 0) & (self.pts_all[:, 0] < width) & (self.pts_all[:, 1] >= 0) & (self.pts_all[:, 1] < height))
        self.pts = self.pts_all[mask]
        defect_labels = all_labels.astype(object)
        for _, vac_idx, scale in self.vacancies:
            if scale == 0.5:
                defect_labels[vac_idx] = 'v1'
            elif scale == 0.0:
                defect_labels[vac_idx] = 'v2'
        for _, dop_idx, _ in self.dopants:
            defect_labels[dop_idx] = 'D'
        self.labels = np.array([str(v) for v in defect_labels[mask]])
        label_to_int = {lbl: k for k, lbl in enumerate(sorted(set(str(v) for v in defect_labels)))}
        self.lbs = np.array([label_to_int[lbl] for lbl in self.labels], dtype=int)
        visible = np.where(mask)[0]
        for label in coords:
            idx = np.where(all_labels == label)[0]
            self.filtered_indices[label] = [int(v) for v in idx[np.isin(idx, visible)]]
        return coords

    def add_single_vacancy(self, label='X', index=None):
        if index is not None:
            self.vacancies.append((label, index, 0.5))

    def add_double_vacancy(self, label='X', indices=None):
        if indices is not None and len(indices) >= 1:
            self.vacancies.append((label, indices[0], 0.0))

    def add_random_vacancies(self, label='X', num_single=0, num_double=0, seed=None):
        if self._cached_coords is None:
            self.generate_lattice()
        indices = self.filtered_indices.get(label, [])
        if len(indices) == 0:
            return
        rng = np.random.default_rng(seed)
        rng.shuffle(indices)                                          # a Python list, like the reference
        selected = 0
        for _ in range(num_single):
            if selected >= len(indices):
                break
            self.add_single_vacancy(label, indices[selected])
            selected += 1
        for _ in range(num_double):
            if selected >= len(indices):
                break
            self.add_double_vacancy(label, [indices[selected]])
            selected += 1

    def add_random_dopants(self, label='TM', num_dopants=0, dopant_intensity=0.8, seed=None):
        if self._cached_coords is None:
            self.generate_lattice()
        indices = self.filtered_indices.get(label, [])
        if len(indices) == 0:
            return
        rng = np.random.default_rng(seed)
        for idx in rng.choice(indices, size=min(num_dopants, len(indices)), replace=False):
            self.dopants.append((label, idx, dopant_intensity))

    def _scales(self, label, count):
        scales = np.ones(count, dtype=np.float32)
        for vac_label, vac_idx, scale in self.vacancies:
            if vac_label == label and 0 <= vac_idx < count:
                scales[vac_idx] = scale
        for dop_label, dop_idx, dop_scale in self.dopants:
            if dop_label == label and 0 <= dop_idx < count:
                scales[dop_idx] = dop_scale
        return scales

    def simulate(self, return_masks: bool = False, as_tensor: bool = False):
        import ctypes as C
        from . import _lib
        torch = _lib.require_cuda()
        lib = _lib.load()
        coords = self.generate_lattice()
        height, width = self.size
        total = torch.zeros((height, width), dtype=torch.float32, device="cuda")
        stream = C.c_void_p(_lib.current_stream_ptr())
        masks = {}
        for label, atoms in coords.items():
            d_pts = torch.from_numpy(np.ascontiguousarray(atoms, dtype=np.float64)).cuda()
            d_sc = torch.from_numpy(self._scales(label, len(atoms))).cuda()
            p = self.species_params.get(label, {'sigma': 2.0, 'A': 1.0})
            if not self.use_fft:
                raise NotImplementedError("use_fft=False (scipy.ndimage.gaussian_filter) is not provided on the GPU")
            ksize = int(6 * p['sigma']) | 1
            _lib.check(lib.zb200_render_stamps_f32(int(d_pts.data_ptr()), int(d_sc.data_ptr()), len(atoms), float(p['A']),
                                                   float(p['sigma']), ksize, height, width, int(total.data_ptr()), 1, stream),
                       "render_stamps")
            if return_masks:
                delta = torch.empty_like(total)
                _lib.check(lib.zb200_render_stamps_f32(int(d_pts.data_ptr()), int(d_sc.data_ptr()), len(atoms), 1.0, 1.0, 1,
                                                       height, width, int(delta.data_ptr()), 0, stream), "render_stamps")
                masks[label] = delta if as_tensor else delta.cpu().numpy()
        img = total if as_tensor else total.cpu().numpy()
        return (img, masks) if return_masks else img

    def get_defect_counts(self):
        if self.labels is None:
            raise ValueError("Run simulate() first to generate labels.")
        unique, counts = np.unique(self.labels, return_counts=True)
        return dict(zip(unique, counts))
