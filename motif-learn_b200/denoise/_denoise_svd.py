"""``denoise_svd`` -- patch-SVD denoising on the GPU.

Host-side mirror of ``mtflearn.denoise`` 's ``extract_patches, reconstruct_patches, denoise_svd, DenoiseSVD``
(mtflearn/denoise/_denoise_svd.py:10-144): same names, arguments and error texts.  Steps and where they run:

* ``extract_patches`` -- the strided patch grid (start indices every ``extraction_step`` pixels plus the last
  start, ``:10-20``) is a gather at known window corners: the K2 kernel of the hot path (``zb200_gather_patches_f32``);
* low-rank approximation -- the reference calls ``sklearn.utils.extmath.randomized_svd(patches, n_components,
  random_state=None)``: Gaussian range finder with 10 oversamples, 7 (or 4) LU-normalised power iterations, QR, small
  SVD.  Restated with library GEMMs / LU / QR / SVD in float64 on the device (cuBLAS / cuSOLVER through torch -- plain
  library calls, there is nothing to fuse); with ``random_state=None`` the reference is not reproducible bit for bit,
  parity is stated on the reconstructed frame and the singular values;
* ``reconstruct_patches`` -- overlap-add and division by the overlap count: ``zb200_overlap_add_f64``.

numpy in -> float64 numpy out; a CUDA tensor frame gives CUDA tensors.
"""
from __future__ import annotations

import ctypes as C
import numbers

import numpy as np

from .. import _lib
from ..features._device import is_torch


def _patch_start_indices(image_extent, patch_extent, step):
    if step <= 0:
        raise ValueError("extraction_step must be a positive integer.")
    if patch_extent >= image_extent:
        raise ValueError("patch_size must be strictly smaller than the image size.")
    last_start = image_extent - patch_extent
    indices = np.arange(0, last_start, step)
    if indices.size == 0 or indices[-1] != last_start:
        indices = np.append(indices, last_start)
    return indices


def _frame_on_device(data):
    torch = _lib.require_cuda()
    if is_torch(data):
        return data.to(device="cuda", dtype=torch.float32).contiguous(), data.is_cuda
    return torch.from_numpy(np.ascontiguousarray(data, dtype=np.float32)).cuda(), False


def _extract_device(dev, patch, step):
    """(patches CUDA float32 (ny*nx, k, k), ys, xs) for a square patch of size ``patch``."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    ys = _patch_start_indices(int(dev.shape[0]), patch, step)
    xs = _patch_start_indices(int(dev.shape[1]), patch, step)
    yy, xx = np.meshgrid(ys, xs, indexing="ij")
    centres = np.stack([xx.ravel() + patch // 2, yy.ravel() + patch // 2], axis=1).astype(np.float64)   # window = [c - k//2, +k)
    pts = torch.from_numpy(np.ascontiguousarray(centres)).cuda()
    out = torch.empty((len(centres), patch, patch), dtype=torch.float32, device=dev.device)
    _lib.check(lib.zb200_gather_patches_f32(int(dev.data_ptr()), int(dev.shape[0]), int(dev.shape[1]), int(pts.data_ptr()),
                                            len(centres), patch, int(out.data_ptr()), C.c_void_p(_lib.current_stream_ptr())),
               "gather_patches")
    return out, ys, xs


def _square(patch_shape):
    if isinstance(patch_shape, numbers.Number):
        return int(patch_shape)
    if len(patch_shape) == 2 and patch_shape[0] == patch_shape[1]:
        return int(patch_shape[0])
    raise NotImplementedError("the GPU patch gather handles square patches")


def extract_patches(data, patch_shape=64, extraction_step=1):
    """Patches of ``data`` on the strided start grid, ``(n_patches, k, k)`` (float32, the gather kernel's type)."""
    dev, on_device = _frame_on_device(data)
    out, _, _ = _extract_device(dev, _square(patch_shape), extraction_step)
    return out if on_device else out.cpu().numpy()


def reconstruct_patches(patches, img_shape, reconstruction_step):
    torch = _lib.require_cuda()
    lib = _lib.load()
    if isinstance(img_shape, numbers.Number):
        img_shape = tuple([img_shape] * 2)
    h, w = (int(v) for v in img_shape)
    on_device = is_torch(patches) and patches.is_cuda
    dev = patches if is_torch(patches) else torch.from_numpy(np.ascontiguousarray(patches))
    dev = dev.to(device="cuda", dtype=torch.float64).contiguous()
    kh, kw = int(dev.shape[1]), int(dev.shape[2])
    ys = _patch_start_indices(h, kh, reconstruction_step)
    xs = _patch_start_indices(w, kw, reconstruction_step)
    if len(ys) * len(xs) != int(dev.shape[0]):
        raise ValueError(f"{int(dev.shape[0])} patches do not match the {len(ys)}x{len(xs)} start grid of this image and step")
    d_ys = torch.from_numpy(ys.astype(np.int32)).cuda()
    d_xs = torch.from_numpy(xs.astype(np.int32)).cuda()
    img = torch.empty((h, w), dtype=torch.float64, device=dev.device)
    _lib.check(lib.zb200_overlap_add_f64(int(dev.data_ptr()), int(d_ys.data_ptr()), len(ys), int(d_xs.data_ptr()), len(xs),
                                         kh, kw, h, w, int(img.data_ptr()), C.c_void_p(_lib.current_stream_ptr())), "overlap_add")
    return img if on_device else img.cpu().numpy()


def _randomized_svd(m, n_components, seed=None):
    """sklearn.utils.extmath.randomized_svd(M, n_components) with its defaults (n_oversamples=10, n_iter='auto',
    LU-normalised power iterations, transpose='auto', svd_flip), library linear algebra on the device."""
    torch = _lib.require_cuda()
    n_random = n_components + 10
    n_iter = 7 if n_components < 0.1 * min(m.shape) else 4
    transpose = m.shape[0] < m.shape[1]
    if transpose:
        m = m.T
    rng = np.random.RandomState(seed)
    q = torch.from_numpy(rng.normal(size=(m.shape[1], n_random))).to(m.device)

    def lu_l(a):                                        # scipy.linalg.lu(a, permute_l=True)[0] = P @ L
        p, l, _ = torch.linalg.lu(a)
        return p @ l

    for _ in range(n_iter):
        q = lu_l(m @ q)
        q = lu_l(m.T @ q)
    q, _ = torch.linalg.qr(m @ q, mode="reduced")
    b = q.T @ m
    uhat, s, vt = torch.linalg.svd(b, full_matrices=False)
    u = q @ uhat
    if not transpose:                                   # svd_flip(u, vt): u-based decision
        idx = torch.argmax(torch.abs(u), dim=0)
        signs = torch.sign(u[idx, torch.arange(u.shape[1], device=u.device)])
    else:                                               # svd_flip(u, vt, u_based_decision=False)
        idx = torch.argmax(torch.abs(vt), dim=1)
        signs = torch.sign(vt[torch.arange(vt.shape[0], device=vt.device), idx])
    signs[signs == 0] = 1
    u, vt = u * signs, vt * signs[:, None]
    if transpose:
        return vt[:n_components].T, s[:n_components], u[:, :n_components].T
    return u[:, :n_components], s[:n_components], vt[:n_components]


def denoise_svd(img, patch_size, n_components, extraction_step=None, verbose=True, return_s=False, random_state=None):
    if isinstance(patch_size, numbers.Number):
        patch_size = tuple([patch_size] * 2)
    img_height, img_width = img.shape
    patch_height, patch_width = patch_size
    if patch_height >= img_height or patch_width >= img_width:
        raise ValueError("patch_size must be strictly smaller than the image dimensions.")
    if extraction_step is None:
        extraction_step = max(1, int(patch_size[0] / 4))
    k = _square(patch_size)
    torch = _lib.require_cuda()
    dev, on_device = _frame_on_device(img)
    if verbose:
        print('Extracting reference patches...')
    patches, _, _ = _extract_device(dev, k, extraction_step)
    flat = patches.reshape(patches.shape[0], -1).double()
    if verbose:
        print('Singular value decomposition...')
    u, s, v = _randomized_svd(flat, n_components, random_state)
    if verbose:
        print('Reconstructing...')
    low = (u * s) @ v
    clean = reconstruct_patches(low.reshape(-1, k, k), (img_height, img_width), extraction_step)
    if not on_device:
        clean = clean.cpu().numpy() if is_torch(clean) else clean
        s = s.cpu().numpy()
    return (clean, s) if return_s else clean


class DenoiseSVD:

    def __init__(self, image, n_components, patch_size, extraction_step):
        self.image = image
        self.n_components = n_components
        self.patch_size = patch_size
        self.extraction_step = extraction_step
        self.patches = None
        self.s_values = None
        self.img_clean = None

    def run(self, verbose=False):
        self.img_clean, self.s_values = denoise_svd(self.image, n_components=self.n_components, patch_size=self.patch_size,
                                                    extraction_step=self.extraction_step, return_s=True, verbose=verbose)
        return self.img_clean
