"""``zmoments`` -- Zernike-moment container whose algebra runs on the B200.

Host-side mirror of ``mtflearn.features.zmoments`` (mtflearn/features/_zmoments.py:238-493):
same constructor, attributes (``data, n, m, patch_size``), properties and method names,
same return shapes and error texts.  What differs is WHERE the numbers are produced: every
method that does arithmetic (to_complex, to_real, normalize, rotate, rot_maps, mirror_map)
launches a CUDA kernel through the C ABI (include/zernike_b200.h); there is no numpy path.

``data`` may be a numpy array (results come back as numpy, float64/complex128 in ->
float64/complex128 out, like the reference) or a CUDA ``torch.Tensor`` (results stay in
HBM, float32/complex64 kept as such).  Mode bookkeeping (sorting, select) is integer work
on the host.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib
from . import _indexing as ix
from ._device import Buf, f32, from_device, i32, is_torch, norm_kind, np_ptr, to_device, u8


class zmoments:

    def __init__(self, data, n, m, patch_size=None):
        self.n = np.asarray(n)
        self.m = np.asarray(m)
        self.data = data if is_torch(data) else np.asarray(data)
        self.patch_size = patch_size

        if self.n.shape != self.m.shape:
            raise ValueError("`n` and `m` must have the same shape.")
        count = len(self.n)
        if self.data.ndim == 2:
            got = self.data.shape[1]
        elif self.data.ndim == 3:
            got = self.data.shape[0]
        else:
            raise ValueError("Data must be 2D or 3D array.")
        if got != count:
            raise ValueError(f"Data shape mismatch: expected {count} moments but got {got}")

        # canonical (n, m) order; the permutation is the identity for ZPs output, in which
        # case no copy is made (the reference always copies -- same values either way)
        order = ix.mode_order(self.n, self.m)
        if not np.array_equal(order, np.arange(count)):
            self.n, self.m = self.n[order], self.m[order]
            self.data = self._take_modes(order)

    # ------------------------------------------------------------------ helpers
    @property
    def on_device(self) -> bool:
        return is_torch(self.data) and self.data.is_cuda

    def _stream(self):
        return C.c_void_p(_lib.current_stream_ptr())

    def _take_modes(self, index):
        """data[:, index] / data[index] -- integer gather (no arithmetic)."""
        index = np.asarray(index)
        if not is_torch(self.data):
            return self.data[:, index] if self.data.ndim == 2 else self.data[index, :, :]
        lib = _lib.load()
        src, _ = to_device(self.data)
        dst = src.like(len(index))
        idx = i32(index)
        _lib.check(lib.zb200_select_modes(src.dtype_code, int(src.is_complex), src.ptr, src.n_items,
                                          src.item_stride, src.mode_stride, np_ptr(idx), len(idx),
                                          dst.ptr, dst.item_stride, dst.mode_stride, self._stream()),
                   "select_modes")
        return dst.t

    def _wrap(self, buf: Buf, host: bool, n, m):
        return zmoments(data=from_device(buf, host), n=n, m=m, patch_size=self.patch_size)

    def numpy(self) -> np.ndarray:
        """Host copy of ``data`` (no-op for numpy data)."""
        return self.data.cpu().numpy() if is_torch(self.data) else self.data

    def __array__(self, dtype=None, copy=None):
        arr = self.numpy()
        return arr.astype(dtype) if dtype is not None else arr

    # --------------------------------------------------------------- properties
    @property
    def valid_mask(self):
        """Pixels whose window lies inside the image, exactly as the reference marks them
        (_zmoments.py:279-294; for even windows that is one pixel off the true valid
        region -- reproduced, not fixed)."""
        if self.data.ndim == 2 or self.patch_size is None:
            return None
        mask = np.ones(tuple(self.data.shape[1:]), dtype=bool)
        lead = (self.patch_size - 1) // 2
        trail = self.patch_size - 1 - lead
        mask[:lead, :] = False
        mask[-trail:, :] = False
        mask[:, :lead] = False
        mask[:, -trail:] = False
        return mask

    @property
    def is_complex(self):
        return self.data.is_complex() if is_torch(self.data) else np.iscomplexobj(self.data)

    # ------------------------------------------------------------ real <-> complex
    def to_complex(self):
        """Zc[n,m] = Z[n,+m] + i Z[n,-m]  (_zmoments.py:300-316)."""
        if self.is_complex:
            return self
        lib = _lib.load()
        pos, neg, n_c, m_c = ix.complex_pairing(self.n, self.m)
        src, host = to_device(self.data)
        dst = src.like(len(pos), complex_out=True)
        _lib.check(lib.zb200_to_complex(src.dtype_code, src.ptr, src.n_items, src.item_stride, src.mode_stride,
                                        np_ptr(pos), np_ptr(neg), len(pos), dst.ptr, dst.item_stride,
                                        dst.mode_stride, self._stream()), "to_complex")
        return self._wrap(dst, host, n_c, m_c)

    def to_real(self):
        """Inverse of :meth:`to_complex` (_zmoments.py:318-341)."""
        if not self.is_complex:
            return self
        lib = _lib.load()
        pick, imag, n_r, m_r = ix.real_pairing(self.n, self.m)
        src, host = to_device(self.data)
        dst = src.like(len(pick), real_out=True)
        _lib.check(lib.zb200_to_real(src.dtype_code, src.ptr, src.n_items, src.item_stride, src.mode_stride,
                                     np_ptr(pick), np_ptr(imag), len(pick), dst.ptr, dst.item_stride,
                                     dst.mode_stride, self._stream()), "to_real")
        return self._wrap(dst, host, n_r, m_r)

    # ------------------------------------------------------------------ algebra
    def normalize(self, order=None):
        """Divide every item by its p-norm over the mode axis (_zmoments.py:344-356)."""
        if self.data.ndim not in (2, 3):
            raise ValueError("Input must be a 2D or 3D array.")
        lib = _lib.load()
        kind, p = norm_kind(order)
        src, host = to_device(self.data)
        dst = src.like(src.n_modes)
        _lib.check(lib.zb200_normalize(src.dtype_code, int(src.is_complex), src.ptr, src.n_items, src.n_modes,
                                       src.item_stride, src.mode_stride, kind, p, dst.ptr, self._stream()),
                   "normalize")
        return self._wrap(dst, host, self.n, self.m)

    def select(self, m_select):
        """Keep the modes whose |m| is listed, in their original order (_zmoments.py:359-369)."""
        if self.data.ndim not in (2, 3):
            raise ValueError("Invalid Zernike moment array shape, it can only be 2D or 3D.")
        keep = ix.select_index(self.m, m_select)
        return zmoments(data=self._take_modes(keep), n=self.n[keep], m=self.m[keep], patch_size=self.patch_size)

    def unselect(self, m_unselect):
        """Drop the modes whose |m| is listed (_zmoments.py:371-374)."""
        if self.data.ndim not in (2, 3):
            raise ValueError("Invalid Zernike moment array shape, it can only be 2D or 3D.")
        keep = ix.select_index(self.m, m_unselect, invert=True)
        return zmoments(data=self._take_modes(keep), n=self.n[keep], m=self.m[keep], patch_size=self.patch_size)

    def rotate(self, theta):
        """Complex moments of the pattern rotated by ``theta`` degrees: Zc * exp(-i m theta)
        (_zmoments.py:377-418)."""
        lib = _lib.load()
        zc = self.to_complex()
        src, host = to_device(zc.data)
        dst = src.like(src.n_modes)
        mm = i32(zc.m)
        _lib.check(lib.zb200_rotate(src.dtype_code, src.ptr, src.n_items, src.n_modes, src.item_stride,
                                    src.mode_stride, np_ptr(mm), float(np.deg2rad(theta)), dst.ptr, self._stream()),
                   "rotate")
        return zmoments(data=from_device(dst, host), n=zc.n, m=zc.m, patch_size=self.patch_size)

    def rot_maps(self, n_folds, p=2, m_unselect=None):
        """n-fold symmetry scores (_zmoments.py:420-462): with d = Z_sel/||Z_sel||_p,
        S_f = sum_j w_f(|m_j|) d_j^2.  One fused kernel: selection, norm, square and the
        weighted sum never leave registers.  Returns (N,F) or (F,H,W)."""
        if self.data.ndim not in (2, 3):
            raise ValueError("Input must be a 2D or 3D array.")
        if m_unselect is None:
            m_unselect = (0, 1)
        elif 0 not in m_unselect:
            raise ValueError("m=0 must be included in m_unselect.")
        zr = self.to_real() if self.is_complex else self
        return _rot_scores(zr, n_folds, p, m_unselect)

    def mirror_map(self, theta=None, p=2, m_unselect=(0, 1)):
        """Mirror-symmetry score: max over angles of sum_c Re(Zc^2 e^{-i m theta})
        (_zmoments.py:464-493).  Returns (N,) or (H,W)."""
        lib = _lib.load()
        if theta is None:
            theta = np.linspace(0, 2 * np.pi, 360, endpoint=False)
        theta = np.ascontiguousarray(np.atleast_1d(theta), dtype=np.float64)
        zm = self.to_real() if self.is_complex else self
        zm = zm.unselect(m_unselect)
        if p is not None:
            zm = zm.normalize(order=p)
        zc = zm.to_complex()
        src, host = to_device(zc.data)
        torch = _lib.require_cuda()
        shape = (src.n_items,) if src.ndim == 2 else tuple(zc.data.shape[1:])
        real_t = torch.float32 if src.dtype_code == _lib.F32 else torch.float64
        out = torch.empty(shape, dtype=real_t, device=src.t.device)
        mm = i32(zc.m)
        _lib.check(lib.zb200_mirror_scores(src.dtype_code, src.ptr, src.n_items, src.n_modes, src.item_stride,
                                           src.mode_stride, np_ptr(mm), np_ptr(theta), len(theta),
                                           int(out.data_ptr()), self._stream()), "mirror_scores")
        return out.cpu().numpy() if host else out


def rot_weight_tables(m_all, n_folds, m_unselect):
    """(W[F, M] float64 over ALL modes with zeros on unselected ones, select[M] uint8)."""
    m_all = np.asarray(m_all)
    keep = ix.select_index(m_all, m_unselect, invert=True)
    sel = np.zeros(len(m_all), dtype=np.uint8)
    sel[keep] = 1
    w = np.zeros((len(ix.check_array1d(n_folds)), len(m_all)), dtype=np.float64)
    if len(keep):
        w[:, keep] = ix.construct_rot_maps_matrix(n_folds, m_all[keep])
    return np.ascontiguousarray(w), u8(sel)


def norm_code(p):
    if p is None:
        return _lib.NORM_NONE
    kind, _ = norm_kind(p)
    if kind < 0:
        raise ValueError("rot_maps supports p in {None, 1, 2, inf}")
    return kind


def _rot_scores(zr: "zmoments", n_folds, p, m_unselect):
    lib = _lib.load()
    torch = _lib.require_cuda()
    w, sel = rot_weight_tables(zr.m, n_folds, m_unselect)
    n_f = w.shape[0]
    src, host = to_device(zr.data)
    real_t = torch.float32 if src.dtype_code == _lib.F32 else torch.float64
    if src.ndim == 2:
        out = torch.empty((src.n_items, n_f), dtype=real_t, device=src.t.device)
        ois, ofs = n_f, 1
    else:
        out = torch.empty((n_f,) + tuple(zr.data.shape[1:]), dtype=real_t, device=src.t.device)
        ois, ofs = 1, src.n_items
    _lib.check(lib.zb200_rot_scores(src.dtype_code, src.ptr, src.n_items, src.n_modes, src.item_stride,
                                    src.mode_stride, np_ptr(w), np_ptr(sel), n_f, norm_code(p),
                                    int(out.data_ptr()), ois, ofs, C.c_void_p(_lib.current_stream_ptr())),
               "rot_scores")
    return out.cpu().numpy() if host else out
