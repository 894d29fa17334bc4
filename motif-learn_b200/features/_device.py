"""Device-buffer plumbing shared by the host mirror classes.

torch is used only for what the spec allows it for: device memory, streams and host<->device
copies.  All moment arithmetic is done by the CUDA kernels behind the C ABI.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib


def is_torch(x) -> bool:
    return type(x).__module__.split(".")[0] == "torch"


class Buf:
    """A moment array resident in HBM plus the addressing the kernels need.

    2-D (N, M): item = row, item_stride = M, mode_stride = 1.
    3-D (M, H, W): item = pixel, item_stride = 1, mode_stride = H*W.
    Strides are in elements (complex elements for complex arrays)."""

    def __init__(self, tensor):
        self.t = tensor
        self.ndim = tensor.ndim
        if tensor.ndim == 2:
            self.n_items, self.n_modes = int(tensor.shape[0]), int(tensor.shape[1])
            self.item_stride, self.mode_stride = self.n_modes, 1
        elif tensor.ndim == 3:
            self.n_modes = int(tensor.shape[0])
            self.n_items = int(tensor.shape[1]) * int(tensor.shape[2])
            self.item_stride, self.mode_stride = 1, self.n_items
        else:
            raise ValueError("Data must be 2D or 3D array.")
        self.is_complex = tensor.is_complex()
        torch = _lib.require_cuda()
        base = tensor.dtype
        self.dtype_code = _lib.F32 if base in (torch.float32, torch.complex64) else _lib.F64

    @property
    def ptr(self) -> int:
        return int(self.t.data_ptr())

    def like(self, n_modes: int, complex_out: bool | None = None, real_out: bool = False):
        """Empty buffer with the same item geometry and ``n_modes`` modes."""
        torch = _lib.require_cuda()
        if self.dtype_code == _lib.F32:
            real_t, cplx_t = torch.float32, torch.complex64
        else:
            real_t, cplx_t = torch.float64, torch.complex128
        want_complex = self.is_complex if complex_out is None else complex_out
        if real_out:
            want_complex = False
        dt = cplx_t if want_complex else real_t
        if self.ndim == 2:
            shape = (self.n_items, n_modes)
        else:
            shape = (n_modes, int(self.t.shape[1]), int(self.t.shape[2]))
        return Buf(torch.empty(shape, dtype=dt, device=self.t.device))


def to_device(data):
    """(Buf, was_host).  numpy -> HBM copy; CUDA tensors are used in place (made contiguous)."""
    torch = _lib.require_cuda()
    if is_torch(data):
        if not data.is_cuda:
            data = data.cuda()
            host = True
        else:
            host = False
        t = data
    else:
        arr = np.asarray(data)
        if arr.dtype.kind not in "fc" or arr.dtype.itemsize < 4 or arr.dtype in (np.longdouble, np.clongdouble):
            arr = arr.astype(np.complex128 if arr.dtype.kind == "c" else np.float64)
        t = torch.from_numpy(np.ascontiguousarray(arr)).cuda()
        host = True
    if t.dtype not in (torch.float32, torch.float64, torch.complex64, torch.complex128):
        t = t.to(torch.float64)
    return Buf(t.contiguous()), host


def from_device(buf: Buf, to_host: bool):
    return buf.t.cpu().numpy() if to_host else buf.t


def np_ptr(arr: np.ndarray) -> C.c_void_p:
    return C.c_void_p(arr.ctypes.data)


def i32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def u8(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.uint8))


def f32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def norm_kind(order):
    """np.linalg.norm ``ord`` -> (ZB200_NORM_* or -1, p)."""
    if order is None or order == 2:
        return _lib.NORM_L2, 2.0
    if order == 1:
        return _lib.NORM_L1, 1.0
    if order == np.inf:
        return _lib.NORM_INF, 0.0
    if isinstance(order, (int, float)) and order > 0 and np.isfinite(order):
        return -1, float(order)
    raise ValueError(f"unsupported norm order {order!r} (supported: None, 1, 2, inf, any finite p > 0)")
