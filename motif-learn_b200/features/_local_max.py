"""``local_max`` -- peak detection on the GPU, the step right before ``KeyPoints`` / ``ZPs``.

Host-side mirror of ``mtflearn.features.local_max`` (mtflearn/features/_local_max_v2.py:46-66):
``skimage.feature.peak_local_max(image, min_distance=1, threshold_abs=threshold)`` followed by the
intensity-ordered radius suppression ``filter_peaks_by_distance`` (``_local_max_v2.py:6-43``).  Same
arguments, same return value: an ``(N, 2)`` int64 array of ``(x, y)`` = (column, row), brightest first.

Differences, stated: the frame is compared as float32 (the type the rest of the path consumes); peaks of
exactly equal intensity are visited in raster order (the reference leaves that to numpy's unstable
``argsort``).  Pass a CUDA tensor to keep the frame in HBM; ``as_tensor=True`` returns the peaks as a CUDA
int32 tensor instead of a host array.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib
from ._device import is_torch


def local_max(image, min_distance, threshold=None, as_tensor: bool = False):
    if image.ndim != 2:
        raise ValueError("image must be a 2D array.")
    torch = _lib.require_cuda()
    lib = _lib.load()
    if is_torch(image):
        dev = image.to(device="cuda", dtype=torch.float32).contiguous()
    else:
        dev = torch.from_numpy(np.ascontiguousarray(image, dtype=np.float32)).cuda()
    h, w = int(dev.shape[0]), int(dev.shape[1])
    stream = C.c_void_p(_lib.current_stream_ptr())
    count, cand = C.c_int64(0), C.c_int64(0)
    # a strict 3x3 maximum occupies at least a 2x2 cell of its own; plateaus can exceed that -> retry
    capacity = h * w // 4 + 16
    out = torch.empty((capacity, 2), dtype=torch.int32, device=dev.device)
    has_thr = 0 if threshold is None else 1
    thr = 0.0 if threshold is None else float(threshold)
    rc = lib.zb200_local_max_f32(int(dev.data_ptr()), h, w, float(min_distance), has_thr, thr, int(out.data_ptr()),
                                 capacity, C.byref(count), C.byref(cand), stream)
    if rc < 0 and count.value > capacity:
        capacity = int(count.value)
        out = torch.empty((capacity, 2), dtype=torch.int32, device=dev.device)
        rc = lib.zb200_local_max_f32(int(dev.data_ptr()), h, w, float(min_distance), has_thr, thr, int(out.data_ptr()),
                                     capacity, C.byref(count), C.byref(cand), stream)
    _lib.check(rc, "local_max")
    pts = out[: int(count.value)]
    return pts if as_tensor else pts.cpu().numpy().astype(np.int64)
