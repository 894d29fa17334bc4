"""``KeyPoints`` -- patch gather at peak coordinates on the GPU.

Host-side mirror of ``mtflearn.features.KeyPoints`` for the two members that sit on the hot
path (mtflearn/features/_keypoint.py:44-78): ``clear_border`` (strict-inequality border
filter, pure comparisons on the host) and ``extract_patches`` (the gather, a CUDA kernel).
The centre-of-mass refinement helpers of the reference are off the path and not provided.

Patches are float32 -- the type the projection kernel consumes -- whatever the image dtype;
pass a CUDA tensor image to keep the result in HBM.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib
from ._device import is_torch


def clear_border(pts, shape, size):
    """Keep points with size//2+1 < x < W-size//2-1 and the same for y; pts are (x, y)."""
    pts = np.asarray(pts)
    x, y = pts[:, 0], pts[:, 1]
    margin = size // 2 + 1
    keep = (x > margin) & (x < shape[1] - size // 2 - 1) & (y > margin) & (y < shape[0] - size // 2 - 1)
    return pts[keep]


class KeyPoints:

    def __init__(self, pts, img, size):
        self.shape = tuple(img.shape)
        self.size = size
        self.img = img
        self.pts = clear_border(pts, self.shape, self.size)
        self.patches = None

    def clear_border(self, size):
        # the reference compares y against shape[1] here (_keypoint.py:80-84); kept as is
        x, y = self.pts[:, 0], self.pts[:, 1]
        keep = ((x > size // 2 + 1) & (x < self.shape[1] - size // 2 - 1)
                & (y > size // 2 + 1) & (y < self.shape[1] - size // 2 - 1))
        self.pts = self.pts[keep]

    def extract_patches(self, size=None, flat=False):
        """patches[i] = img[y-size//2 : y-size//2+size, x-size//2 : x-size//2+size] with
        (x, y) = rint(pts[i]); (P, size, size) or (P, size*size) when ``flat``."""
        size = self.size if size is None else size
        torch = _lib.require_cuda()
        lib = _lib.load()
        host_img = not (is_torch(self.img) and self.img.is_cuda)
        if is_torch(self.img):
            img = self.img.to(device="cuda", dtype=torch.float32).contiguous()
        else:
            img = torch.from_numpy(np.ascontiguousarray(self.img, dtype=np.float32)).cuda()
        pts = torch.from_numpy(np.ascontiguousarray(self.pts, dtype=np.float64).reshape(-1, 2)).to(img.device)
        count = int(pts.shape[0])
        out = torch.empty((count, size, size), dtype=torch.float32, device=img.device)
        _lib.check(lib.zb200_gather_patches_f32(int(img.data_ptr()), int(img.shape[0]), int(img.shape[1]),
                                                int(pts.data_ptr()), count, int(size), int(out.data_ptr()),
                                                C.c_void_p(_lib.current_stream_ptr())), "gather_patches")
        if flat:
            out = out.reshape(count, size * size)
        self.patches = out.cpu().numpy() if host_img else out
        return self.patches
