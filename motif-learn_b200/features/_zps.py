"""``ZPs`` -- Zernike-polynomial transformer backed by sm_100a CUDA kernels.

Host-side mirror of ``mtflearn.features.ZPs`` (mtflearn/features/_zps.py:11-197): same
constructor arguments, validation messages, attributes (``n_max, size, n, m, polynomials``)
and methods (``get_polynomials, fit, transform, fit_transform``), still an sklearn
``BaseEstimator``/``TransformerMixin`` so ``get_params``/``clone``/pipelines keep working.

Where the reference calls ``numpy.dot`` (3-D input) or ``scipy.signal.fftconvolve`` (2-D
input), this class calls the C ABI of libzernike_b200.so: the basis is generated once in
fp64 on the GPU and cached in HBM (plan), patch stacks go through the projection kernel,
images through the sliding-window map kernel.  Extra, keyword-only options:

``precision``  'auto' | 'fp32' | 'tf32x3' | 'tf32' | 'f16x3' | 'f16' -- arithmetic of the contraction.
               'auto' picks the fp32-grade tensor-core path when the plan supports it (patch stacks:
               'tf32x3'; dense maps: the fp16-split 'f16x3', then 'tf32x3'), otherwise the fp32 SIMT
               kernel.  'f16x3'/'f16' are dense-map kernels; for patch stacks they mean 'tf32x3'/'tf32'.
               Stated error bounds: DESIGN.md.
``value_max``  None | float -- an upper bound of ``|pixel values|`` of the patch stacks this transformer will see
               (e.g. 1.0 for frames normalised to [0, 1]).  With it, CUDA patch stacks run through the fp16-split
               projection ('f16x3': three fp16 tensor-core passes on x = x1 + x2, V = b1 + b2 -- fp32-grade like
               'tf32x3', a quarter fewer tensor-core instructions and half the basis traffic).  fp16 has a narrow
               exponent range, so the inputs are scaled by a power of two derived from this bound; values beyond it
               by more than 4x overflow to inf / NaN in the result (loud, not silent).  Without it 'auto' is 'tf32x3',
               except for 64- and 128-pixel windows (n_max <= 20): those run the mirror-folded fp16-split kernel
               (a quarter of the multiply-adds), which takes its scale from a sample of the stack when no bound is
               given and is backed by 'tf32x3' on the same stream should an unsampled value overflow.
``output``     'auto' | 'numpy' | 'torch' -- 'auto' returns what it was given: numpy in ->
               float64 numpy out (like the reference); CUDA tensor in -> float32 CUDA tensor.
"""
from __future__ import annotations

import ctypes as C
import threading
import warnings
from typing import Optional

import numpy as np
from sklearn.base import BaseEstimator, TransformerMixin

from .. import _lib
from ._device import is_torch, np_ptr
from ._zmoments import norm_code, rot_weight_tables, zmoments

def _host_f64(t):
    """float32 CUDA tensor -> float64 numpy: chunked D2H through pinned staging, widened on the host by a
    few threads while the next chunk is in flight (zb200_download_as_f64)."""
    _lib.require_cuda()
    t = t.contiguous()
    out = np.empty(tuple(t.shape), dtype=np.float64)
    _lib.check(_lib.load().zb200_download_as_f64(int(t.data_ptr()), t.numel(), np_ptr(out),
                                                 C.c_void_p(_lib.current_stream_ptr())), "download_as_f64")
    return out


_plans: dict = {}
_plans_lock = threading.Lock()


def _plan_for(n_max: int, size: int):
    """The HBM-resident plan (basis + packed operands) for (n_max, size) on the current device."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    key = (int(n_max), int(size), torch.cuda.current_device())
    with _plans_lock:
        plan = _plans.get(key)
        if plan is None:
            handle = C.c_void_p()
            _lib.check(lib.zb200_plan_create(int(n_max), int(size), C.byref(handle)), "plan_create")
            plan = handle
            _plans[key] = plan
    return plan


def release_plans(device=None) -> int:
    """Destroy the cached plans (fp64 basis, packed operands, host staging buffers: a few MB for n_max=12 up to
    1-2 GB once the host frame route has run) of ``device`` (all devices when None); returns how many were freed.
    Plans are rebuilt on demand.  No call that uses a plan may be in flight."""
    _lib.require_cuda()
    lib = _lib.load()
    freed = 0
    with _plans_lock:
        for key in [k for k in _plans if device is None or k[2] == device]:
            lib.zb200_plan_destroy(_plans.pop(key))
            freed += 1
    lib.zb200_trim_scratch()          # the library's cached stream-ordered scratch of the current device
    return freed


def _mode_table(n_max: int):
    lib = _lib.load()
    count = lib.zb200_num_modes(int(n_max))
    n = np.empty(count, dtype=np.int32)
    m = np.empty(count, dtype=np.int32)
    lib.zb200_mode_table(int(n_max), n.ctypes.data_as(C.POINTER(C.c_int32)), m.ctypes.data_as(C.POINTER(C.c_int32)))
    return n.astype(int), m.astype(int)


class ZPs(BaseEstimator, TransformerMixin):
    """
    Zernike Polynomials transformer for computing Zernike moments on a B200.

    Parameters
    ----------
    n_max : int
        Maximum radial order.
    size : int
        Size of the polynomial grid (size x size).
    """

    def __init__(self, n_max: int, size: int, *, precision: str = "auto", output: str = "auto", value_max=None):
        if n_max < 0:
            raise ValueError("n_max must be non-negative.")
        if size <= 0:
            raise ValueError("size must be positive.")
        if n_max > size:
            raise ValueError(
                f"n_max={n_max} exceeds size={size}. This will produce "
                f"meaningless results. Use n_max <= {size//2} for accurate moments."
            )
        if n_max > size / 2:
            recommended_max = size // 2
            warnings.warn(
                f"n_max={n_max} exceeds recommended limit of size/2≈{recommended_max}. "
                f"High-order Zernike moments may suffer from aliasing and numerical "
                f"errors. For accurate results, use n_max <= {size//2}; for maximum "
                f"stability, use n_max <= {recommended_max}.",
                UserWarning,
                stacklevel=2
            )
        if precision not in ("auto", "fp32", "tf32", "tf32x3", "f16", "f16x3"):
            raise ValueError("precision must be one of 'auto', 'fp32', 'tf32x3', 'tf32', 'f16x3', 'f16'")
        if output not in ("auto", "numpy", "torch"):
            raise ValueError("output must be one of 'auto', 'numpy', 'torch'")
        self.n_max = n_max
        self.size = size
        if value_max is not None and not (np.isfinite(value_max) and value_max > 0):
            raise ValueError("value_max must be a positive finite bound of |pixel values| (or None)")
        self.precision = precision
        self.output = output
        self.value_max = value_max
        self.n, self.m = _mode_table(n_max)
        self._basis_host = None

    # ------------------------------------------------------------------ basis
    @property
    def _plan(self):
        return _plan_for(self.n_max, self.size)

    @property
    def polynomials(self) -> np.ndarray:
        """(M, size, size) float64 basis, generated on the GPU (K1) and copied out on first use."""
        if self._basis_host is None:
            out = np.empty((len(self.n), self.size, self.size), dtype=np.float64)
            _lib.check(_lib.load().zb200_plan_basis_to_host(self._plan, np_ptr(out)), "plan_basis_to_host")
            self._basis_host = out
        return self._basis_host

    def get_polynomials(self) -> np.ndarray:
        """Return the generated Zernike polynomials."""
        return self.polynomials

    def polynomials_device(self):
        """The fp64 basis as a CUDA tensor (M, size, size)."""
        torch = _lib.require_cuda()
        return torch.from_numpy(self.polynomials).cuda()

    # --------------------------------------------------------------- sklearn API
    def fit(self, X, y: Optional[np.ndarray] = None) -> "ZPs":
        """Fit method (no-op for compatibility with sklearn)."""
        return self

    def fit_transform(self, X, y: Optional[np.ndarray] = None) -> zmoments:
        """Fit and transform (fit is no-op)."""
        return self.fit(X).transform(X)

    def transform(self, images) -> zmoments:
        """2-D image -> dense moment map (M,H,W); 3-D patch stack (N,k,k) -> moments (N,M)."""
        if images.ndim == 2:
            return self._transform_map(images)
        if images.ndim == 3:
            return self._transform_patches(images)
        raise ValueError("Images must be 2D or 3D array.")

    # ------------------------------------------------------------------ checks
    def _validate_size(self, images) -> None:
        if images.ndim == 2:
            height, width = images.shape
        elif images.ndim == 3:
            _, height, width = images.shape
        else:
            raise ValueError("Images must be 2D or 3D array.")
        if images.ndim == 3:
            if height != self.size or width != self.size:
                raise ValueError(
                    f"For batch processing, image size ({height}x{width}) must match "
                    f"polynomial size ({self.size}x{self.size})"
                )
        elif height < self.size or width < self.size:
            raise ValueError(
                f"For FFT convolution, image size ({height}x{width}) must be at least "
                f"as large as polynomial size ({self.size}x{self.size})"
            )

    def _precision_code(self, for_map: bool = False, device_stack: bool = False, host_route: bool = False) -> int:
        """``device_stack``: a CUDA patch stack goes through ``_project`` (which can pass ``value_max``);
        ``host_route``: the library's own host pipelines (no ``value_max`` argument: auto-range only)."""
        lib = _lib.load()
        auto_range = bool(lib.zb200_plan_supports_autorange(self._plan))     # the mirror-folded kernel: 64 / 128 px windows
        if self.precision == "auto":
            if for_map:
                for code in (_lib.PREC_F16X3, _lib.PREC_TF32X3):
                    if lib.zb200_plan_supports_map(self._plan, code):
                        return code
                return _lib.PREC_FP32
            ok = lib.zb200_plan_supports(self._plan, _lib.PREC_TF32X3, _lib.OUT_REAL)
            if ok and ((device_stack and self.value_max is not None) or ((device_stack or host_route) and auto_range)):
                return _lib.PREC_F16X3
            return _lib.PREC_TF32X3 if ok else _lib.PREC_FP32
        code = _lib.PRECISIONS[self.precision]
        if not for_map and code in (_lib.PREC_F16, _lib.PREC_F16X3):
            if code == _lib.PREC_F16X3 and ((device_stack and self.value_max is not None)
                                            or ((device_stack or host_route) and auto_range)):
                return code
            code = _lib.PREC_TF32 if code == _lib.PREC_F16 else _lib.PREC_TF32X3
        return code

    def _project(self, dev, code, out, out2=None):
        """One launch of the projection on a CUDA stack with the epilogue ``code``; picks the fp16-split kernel when
        ``value_max`` allows it."""
        lib = _lib.load()
        prec = self._precision_code(device_stack=True)
        if prec == _lib.PREC_F16X3 and not lib.zb200_plan_supports(self._plan, prec, code):
            prec = self._precision_code()
        o2 = None if out2 is None else int(out2.data_ptr())
        if prec == _lib.PREC_F16X3:
            _lib.check(lib.zb200_project_patches_ranged_f32(self._plan, int(dev.data_ptr()), int(dev.shape[0]),
                                                            float(self.value_max or 0.0), code, int(out.data_ptr()), o2,
                                                            self._stream()), "project_patches_ranged")
        else:
            _lib.check(lib.zb200_project_patches_f32(self._plan, int(dev.data_ptr()), int(dev.shape[0]), prec, code,
                                                     int(out.data_ptr()), o2, self._stream()), "project_patches")
        return prec

    def _want_host(self, given) -> bool:
        if self.output == "auto":
            return not (is_torch(given) and given.is_cuda)
        return self.output == "numpy"

    @staticmethod
    def _stream():
        return C.c_void_p(_lib.current_stream_ptr())

    # ----------------------------------------------------- K3: patch projection
    def _transform_patches(self, images) -> zmoments:
        self._validate_size(images)
        torch = _lib.require_cuda()
        lib = _lib.load()
        n_img = int(images.shape[0])
        n_modes = len(self.n)
        host_in = not is_torch(images)
        if host_in and self._want_host(images):
            # the numpy user's call: host buffers in, float64 host buffer out, chunked
            # H2D/kernel/D2H pipeline inside the library
            src = np.ascontiguousarray(images, dtype=np.float32)
            out = np.empty((n_img, n_modes), dtype=np.float64)
            _lib.check(lib.zb200_project_patches_host(self._plan, np_ptr(src), n_img, self._precision_code(host_route=True),
                                                      np_ptr(out)), "project_patches_host")
            return zmoments(data=out, n=self.n, m=self.m, patch_size=self.size)
        dev = torch.from_numpy(np.ascontiguousarray(images, dtype=np.float32)).cuda() if host_in else images
        dev = dev.to(device="cuda", dtype=torch.float32).contiguous()
        out = torch.empty((n_img, n_modes), dtype=torch.float32, device=dev.device)
        self._project(dev, _lib.OUT_REAL, out)
        data = _host_f64(out) if self._want_host(images) else out
        return zmoments(data=data, n=self.n, m=self.m, patch_size=self.size)

    def _project_device(self, dev):
        """Real moments (N, M) float32 of a CUDA patch stack, always a CUDA tensor (internal)."""
        torch = _lib.require_cuda()
        out = torch.empty((int(dev.shape[0]), len(self.n)), dtype=torch.float32, device=dev.device)
        self._project(dev, _lib.OUT_REAL, out)
        return zmoments(data=out, n=self.n, m=self.m, patch_size=self.size)

    def transform_features(self, images, kind: str = "abs"):
        """Fused projection epilogues on a CUDA patch stack (tensor-core path only).

        kind='complex' -> complex64 (N,Mc); 'abs' -> |Zc| (N,Mc) the rotation-invariant
        features of notebook "2 How to use ZPs" cell 13; 'abs_phase' -> (|Zc|, angle(Zc))."""
        self._validate_size(images)
        torch = _lib.require_cuda()
        lib = _lib.load()
        code = {"complex": _lib.OUT_COMPLEX, "abs": _lib.OUT_ABS, "abs_phase": _lib.OUT_ABS_PHASE}[kind]
        prec = self._precision_code()
        if not lib.zb200_plan_supports(self._plan, prec, code):
            # the fused epilogue is not available for this shape/precision: real moments from the
            # projection kernel, then the packing kernel (two launches, still all on the GPU)
            dev_in = images if is_torch(images) else torch.from_numpy(np.ascontiguousarray(images, dtype=np.float32))
            dev_in = dev_in.to(device="cuda", dtype=torch.float32).contiguous()
            zc = self._project_device(dev_in).to_complex().data.contiguous()
            if kind == "complex":
                return zc
            real_t = torch.float32 if zc.dtype == torch.complex64 else torch.float64
            mag = torch.empty(zc.shape, dtype=real_t, device=zc.device)
            ph = torch.empty_like(mag) if kind == "abs_phase" else None
            _lib.check(lib.zb200_complex_abs_phase(_lib.F32 if real_t == torch.float32 else _lib.F64, int(zc.data_ptr()),
                                                   zc.numel(), int(mag.data_ptr()),
                                                   None if ph is None else int(ph.data_ptr()), self._stream()),
                       "complex_abs_phase")
            return mag if ph is None else (mag, ph)
        dev = images if is_torch(images) else torch.from_numpy(np.ascontiguousarray(images, dtype=np.float32))
        dev = dev.to(device="cuda", dtype=torch.float32).contiguous()
        n_img, n_c = int(dev.shape[0]), lib.zb200_num_complex_modes(self.n_max)
        out2 = None
        if kind == "complex":
            out = torch.empty((n_img, n_c), dtype=torch.complex64, device=dev.device)
        else:
            out = torch.empty((n_img, n_c), dtype=torch.float32, device=dev.device)
            if kind == "abs_phase":
                out2 = torch.empty_like(out)
        self._project(dev, code, out, out2)
        return out if out2 is None else (out, out2)

    def transform_allgather(self, images, peers, row0: int, kind: str = "real"):
        """Multi-GPU ``transform`` whose final feature gather is part of the projection kernel:
        this rank's CUDA patch stack ``images`` (n, k, k) is projected and the rows land at ``[row0, row0 + n)`` of
        ``peers`` (a ``motif_learn_b200.parallel.PeerArray`` of ``(total, cols)`` float32, cols = M | 2*Mc | Mc for
        kind 'real' | 'complex' | 'abs') ON EVERY RANK: finished tiles are forwarded to the other GPUs over NVLink by
        a warp of the kernel while the next tiles are computed.  Call ``peers.fence()`` afterwards; then
        ``peers.local`` holds all ranks' rows.  Tensor-core precisions only."""
        self._validate_size(images)
        torch = _lib.require_cuda()
        lib = _lib.load()
        code = {"real": _lib.OUT_REAL, "complex": _lib.OUT_COMPLEX, "abs": _lib.OUT_ABS}[kind]
        cols = {"real": len(self.n), "complex": 2 * lib.zb200_num_complex_modes(self.n_max),
                "abs": lib.zb200_num_complex_modes(self.n_max)}[kind]
        if peers.cols != cols:
            raise ValueError(f"peer array has {peers.cols} columns, kind={kind!r} needs {cols}")
        dev = images.to(device="cuda", dtype=torch.float32).contiguous()
        n = int(dev.shape[0])
        if row0 < 0 or row0 + n > peers.rows:
            raise ValueError("rows [row0, row0 + n) fall outside the peer array")
        prec = self._precision_code(device_stack=True)
        if prec not in (_lib.PREC_TF32, _lib.PREC_TF32X3, _lib.PREC_F16X3) or not lib.zb200_plan_supports(self._plan, prec, code):
            raise ValueError("transform_allgather needs a tensor-core precision supported for this shape")
        others = [r for r in range(peers.world) if r != peers.rank]
        ptrs = (C.c_void_p * max(1, len(others)))(*[peers.row_ptr(r, row0) for r in others])
        _lib.check(lib.zb200_project_patches_push_f32(self._plan, int(dev.data_ptr()), n, prec, code,
                                                      float(self.value_max or 0.0),
                                                      C.c_void_p(peers.row_ptr(peers.rank, row0)), ptrs, len(others),
                                                      self._stream()), "project_patches_push")
        return peers.local[row0:row0 + n]

    def symmetry_scores(self, images, n_folds, p=2, m_unselect=None):
        """Fused ``transform(patches).rot_maps(n_folds, p, m_unselect)`` -> (N, F): the n-fold scores of a
        patch stack computed in the projection kernel's epilogue (the moments never leave the SM).
        Falls back to projection + score kernel when the fused epilogue is unavailable."""
        self._validate_size(images)
        if images.ndim != 3:
            raise ValueError("Images must be 2D or 3D array.")
        if m_unselect is None:
            m_unselect = (0, 1)
        elif 0 not in m_unselect:
            raise ValueError("m=0 must be included in m_unselect.")
        torch = _lib.require_cuda()
        lib = _lib.load()
        host_in = not is_torch(images)
        dev = torch.from_numpy(np.ascontiguousarray(images, dtype=np.float32)) if host_in else images
        dev = dev.to(device="cuda", dtype=torch.float32).contiguous()
        wts, sel = rot_weight_tables(self.m, n_folds, m_unselect)
        n_f = wts.shape[0]
        prec = self._precision_code()
        kind = norm_code(p)
        fusable = n_f <= 8 and (prec == _lib.PREC_FP32 or len(self.n) <= (128 if prec == _lib.PREC_TF32X3 else 256))
        if not fusable:
            out = self._project_device(dev).rot_maps(n_folds, p=p, m_unselect=m_unselect)
            return _host_f64(out) if self._want_host(images) else out
        out = torch.empty((int(dev.shape[0]), n_f), dtype=torch.float32, device=dev.device)
        w32 = np.ascontiguousarray(wts, dtype=np.float32)
        _lib.check(lib.zb200_project_patches_scores_f32(self._plan, int(dev.data_ptr()), int(dev.shape[0]), prec,
                                                        np_ptr(w32), np_ptr(sel), n_f, kind, int(out.data_ptr()),
                                                        self._stream()), "project_patches_scores")
        return _host_f64(out) if self._want_host(images) else out

    def transform_peaks(self, image, pts, kind: str = "real", fused=None):
        """Moments of the ``size x size`` windows centred at ``pts`` (rows of (x, y), as kept by
        ``clear_border``) of one frame -- ``KeyPoints(pts, image, size).extract_patches()`` followed by
        ``transform`` (BASELINE config 5), without the patch stack ever crossing the bus.

        kind='real' -> ``zmoments`` (N, M); 'complex' | 'abs' | 'abs_phase' as ``transform_features``.
        Where the result lives is decided once, from ``image`` and ``self.output``: a numpy frame gives
        float64 / complex128 numpy results like the reference (through ``zb200_project_peaks_host``: frame and
        coordinates up, features down), a CUDA frame float32 / complex64 CUDA tensors.
        ``fused=True`` runs gather and projection as ONE kernel (windows go from the L2-resident frame straight
        into the tensor-core operand, no patch stack in HBM), ``False`` the gather kernel followed by the
        projection.  ``None`` (default): 64-pixel windows with n_max <= 13 use the mirror-folded kernel with its
        gathering warpgroup (the fastest route); other shapes the two-kernel route, or the tf32x3 fused kernel
        when the intermediate patch stack would exceed 8 GiB."""
        if image.ndim != 2:
            raise ValueError("Images must be 2D or 3D array.")
        torch = _lib.require_cuda()
        lib = _lib.load()
        codes = {"real": _lib.OUT_REAL, "complex": _lib.OUT_COMPLEX, "abs": _lib.OUT_ABS, "abs_phase": _lib.OUT_ABS_PHASE}
        code = codes[kind]
        prec = self._precision_code()
        to_host = self._want_host(image)
        pts_np = np.ascontiguousarray(np.asarray(pts, dtype=np.float64).reshape(-1, 2))
        can_fuse = (prec == _lib.PREC_TF32X3 and self.size >= 32 and lib.zb200_plan_supports(self._plan, prec, code))
        # 64-pixel windows, n_max <= 13: the mirror-folded kernel gathers the windows itself -- faster than the gather
        # kernel + projection (no patch stack in HBM), so it is the default there
        fold_gather = (self.precision in ("auto", "f16x3") and bool(lib.zb200_plan_supports_folded_gather(self._plan))
                       and lib.zb200_plan_supports(self._plan, _lib.PREC_F16X3, code))
        if fold_gather:
            can_fuse, prec = True, _lib.PREC_F16X3
        host_route = (to_host and not is_torch(image) and kind != "abs_phase"
                      and (prec == _lib.PREC_FP32 and kind == "real" or lib.zb200_plan_supports(self._plan, prec, code)))
        if fused is None:
            # a numpy frame goes through the library's host pipeline (which fuses on its own where it can)
            fused = (not host_route) and (fold_gather or (can_fuse and pts_np.shape[0] * self.size * self.size * 4 > (8 << 30)))
        if fused and not can_fuse:
            raise ValueError("fused gather+projection needs precision 'tf32x3' (or 'auto' on a supported shape) and size >= 32")
        if host_route and not fused:
            return self.transform_peaks_batch([image], [pts_np], kind)[0]
        dev = self._image_on_device(image)
        if not fused:
            # two kernels: gather to HBM, then the projection at the requested precision
            from ._keypoint import KeyPoints
            kp = KeyPoints.__new__(KeyPoints)
            kp.shape, kp.size, kp.img, kp.pts, kp.patches = tuple(dev.shape), self.size, dev, pts_np, None
            patches = kp.extract_patches()
            res = self._project_device(patches) if kind == "real" else self.transform_features(patches, kind)
        else:
            dpts = torch.from_numpy(pts_np).to(dev.device)
            count = int(dpts.shape[0])
            n_c = lib.zb200_num_complex_modes(self.n_max)
            out2 = None
            if kind == "real":
                out = torch.empty((count, len(self.n)), dtype=torch.float32, device=dev.device)
            elif kind == "complex":
                out = torch.empty((count, n_c), dtype=torch.complex64, device=dev.device)
            else:
                out = torch.empty((count, n_c), dtype=torch.float32, device=dev.device)
                if kind == "abs_phase":
                    out2 = torch.empty_like(out)
            _lib.check(lib.zb200_project_peaks_f32(self._plan, int(dev.data_ptr()), int(dev.shape[0]), int(dev.shape[1]),
                                                   int(dpts.data_ptr()), count, prec, code, int(out.data_ptr()),
                                                   None if out2 is None else int(out2.data_ptr()), self._stream()),
                       "project_peaks")
            res = zmoments(data=out, n=self.n, m=self.m, patch_size=self.size) if kind == "real" else (
                out if out2 is None else (out, out2))
        if not to_host:
            return res
        if kind == "real":
            return zmoments(data=_host_f64(res.data), n=self.n, m=self.m, patch_size=self.size)
        if kind == "complex":
            return res.cpu().numpy().astype(np.complex128)
        if kind == "abs":
            return _host_f64(res)
        return tuple(_host_f64(t) for t in res)

    def transform_peaks_batch(self, frames, pts_list, kind: str = "real", out=None):
        """``transform_peaks`` for a series of equally shaped HOST frames (an in-situ series, BASELINE config 5):
        one call of ``zb200_project_peaks_host``, which pipelines the frames over two streams (upload of frame
        f+1 and download + float64 widening of frame f-1 overlap the kernels of frame f).  ``frames`` is a
        sequence of (H, W) numpy arrays (or one (F, H, W) array), ``pts_list`` one (P_f, 2) array of (x, y) per
        frame.  Returns a list with one entry per frame: ``zmoments`` for kind='real', complex128 (P_f, Mc) for
        'complex', float64 (P_f, Mc) for 'abs'.  ``out`` (optional) is a C-contiguous array of the result dtype with
        at least sum(P_f) rows that receives the features -- a frame loop that reuses one buffer does not pay the
        page faults of a fresh 100 MB allocation per call; the returned entries are views of it."""
        _lib.require_cuda()
        lib = _lib.load()
        code = {"real": _lib.OUT_REAL, "complex": _lib.OUT_COMPLEX, "abs": _lib.OUT_ABS}[kind]
        frames = [np.ascontiguousarray(f, dtype=np.float32) for f in frames]
        if len(frames) != len(pts_list):
            raise ValueError("one array of peak coordinates per frame is required")
        if not frames:
            return []
        h, w = frames[0].shape
        for f in frames:
            if f.ndim != 2 or f.shape != (h, w):
                raise ValueError("all frames of a batch must be 2D arrays of one shape")
        pts = [np.ascontiguousarray(np.asarray(q, dtype=np.float64).reshape(-1, 2)) for q in pts_list]
        counts = np.array([len(q) for q in pts], dtype=np.int64)
        flat = np.ascontiguousarray(np.concatenate(pts)) if counts.sum() else np.zeros((0, 2))
        n_c = lib.zb200_num_complex_modes(self.n_max)
        shape = (int(counts.sum()), len(self.n) if kind == "real" else n_c)
        dtype = np.complex128 if kind == "complex" else np.float64
        if out is None:
            out = np.empty(shape, dtype=dtype)
        else:
            if (not isinstance(out, np.ndarray) or out.dtype != dtype or out.ndim != 2 or out.shape[1] != shape[1]
                    or out.shape[0] < shape[0] or not out.flags.c_contiguous):
                raise ValueError(f"out must be a C-contiguous {np.dtype(dtype).name} array of shape (>= {shape[0]}, {shape[1]})")
            out = out[:shape[0]]
        ptrs = (C.c_void_p * len(frames))(*[f.ctypes.data for f in frames])
        prec = self._precision_code(host_route=True)
        _lib.check(lib.zb200_project_peaks_host(self._plan, ptrs, len(frames), int(h), int(w), np_ptr(flat), np_ptr(counts),
                                                prec, code, _lib.F64, np_ptr(out)), "project_peaks_host")
        edges = np.concatenate([[0], np.cumsum(counts)])
        parts = [out[a:b] for a, b in zip(edges[:-1], edges[1:])]
        if kind == "real":
            return [zmoments(data=q, n=self.n, m=self.m, patch_size=self.size) for q in parts]
        return parts

    # ------------------------------------------------------------ K4: dense map
    def _image_on_device(self, image):
        torch = _lib.require_cuda()
        if is_torch(image):
            return image.to(device="cuda", dtype=torch.float32).contiguous()
        return torch.from_numpy(np.ascontiguousarray(image, dtype=np.float32)).cuda()

    def _transform_map(self, image, row0: int = 0, rows: Optional[int] = None) -> zmoments:
        self._validate_size(image)
        torch = _lib.require_cuda()
        lib = _lib.load()
        out = self._map_device(self._image_on_device(image), row0, rows)
        data = _host_f64(out) if self._want_host(image) else out
        return zmoments(data=data, n=self.n, m=self.m, patch_size=self.size)

    def _map_device(self, dev, row0: int = 0, rows: Optional[int] = None):
        """Moment maps (M, rows, W) of a CUDA frame as a CUDA tensor (internal)."""
        torch = _lib.require_cuda()
        h, w = int(dev.shape[0]), int(dev.shape[1])
        rows = h - row0 if rows is None else rows
        out = torch.empty((len(self.n), rows, w), dtype=torch.float32, device=dev.device)
        _lib.check(_lib.load().zb200_moment_map_f32(self._plan, int(dev.data_ptr()), h, w, row0, rows,
                                                    self._precision_code(for_map=True), int(out.data_ptr()), self._stream()),
                   "moment_map")
        return out

    def mirror_map(self, image, theta=None, p=2, m_unselect=(0, 1), band_rows: int = 64):
        """``transform(image).mirror_map(theta, p, m_unselect)`` (mtflearn/features/_zmoments.py:464-493) without the
        (M, H, W) moment maps ever existing in full: the frame is processed in bands of ``band_rows`` output rows --
        dense moment map of the band (K4, bit-identical to the corresponding rows of the full map), then the
        mirror-score kernel on that band -- so the working set is ``M * band_rows * W`` floats (95 MB at W = 4096)
        instead of 6.1 GB for a 4096^2 frame.  Returns (H, W)."""
        if image.ndim != 2:
            raise ValueError("Images must be 2D or 3D array.")
        self._validate_size(image)
        torch = _lib.require_cuda()
        dev = self._image_on_device(image)
        h, w = int(dev.shape[0]), int(dev.shape[1])
        band_rows = max(2, int(band_rows) & ~1)                       # even: the map kernel pairs rows by absolute parity
        out = torch.empty((h, w), dtype=torch.float32, device=dev.device)
        for r0 in range(0, h, band_rows):
            rows = min(band_rows, h - r0)
            band = zmoments(data=self._map_device(dev, r0, rows), n=self.n, m=self.m, patch_size=self.size)
            out[r0:r0 + rows] = band.mirror_map(theta=theta, p=p, m_unselect=m_unselect)
        return _host_f64(out) if self._want_host(image) else out

    def symmetry_map(self, image, n_folds, p=2, m_unselect=None, row0: int = 0, rows: Optional[int] = None):
        """Fused ``transform(image).rot_maps(n_folds, p, m_unselect)`` that never writes the
        (M,H,W) moment maps to HBM.  Returns (F, rows, W); ``row0/rows`` select a row band
        (image-tile sharding across GPUs)."""
        if image.ndim != 2:
            raise ValueError("Images must be 2D or 3D array.")
        self._validate_size(image)
        if m_unselect is None:
            m_unselect = (0, 1)
        elif 0 not in m_unselect:
            raise ValueError("m=0 must be included in m_unselect.")
        torch = _lib.require_cuda()
        lib = _lib.load()
        wts, sel = rot_weight_tables(self.m, n_folds, m_unselect)
        if not is_torch(image) and self._want_host(image) and row0 == 0 and (rows is None or rows == image.shape[0]):
            # the numpy user's call: frame up once, the map in row bands, each band's scores downloaded and widened
            # to float64 while the next band is computed (zb200_symmetry_map_host)
            src = np.ascontiguousarray(image, dtype=np.float32)
            h, w = src.shape
            w32 = np.ascontiguousarray(wts, dtype=np.float32)
            host = np.empty((w32.shape[0], h, w), dtype=np.float64)
            _lib.check(lib.zb200_symmetry_map_host(self._plan, np_ptr(src), int(h), int(w), self._precision_code(for_map=True),
                                                   np_ptr(w32), np_ptr(sel), w32.shape[0], norm_code(p), np_ptr(host)),
                       "symmetry_map_host")
            return host
        dev = self._image_on_device(image)
        h, w = int(dev.shape[0]), int(dev.shape[1])
        rows = h - row0 if rows is None else rows
        out = torch.empty((wts.shape[0], rows, w), dtype=torch.float32, device=dev.device)
        wts = np.ascontiguousarray(wts, dtype=np.float32)
        _lib.check(lib.zb200_symmetry_map_f32(self._plan, int(dev.data_ptr()), h, w, row0, rows,
                                              self._precision_code(for_map=True), np_ptr(wts), np_ptr(sel),
                                              wts.shape[0], norm_code(p), int(out.data_ptr()), self._stream()),
                   "symmetry_map")
        return _host_f64(out) if self._want_host(image) else out
