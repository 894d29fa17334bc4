"""Worst element of each projection shape against the strict parity gate allclose(rtol=1e-4, atol=1e-6*max|ref|):
ratio = |err| / (atol + rtol*|ref|) (must be <= 1), plus max|err|/max|ref| and which mode holds the worst element."""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import zernike_oracle as zo
from motif_learn_b200.features import ZPs
shapes = [(12, 64, 1000), (20, 64, 300), (8, 33, 257), (3, 6, 5), (12, 48, 129), (0, 4, 3), (21, 64, 130), (23, 64, 100), (10, 32, 1000)]
for n_max, size, count in shapes:
    rng = np.random.default_rng(n_max * 100 + size)
    patches = rng.random((count, size, size), dtype=np.float32)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        n, m, v = zo.zernike_basis(n_max, size)
        ref = zo.project_patches(patches.astype(np.float64), v)
        exact = None
        for prec in ("fp32", "tf32x3"):
            try:
                z = ZPs(n_max, size, precision=prec)
                got = z.transform(patches).data
            except Exception as exc:
                print(n_max, size, prec, "unsupported:", str(exc)[:60]); continue
            # the same contraction with OUR fp64 basis: separates basis differences from arithmetic error
            own = zo.project_patches(patches.astype(np.float64), z.polynomials)
            for label, r in (("vs reference basis", ref), ("vs own fp64 basis", own)):
                err = np.abs(got - r)
                tol = 1e-6 * np.abs(r).max() + 1e-4 * np.abs(r)
                i = np.unravel_index(np.argmax(err / tol), err.shape)
                print(f"n_max={n_max:2d} k={size:2d} {prec:6s} {label:20s}: worst ratio {float((err/tol).max()):6.2f} at mode j={i[1]} (n={n[i[1]]},m={m[i[1]]}) "
                      f"|ref|={abs(r[i]):.3e} err={err[i]:.3e}; max err/max|ref| = {err.max()/np.abs(r).max():.2e}", flush=True)
