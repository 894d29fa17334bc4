"""Folded projection: worst element against the strict gate and kernel time, per accumulation-chunk length
(ZB200_EXPERIMENT=1 ZB200_TC_CHUNK=<super-blocks per chunk>; default 8).  Lattice patches (the bench's batch) and
uniform random patches."""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import zernike_oracle as zo
from motif_learn_b200.features import ZPs
from motif_learn_b200.datasets import honeycomb_image
from motif_learn_b200.features import clear_border, KeyPoints
img, pts = honeycomb_image(1024, bond=12.0, seed=3)
kept = clear_border(pts, img.shape, 64)[:3000]
lattice = KeyPoints(kept, torch.from_numpy(img).cuda(), 64).extract_patches()
rng = np.random.default_rng(0)
uniform = torch.from_numpy(rng.random((3000, 64, 64), dtype=np.float32)).cuda()
big = torch.from_numpy(rng.random((262144, 64, 64), dtype=np.float32)).cuda()
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for n_max in (12, 20):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        n, m, v = zo.zernike_basis(n_max, 64)
    z = ZPs(n_max, 64, value_max=1.0)
    for name, x in (("lattice", lattice), ("uniform", uniform)):
        ref = zo.project_patches(x.cpu().numpy().astype(np.float64), v)
        got = z.transform(x).data.double().cpu().numpy()
        err = np.abs(got - ref); tol = 1e-6 * np.abs(ref).max() + 1e-4 * np.abs(ref)
        i = np.unravel_index(np.argmax(err / tol), err.shape)
        print(f"chunk={os.environ.get('ZB200_TC_CHUNK', 'default')} n_max={n_max} {name}: worst ratio {(err/tol).max():.3f} at (n={n[i[1]]},m={m[i[1]]}), "
              f"max err/max = {err.max()/np.abs(ref).max():.2e}", flush=True)
    t = timeit(lambda: z.transform(big))
    print(f"chunk={os.environ.get('ZB200_TC_CHUNK', 'default')} n_max={n_max}: 262144 patches in {t:.4f} ms", flush=True)
