"""Fused gather -> folded projection: one 2048^2 frame (20.9 k windows) and 262 k windows, per epilogue."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from motif_learn_b200.datasets import honeycomb_image
from motif_learn_b200.features import ZPs, clear_border
img, pts = honeycomb_image(2048, bond=12.0, seed=0)
dimg = torch.from_numpy(img).cuda()
kept = clear_border(pts, img.shape, 64)
rng = np.random.default_rng(0)
many = clear_border(np.concatenate([kept + rng.normal(0, 2.0, kept.shape) for _ in range(13)])[:262144], img.shape, 64)
z = ZPs(12, 64)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, p in (("one frame", kept), ("262k", many)):
    for kind in ("real", "abs", "complex"):
        t = timeit(lambda: z.transform_peaks(dimg, p, kind, fused=True))
        print(f"chunk={os.environ.get('ZB200_TC_CHUNK','default')} {name} N={len(p)} {kind}: fused {t:.3f} ms ({len(p)/t/1e3:.1f} M/s)", flush=True)
