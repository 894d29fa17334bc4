"""Throughput of the fused gather->projection kernel vs gather + projection (BASELINE config 5 shape:
64-px windows at peak positions of 2048^2 frames, n_max=12)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from motif_learn_b200.datasets import honeycomb_image
from motif_learn_b200.features import ZPs, KeyPoints, clear_border
img, pts = honeycomb_image(2048, bond=12.0, seed=0)
dimg = torch.from_numpy(img).cuda()
kept = clear_border(pts, img.shape, 64)
rng = np.random.default_rng(0)
many = np.concatenate([kept + rng.normal(0, 2.0, kept.shape) for _ in range(13)])[:262144]
many = clear_border(many, img.shape, 64)
z = ZPs(12, 64)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, p in (("one frame", kept), ("262k peaks", many)):
    kp = KeyPoints(p, dimg, 64)
    t_unf = timeit(lambda: z.transform(kp.extract_patches()))
    t_fus = timeit(lambda: z.transform_peaks(dimg, p, fused=True))
    t_g = timeit(lambda: kp.extract_patches())
    print(f"{name}: N={len(p)}  gather+project {t_unf:.3f} ms ({len(p)/t_unf/1e3:.1f} M patches/s)   fused {t_fus:.3f} ms ({len(p)/t_fus/1e3:.1f} M patches/s)   gather alone {t_g:.3f} ms ({len(p)*2*16384/t_g/1e6:.0f} GB/s r+w)", flush=True)
