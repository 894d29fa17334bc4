"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from motif_learn_b200.datasets import honeycomb_image
from motif_learn_b200.features import ZPs, KeyPoints, local_max
img, pts = honeycomb_image((200, 333), bond=12.0, seed=1, jitter=0.2, noise=0.01)
dimg = torch.from_numpy(img).cuda()
for prec in ("f16x3", "f16", "tf32x3", "fp32"):
    z = ZPs(12, 48, precision=prec)
    s = z.symmetry_map(dimg, [2, 3, 4, 6], row0=3, rows=101)
    m = z._transform_map(dimg, row0=10, rows=7).data
    print(prec, float(s.nansum()), float(m.sum()))
z = ZPs(8, 33)
print("odd window", float(z.symmetry_map(dimg, [3, 6]).nansum()))
p = local_max(img, 5.0, 0.3)
kp = KeyPoints(p, dimg, 32)
patches = kp.extract_patches()
for prec in ("tf32x3", "tf32", "fp32"):
    print(prec, float(ZPs(10, 32, precision=prec).transform(patches).data.sum()))
zz = ZPs(12, 64)
kp64 = KeyPoints(p, dimg, 64)
print("fused gather", float(zz.transform_peaks(dimg, kp64.pts, fused=True).data.sum()))
print("features", float(zz.transform_features(kp64.extract_patches(), "abs").sum()))
torch.cuda.synchronize()
print("ok")
