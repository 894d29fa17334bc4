"""BASELINE config 5 per frame, all on the device: peaks (local_max) -> clear_border -> gather -> ZPs(12,64) |Zc|."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from motif_learn_b200.datasets import honeycomb_image
from motif_learn_b200.features import ZPs, KeyPoints, local_max
img, pts = honeycomb_image(2048, bond=12.0, seed=0, jitter=0.3, noise=0.01)
dimg = torch.from_numpy(img).cuda()
z = ZPs(12, 64)
def wall(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, out
ms_p, peaks = wall(lambda: local_max(dimg, 5.0, 0.3))
ms_pt, peaks_t = wall(lambda: local_max(dimg, 5.0, 0.3, as_tensor=True))
kp = KeyPoints(peaks, dimg, 64)
ms_g, patches = wall(lambda: kp.extract_patches())
ms_f, feats = wall(lambda: z.transform_features(patches, "abs"))
ms_all, _ = wall(lambda: z.transform_features(KeyPoints(local_max(dimg, 5.0, 0.3), dimg, 64).extract_patches(), "abs"))
print(f"peaks {len(peaks)} (truth {len(pts)}): local_max {ms_p:.2f} ms (device result {ms_pt:.2f} ms), gather {ms_g:.3f} ms, "
      f"features {ms_f:.3f} ms, whole frame {ms_all:.2f} ms -> {1e3 / ms_all:.0f} frames/s per GPU", flush=True)
