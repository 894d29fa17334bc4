"""Same-box A/B of the accumulator chunk length of the fp32-grade projection: time and error vs the oracle.
The knobs are read once per process and only with ZB200_EXPERIMENT=1, so every setting is its own process:
    for c in 4 8 16 32; do ZB200_EXPERIMENT=1 ZB200_TC_CHUNK=$c python scripts/tc_chunk_ab.py; done"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import zernike_oracle as zo
from motif_learn_b200.datasets import honeycomb_image
from motif_learn_b200.features import ZPs, KeyPoints
img, pts = honeycomb_image(2048, bond=12.0, seed=0)
base = KeyPoints(pts, torch.from_numpy(img).cuda(), 64).extract_patches()
x = base.repeat(13, 1, 1)[:262144].contiguous()
n, m, v = zo.zernike_basis(12, 64)
sub = base[:2048].cpu().numpy()
ref = zo.project_patches(sub.astype(np.float64), v)
z = ZPs(12, 64)
def t(nrep=100):
    for _ in range(3): z.transform(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(nrep): z.transform(x)
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / nrep, 4)
chunk = os.environ.get("ZB200_TC_CHUNK", "8 (default)") if os.environ.get("ZB200_EXPERIMENT") == "1" else "8 (default; knobs off)"
got = z.transform(base[:2048].contiguous()).data.cpu().numpy()
err = np.abs(got - ref)
ok = np.allclose(got, ref, rtol=1e-4, atol=1e-6 * np.abs(ref).max())
worst = (err - 1e-4 * np.abs(ref)).max() / np.abs(ref).max()
print("chunk", chunk, "ms", t(), "max err/max", err.max() / np.abs(ref).max(), "fp32_close", ok, "worst (err - rtol|ref|)/max", worst, flush=True)
