"""One launch of the fused gather -> folded projection on 262 k windows of a 2048^2 frame (for ncu)."""
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
many = np.concatenate([kept + rng.normal(0, 2.0, kept.shape) for _ in range(13)])[:262144]
many = clear_border(many, img.shape, 64)
z = ZPs(12, 64)
for _ in range(3):
    out = z.transform_peaks(dimg, many, fused=True)
torch.cuda.synchronize()
print("ok", out.data.shape)
