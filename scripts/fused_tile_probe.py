"""Gathering folded kernel: time vs number of windows (how much a full / nearly empty 128-row tile costs).
Run under ncu --metrics gpu__time_duration.sum --cache-control none and read the project_fold launches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from motif_learn_b200.datasets import honeycomb_image
from motif_learn_b200.features import ZPs, clear_border
img, pts = honeycomb_image(2048, bond=12.0, seed=0)
dimg = torch.from_numpy(img).cuda()
kept = clear_border(pts, img.shape, 64)
z = ZPs(12, 64)
for count in (1924, 9472, 18944, 20900):
    for _ in range(4):
        z.transform_peaks(dimg, kept[:count], "abs", fused=True)
torch.cuda.synchronize()
