"""Ablation timings of map_h_kernel through ZB200_MAP_DEBUG bits (1 no frame TMA, 2 no MMA, 4 no score
phase, 8 no TMEM drain, 16 no basis TMA).  Results are wrong when a bit is set; timings only."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from motif_learn_b200.datasets import honeycomb_image
from motif_learn_b200.features import ZPs
img = torch.from_numpy(honeycomb_image(2048, bond=12.0, seed=0)[0]).cuda()
def t(prec, n=5):
    z = ZPs(12, 48, precision=prec)
    for _ in range(2): z.symmetry_map(img, [2, 3, 4, 6])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): z.symmetry_map(img, [2, 3, 4, 6])
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n, 3)
for env in sys.argv[1:]:
    for kv in env.split(","):
        k, v = kv.split("=")
        os.environ[k] = v
    print(env, {p: t(p) for p in ("f16x3", "f16")}, flush=True)
