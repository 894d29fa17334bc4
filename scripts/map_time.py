"""Times the tcgen05 symmetry map under different conditions (random vs lattice frame, L2 flush)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from motif_learn_b200.datasets import honeycomb_image
from motif_learn_b200.features import ZPs
def t(img, prec, flush, n=8):
    z = ZPs(12, 48, precision=prec)
    scratch = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(2): z.symmetry_map(img, [2, 3, 4, 6])
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        if flush: scratch.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); z.symmetry_map(img, [2, 3, 4, 6]); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return round(tot / n, 3)
rnd = torch.rand((2048, 2048), device="cuda")
lat = torch.from_numpy(honeycomb_image(2048, bond=12.0, seed=0)[0]).cuda()
for name, img in (("random", rnd), ("lattice", lat)):
    for prec in ("tf32", "tf32x3"):
        print(name, prec, "no flush", t(img, prec, False), "ms | flush", t(img, prec, True), "ms", flush=True)
