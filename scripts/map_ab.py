"""Same-box A/B of an environment toggle on the fused symmetry map (plans and kernels read the variable per call)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from motif_learn_b200.datasets import honeycomb_image
from motif_learn_b200.features import ZPs
var = sys.argv[1]
def t(z, img, n):
    for _ in range(3): z.symmetry_map(img, [2, 3, 4, 6])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): z.symmetry_map(img, [2, 3, 4, 6])
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n, 3)
for size, k, n in ((2048, 48, 40), (4096, 64, 12)):
    img = torch.from_numpy(honeycomb_image(size, bond=12.0, seed=0)[0]).cuda()
    z = ZPs(12, k)
    res = []
    for rep in range(3):
        for val in ("1", "0"):
            os.environ[var] = val
            res.append((val, t(z, img, n)))
    print(size, k, var, res, flush=True)
