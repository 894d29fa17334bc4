"""Times ZPs.transform(frame) (moment maps written to HBM) against symmetry_map (scores only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from motif_learn_b200.datasets import honeycomb_image
from motif_learn_b200.features import ZPs
def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n, 3)
for size, k in ((2048, 48), (4096, 64)):
    img = torch.from_numpy(honeycomb_image(size, bond=12.0, seed=0)[0]).cuda()
    for prec in sys.argv[1:] or ["f16x3"]:
        z = ZPs(12, k, precision=prec)
        out_gb = 91 * size * size * 4 / 1e9
        ms_m = t(lambda: z.transform(img))
        ms_s = t(lambda: z.symmetry_map(img, [2, 3, 4, 6]))
        print(f"{size}^2 k={k} {prec}: moments {ms_m} ms ({out_gb:.2f} GB out -> {out_gb / ms_m * 1e3:.0f} GB/s incl. compute), scores {ms_s} ms", flush=True)
