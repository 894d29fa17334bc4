"""mirror_map of a dense moment map (a17): time of the score kernel chain on a 1024^2 frame."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from motif_learn_b200.datasets import honeycomb_image
from motif_learn_b200.features import ZPs
img = torch.from_numpy(honeycomb_image(1024, bond=12.0, seed=0)[0]).cuda()
z = ZPs(12, 48)
zm = z.transform(img)
for _ in range(2): out = zm.mirror_map()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3): out = zm.mirror_map()
torch.cuda.synchronize()
print("mirror_map 1024^2 n_max=12:", round((time.perf_counter() - t0) / 3 * 1e3, 2), "ms", float(out.max()), flush=True)
