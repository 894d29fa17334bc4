"""Where the time of ZPs.symmetry_map(numpy) goes (ZB200_HOST_TRACE=2 prints the band pipeline's timeline)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from motif_learn_b200.datasets import honeycomb_image
from motif_learn_b200.features import ZPs
img, _ = honeycomb_image(2048, bond=12.0, seed=0)
pin = torch.empty((2048, 2048), dtype=torch.float32, pin_memory=True); pin.copy_(torch.from_numpy(img))
z = ZPs(12, 48, output="numpy")
for name, frame in (("pinned", pin.numpy()), ("pageable", img)):
    z.symmetry_map(frame, [2, 3, 4, 6])
    t0 = time.perf_counter()
    for _ in range(3): r = z.symmetry_map(frame, [2, 3, 4, 6])
    dt = (time.perf_counter() - t0) / 3
    print(f"{name}: {dt*1e3:.2f} ms per 2048^2 frame -> {2048*2048/dt/1e6:.0f} Mpix/s", flush=True)
d = torch.from_numpy(img).cuda()
zd = ZPs(12, 48)
zd.symmetry_map(d, [2, 3, 4, 6]); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3): zd.symmetry_map(d, [2, 3, 4, 6])
torch.cuda.synchronize()
print(f"device-resident: {(time.perf_counter()-t0)/3*1e3:.2f} ms")
t0 = time.perf_counter(); a = np.empty((4, 2048, 2048)); a[:] = 1.0; print(f"first touch of a fresh 134 MB array by one thread: {(time.perf_counter()-t0)*1e3:.1f} ms")
