"""Where the time of the host frame route goes (ZB200_HOST_TRACE=1 prints the pipeline's own breakdown)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from motif_learn_b200.datasets import honeycomb_image
from motif_learn_b200.features import ZPs, clear_border
print("THP:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip(), "cpus", os.cpu_count(), flush=True)
img, pts = honeycomb_image(1024, bond=12.0, seed=0)
tile = np.tile(img, (2, 2))
kept = np.concatenate([clear_border(pts + np.array([dx, dy]), tile.shape, 64) for dx in (0, 1024) for dy in (0, 1024)])
nf = 8
pinned = torch.empty((nf, 2048, 2048), dtype=torch.float32, pin_memory=True)
for f in range(nf): pinned[f].copy_(torch.from_numpy(tile))
frames = [pinned[f].numpy() for f in range(nf)]
pageable = [np.array(f) for f in frames]
pts_list = [kept] * nf
z = ZPs(12, 64, output="numpy")
n = len(kept) * nf
def run(label, fn, reps=4):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    dt = (time.perf_counter() - t0) / reps
    print(f"{label}: {dt*1e3:.2f} ms per call of {nf} frames -> {n/dt/1e6:.1f} M patches/s", flush=True)
only = os.environ.get("PROBE_ONLY", "")
if only == "abs":
    bufa = np.zeros((n, 49))
    run("pinned frames, abs features, reused out (first)", lambda: z.transform_peaks_batch(frames, pts_list, "abs", out=bufa))
    buf = np.zeros((n, 91))
    run("pinned frames, reused out", lambda: z.transform_peaks_batch(frames, pts_list, out=buf))
    run("pinned frames, abs features, reused out (again)", lambda: z.transform_peaks_batch(frames, pts_list, "abs", out=bufa))
    sys.exit(0)
run("pinned frames, fresh out", lambda: z.transform_peaks_batch(frames, pts_list))
buf = np.zeros((n, 91))
run("pinned frames, reused out", lambda: z.transform_peaks_batch(frames, pts_list, out=buf))
run("pageable frames, reused out", lambda: z.transform_peaks_batch(pageable, pts_list, out=buf))
bufa = np.zeros((n, 49))
run("pinned frames, abs features, reused out", lambda: z.transform_peaks_batch(frames, pts_list, "abs", out=bufa))
d = torch.from_numpy(tile).cuda()
def dev():
    for f in range(nf): z.transform_peaks(d, kept)
    torch.cuda.synchronize()
zd = ZPs(12, 64)
def dev2():
    for f in range(nf): zd.transform_peaks(d, kept)
    torch.cuda.synchronize()
run("device-resident frames (torch out)", dev2)
