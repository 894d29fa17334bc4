"""Why is the host widening erratic inside bench.py?  Same call under different process conditions."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
from motif_learn_b200.datasets import honeycomb_image
from motif_learn_b200.features import ZPs, clear_border
img, pts = honeycomb_image(1024, bond=12.0, seed=0)
tile = np.tile(img, (2, 2))
kept = np.concatenate([clear_border(pts + np.array([dx, dy]), tile.shape, 64) for dx in (0, 1024) for dy in (0, 1024)])
nf = 16
pinned = torch.empty((nf, 2048, 2048), dtype=torch.float32, pin_memory=True)
for f in range(nf): pinned[f].copy_(torch.from_numpy(tile))
frames = [pinned[f].numpy() for f in range(nf)]
pts_list = [kept] * nf
z = ZPs(12, 64, output="numpy")
n = len(kept) * nf
buf = np.zeros((n, 91))
def run(label, reps=6):
    z.transform_peaks_batch(frames, pts_list, out=buf)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); z.transform_peaks_batch(frames, pts_list, out=buf); ts.append((time.perf_counter() - t0) * 1e3)
    print(f"{label}: ms per call {[round(t, 1) for t in ts]} -> best {n/min(ts)/1e3:.1f} M/s, mean {n*len(ts)/sum(ts)/1e3:.1f} M/s", flush=True)
run("baseline")
a = np.random.rand(4096, 4096); b = np.random.rand(4096, 512)
for _ in range(3): a @ b
run("right after numpy.dot (BLAS threads awake)")
time.sleep(1.0)
run("1 s later")
from threadpoolctl import threadpool_info, threadpool_limits
print([(p["internal_api"], p.get("threading_layer"), p["num_threads"]) for p in threadpool_info()], flush=True)
for _ in range(3): a @ b
with threadpool_limits(limits=1):
    run("after numpy.dot, BLAS limited to 1 thread during the calls")
x = torch.rand((262144, 64, 64), device="cuda"); zd = ZPs(12, 64)
for _ in range(3000): zd.transform(x)
torch.cuda.synchronize()
run("after 3000 projection launches (GPU power capped)")
import pynvml; pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
for _ in range(1000): pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
run("after NVML polling")
