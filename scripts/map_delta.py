"""Debug aid: compares the tcgen05 dense map with the SIMT map on dense random frames and prints
where (rows / columns / pixel phase / mode) the two disagree."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from motif_learn_b200.features import ZPs
np.set_printoptions(linewidth=250, precision=3, suppress=True)
def run(n_max, k, H, W, prec="tf32"):
    rng = np.random.default_rng(0)
    img = rng.random((H, W), dtype=np.float32)
    ref = ZPs(n_max, k, precision="fp32").transform(img).data
    got = ZPs(n_max, k, precision=prec).transform(img).data
    err = np.abs(got - ref) / np.abs(ref).max()
    print(f"n_max={n_max} k={k} {H}x{W} {prec}: max rel err {err.max():.3e}")
    if err.max() > 1e-2:
        print("  per mode (first 12):", err.max(axis=(1, 2))[:12])
        print("  per row (first 12):", err.max(axis=(0, 2))[:12])
        print("  per phase:", [float(err[:, :, p::4].max()) for p in range(4)])
        cols = err.max(axis=(0, 1))
        bad = np.nonzero(cols > 1e-2)[0]
        print("  bad cols:", bad[:10], "...", bad[-10:], "count", len(bad))
run(4, 16, 40, 600)
run(8, 16, 40, 600)
run(4, 16, 200, 600)
run(4, 48, 60, 600)
run(12, 48, 60, 600)
run(12, 48, 60, 600, "tf32x3")
