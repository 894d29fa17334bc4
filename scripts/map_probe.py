"""On-GPU probe of the tcgen05 dense map: accuracy vs the SIMT kernel / oracle on a small frame and
timings at BASELINE config-2 / config-4 sizes.  Run under `timeout`."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import zernike_oracle as zo
from motif_learn_b200.datasets import honeycomb_image
from motif_learn_b200.features import ZPs

def check(n_max, k, shape, seed=0):
    img, _ = honeycomb_image(shape, bond=12.0, seed=seed, angle=7.0)
    n, m, v = zo.zernike_basis(n_max, k)
    ref = zo.moment_map_fft(img.astype(np.float64), v, n)
    rot = zo.rot_maps(ref, n, m, [2, 3, 4, 6])
    out = {}
    for prec in ("fp32", "tf32", "tf32x3"):
        z = ZPs(n_max, k, precision=prec)
        got = z.transform(img).data
        sc = z.symmetry_map(img, [2, 3, 4, 6])
        out[prec] = (float(np.abs(got - ref).max() / np.abs(ref).max()), float(np.nanmax(np.abs(sc - rot))))
    print(f"check n_max={n_max} k={k} shape={shape}: (moment err / max, score abs err) {out}", flush=True)

def timing(n_max, k, size, folds=(2, 3, 4, 6)):
    img = torch.rand((size, size), device="cuda")
    res = {}
    for prec in ("fp32", "tf32", "tf32x3"):
        if prec == "fp32" and size > 2048: continue
        z = ZPs(n_max, k, precision=prec)
        for _ in range(2): z.symmetry_map(img, list(folds))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): z.symmetry_map(img, list(folds))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        tf = 2.0 * size * size * k * k * len(z.n) / ms / 1e9
        res[prec] = {"ms": round(ms, 3), "Mpix_s": round(size * size / ms / 1e3, 1), "TFLOPs": round(tf, 1)}
    print(f"timing n_max={n_max} k={k} {size}x{size}: {json.dumps(res)}", flush=True)

if __name__ == "__main__":
    torch.cuda.set_device(0)
    check(12, 48, (100, 300))
    check(12, 48, (160, 700), seed=2)
    check(8, 16, (70, 530), seed=3)
    check(12, 64, (130, 1100), seed=4)
    timing(12, 48, 2048)
    timing(12, 64, 4096)
