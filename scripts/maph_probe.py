"""On-GPU probe of the fp16-split tcgen05 dense map (zb200_map_h.cu): accuracy vs the oracle on small
frames (even/odd/non-multiple-of-16 windows, scaled frames) and timings at BASELINE config-2 / config-4
sizes.  Run under `timeout`."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import zernike_oracle as zo
from motif_learn_b200.datasets import honeycomb_image
from motif_learn_b200.features import ZPs

PRECS = tuple(os.environ.get("PROBE_PRECS", "tf32x3,f16x3,f16").split(","))

def check(n_max, k, shape, seed=0, gain=1.0):
    img, _ = honeycomb_image(shape, bond=12.0, seed=seed, angle=7.0)
    img = (img * gain).astype(np.float32)
    n, m, v = zo.zernike_basis(n_max, k)
    ref = zo.moment_map_fft(img.astype(np.float64), v, n)
    rot = zo.rot_maps(ref, n, m, [2, 3, 4, 6])
    out = {}
    for prec in PRECS:
        z = ZPs(n_max, k, precision=prec)
        try:
            got = z.transform(img).data
            sc = z.symmetry_map(img, [2, 3, 4, 6])
        except Exception as e:                       # unsupported shape for this kernel
            out[prec] = str(e)[:60]
            continue
        err = np.abs(got - ref)
        out[prec] = (float(err.max() / np.abs(ref).max()), float(np.nanmax(np.abs(sc - rot))),
                     bool(np.allclose(got, ref, rtol=1e-4, atol=1e-6 * np.abs(ref).max())))
    print(f"check n_max={n_max} k={k} shape={shape} gain={gain}: (moment err/max, score abs err, fp32_close) {out}", flush=True)

def timing(n_max, k, size, folds=(2, 3, 4, 6), lattice=True):
    if lattice:
        img = torch.from_numpy(honeycomb_image(size, bond=12.0, seed=0)[0]).cuda()
    else:
        img = torch.rand((size, size), device="cuda")
    scratch = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    res = {}
    for prec in PRECS:
        z = ZPs(n_max, k, precision=prec)
        for _ in range(2): z.symmetry_map(img, list(folds))
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(5):
            scratch.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); z.symmetry_map(img, list(folds)); e1.record(); torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        ms = tot / 5
        tf = 2.0 * size * size * k * k * len(z.n) / ms / 1e9
        res[prec] = {"ms": round(ms, 3), "Mpix_s": round(size * size / ms / 1e3, 1), "TFLOPs": round(tf, 1)}
    print(f"timing n_max={n_max} k={k} {size}x{size}: {json.dumps(res)}", flush=True)

if __name__ == "__main__":
    torch.cuda.set_device(0)
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "check"):
        check(12, 48, (100, 300))
        check(12, 48, (160, 700), seed=2)
        check(8, 16, (70, 530), seed=3)
        check(12, 64, (130, 1100), seed=4)
        check(8, 33, (90, 200), seed=5)
        check(10, 40, (90, 515), seed=6)
        check(6, 20, (150, 203), seed=7)
        check(12, 48, (100, 300), gain=3.7e5)
        check(12, 48, (100, 300), gain=2.1e-7)
    if what in ("all", "time"):
        timing(12, 48, 2048)
        timing(12, 64, 4096)
