"""Quick on-GPU probe of the tcgen05 projection: correctness vs the SIMT kernel and the oracle for a
few shapes, then kernel timings per precision.  Run under `timeout` (a pipeline bug would hang)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import zernike_oracle as zo
from motif_learn_b200 import _lib
from motif_learn_b200.features import ZPs

def run(n_max, size, count, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.random((count, size, size), dtype=np.float32)
    dx = torch.from_numpy(x).cuda()
    _, _, v = zo.zernike_basis(n_max, size)
    ref = zo.project_patches(x.astype(np.float64), v)
    scale = np.abs(ref).max()
    res = {}
    for prec in ("fp32", "tf32", "tf32x3"):
        z = ZPs(n_max, size, precision=prec)
        if prec != "fp32" and not _lib.load().zb200_plan_supports(z._plan, _lib.PRECISIONS[prec], _lib.OUT_REAL):
            res[prec] = "unsupported"; continue
        got = z.transform(dx).data.cpu().numpy()
        res[prec] = float(np.abs(got - ref).max() / scale)
    print(f"n_max={n_max} size={size} N={count}: max|err|/max|ref| {res}", flush=True)

def timing(n_max, size, count):
    dx = torch.rand((count, size, size), device="cuda")
    out = {}
    for prec in ("fp32", "tf32", "tf32x3"):
        z = ZPs(n_max, size, precision=prec)
        for _ in range(3): z.transform(dx)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): z.transform(dx)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        gbs = count * (size * size * 4 + len(z.n) * 4) / ms / 1e6
        out[prec] = {"ms": round(ms, 4), "Mpatch_s": round(count / ms / 1e3, 1), "GBs": round(gbs, 1)}
    print(f"timing n_max={n_max} size={size} N={count}: {json.dumps(out)}", flush=True)

if __name__ == "__main__":
    torch.cuda.set_device(0)
    for args in [(12, 64, 128), (12, 64, 1000), (10, 32, 5000), (20, 64, 700), (12, 48, 40000), (4, 8, 3), (21, 64, 300)]:
        run(*args)
    timing(12, 64, 262144)
    timing(20, 64, 131072)
    timing(10, 32, 262144)
