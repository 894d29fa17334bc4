"""BASELINE config 3 shape: complex ZPs n_max=20 on 64x64 patches (262144-patch slice of the 1M stack)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from motif_learn_b200.features import ZPs
N = 262144
x = torch.rand((N, 64, 64), device="cuda")
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for prec in ("tf32x3", "tf32"):
    z = ZPs(20, 64, precision=prec)
    ms_r = t(lambda: z.transform(x))
    ms_c = t(lambda: z.transform_features(x, "complex"))
    ms_a = t(lambda: z.transform_features(x, "abs"))
    fl = 2.0 * N * 4096 * 231
    print(f"{prec}: real {ms_r:.3f} ms ({N/ms_r/1e3:.1f} M patches/s, {fl/ms_r/1e9:.0f} TFLOP/s, {N*(16384+924)/ms_r/1e6:.0f} GB/s)  "
          f"complex {ms_c:.3f} ms  abs {ms_a:.3f} ms", flush=True)
