"""Measured denominators that MEASURED_PEAKS.json does not carry (SURVEY.md 8d asks for them): dense TF32 and
FP32-SIMT matmul rates and the SUSTAINED copy bandwidth, taken the way the driver takes its own numbers
(torch.matmul 8192^3, best of 10 = burst; back to back for 4 s = sustained; b.copy_(a) over 1 Gi bf16 elements).
Writes profiles/peaks_tf32.json; bench.py reads it for the tf32 rooflines."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def timed(fn, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 1e3 / reps

def matmul_rates(dtype, allow_tf32):
    torch.backends.cuda.matmul.allow_tf32 = allow_tf32
    n = 8192
    a = torch.randn((n, n), device="cuda", dtype=dtype); b = torch.randn((n, n), device="cuda", dtype=dtype)
    c = torch.empty((n, n), device="cuda", dtype=dtype)
    f = lambda: torch.matmul(a, b, out=c)
    for _ in range(3): f()
    torch.cuda.synchronize()
    burst = max(2.0 * n ** 3 / timed(f, 1) for _ in range(10)) / 1e12
    reps = max(10, int(4.0 / timed(f, 5)))
    sustained = 2.0 * n ** 3 / timed(f, reps) / 1e12
    return burst, sustained

out = {"gpu_name": torch.cuda.get_device_name(0), "torch": torch.__version__, "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
       "how": "torch.matmul 8192^3 (2*N^3 flop): best of 10 single launches (burst) and back to back for ~4 s (sustained); copy: b.copy_(a) over 1 Gi bf16 elements, read+write bytes"}
out["tf32_tflops"], out["tf32_tflops_sustained"] = matmul_rates(torch.float32, True)
out["fp32_simt_tflops"], out["fp32_simt_tflops_sustained"] = matmul_rates(torch.float32, False)
out["bf16_tflops_check"], out["bf16_tflops_sustained_check"] = matmul_rates(torch.bfloat16, True)
a = torch.empty(1 << 30, dtype=torch.bfloat16, device="cuda"); b = torch.empty_like(a)
f = lambda: b.copy_(a)
for _ in range(3): f()
torch.cuda.synchronize()
nbytes = 2 * a.numel() * 2
out["hbm_gbs_burst_check"] = max(nbytes / timed(f, 1) for _ in range(10)) / 1e9
out["hbm_gbs_sustained"] = nbytes / timed(f, max(10, int(4.0 / timed(f, 5)))) / 1e9
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "peaks_tf32.json"), "w") as fh: json.dump(out, fh, indent=1)
