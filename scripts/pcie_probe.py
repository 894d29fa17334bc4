"""Raw PCIe rates of the box (pinned memory): H2D of one 2048^2 frame, D2H of one frame's features, both at once."""
import torch, time
f = torch.empty((2048, 2048), dtype=torch.float32, pin_memory=True)
d = torch.empty_like(f, device="cuda")
o = torch.empty((21000, 91), dtype=torch.float32, device="cuda")
h = torch.empty((21000, 91), dtype=torch.float32, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(fn, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
def h2d():
    with torch.cuda.stream(s1): d.copy_(f, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h.copy_(o, non_blocking=True)
def both():
    h2d(); d2h()
t = run(h2d); print(f"H2D 16.8 MB: {t:.3f} ms = {f.numel()*4/t/1e6:.1f} GB/s")
t = run(d2h); print(f"D2H 7.6 MB: {t:.3f} ms = {o.numel()*4/t/1e6:.1f} GB/s")
t = run(both); print(f"both: {t:.3f} ms per pair")
