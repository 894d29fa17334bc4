"""Pure-write HBM bandwidth (fill of a 4.3 GB buffer) against the copy peak: the denominator that applies to
the write-only gather kernel."""
import torch
n = 262144 * 4096
buf = torch.empty(n, dtype=torch.float32, device="cuda")
src = torch.empty(n, dtype=torch.float32, device="cuda")
def t(fn, k=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms_fill = t(lambda: buf.fill_(1.0))
ms_copy = t(lambda: buf.copy_(src))
ms_read = t(lambda: src.sum())
print(f"fill {n*4/ms_fill/1e6:.0f} GB/s   copy {2*n*4/ms_copy/1e6:.0f} GB/s (r+w)   read(sum) {n*4/ms_read/1e6:.0f} GB/s")
