"""Experiment driver for the tcgen05 projection kernels: times the metric shape under different
ZB200_TC_CLUSTER (B-operand multicast width) and ZB200_TC_DEBUG (ablation bits: 1 skip the
correction MMA, 2 idle splitter, 4 no K-chunking, 8 128-patch tiles) settings."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os, torch
sys.path.insert(0, %r)
from motif_learn_b200.features import ZPs
dx = torch.rand((262144, 64, 64), device="cuda")
out = []
for prec in ("tf32", "tf32x3"):
    z = ZPs(12, 64, precision=prec)
    for _ in range(3): z.transform(dx)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): z.transform(dx)
    e1.record(); torch.cuda.synchronize()
    out.append("%%s %%.4f ms" %% (prec, e0.elapsed_time(e1)/20))
print("bstages", os.environ.get("ZB200_TC_BSTAGES","-"), "stages", os.environ.get("ZB200_TC_STAGES","-"), "cluster", os.environ.get("ZB200_TC_CLUSTER","-"), "dbg", os.environ.get("ZB200_TC_DEBUG","0"), " | ".join(out), flush=True)
''' % ROOT
for cl, dbg, st, bs in ((2, 0, 8, 2), (2, 0, 8, 3), (2, 0, 8, 4), (2, 0, 4, 3), (2, 0, 3, 4), (1, 0, 8, 3)):
    env = dict(os.environ, ZB200_TC_DEBUG=str(dbg), ZB200_TC_CLUSTER=str(cl), ZB200_TC_STAGES=str(st), ZB200_TC_BSTAGES=str(bs))
    try:
        subprocess.run([sys.executable, "-c", code], env=env, timeout=100)
    except subprocess.TimeoutExpired:
        print("cluster", cl, "dbg", dbg, "TIMEOUT", flush=True)
