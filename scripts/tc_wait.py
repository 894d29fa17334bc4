"""Prints, per CTA, the cycles each warp role of the fp32-grade projection kernel spends blocked
on its mbarriers (ZB200_TC_DEBUG=16 instrumentation) -- shows which hand-off limits the pipeline."""
import os, sys
os.environ.setdefault("ZB200_TC_DEBUG", "16")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from motif_learn_b200.features import ZPs
dx = torch.rand((262144, 64, 64), device="cuda")
z = ZPs(12, 64, precision="tf32x3")
for _ in range(3):
    z.transform(dx)
torch.cuda.synchronize()
