import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from motif_learn_b200.features import ZPs
img = torch.rand((100, 300), device="cuda")
z = ZPs(12, 48, precision=os.environ.get("PREC", "tf32"))
out = z.symmetry_map(img, [2, 3])
torch.cuda.synchronize()
print("OK", float(out.nan_to_num().abs().max()))
