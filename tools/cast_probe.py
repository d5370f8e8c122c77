import sys, os, torch
sys.path.insert(0, os.getcwd())
from laff_b200 import ops
x = torch.randn(131072, 2048, device="cuda")
out = torch.empty(131072, 2048, dtype=torch.bfloat16, device="cuda")
for _ in range(4):
    ops.cast_pad_16(x, torch.bfloat16, out=out)
w = torch.randn(4096, 3981, device="cuda")
for _ in range(4):
    ops.split3_16(w, 1, torch.bfloat16)
torch.cuda.synchronize()
