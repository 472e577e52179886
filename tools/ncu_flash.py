"""ncu target: level-0 self-attention (B*heads = 64, N = 4096, d = 40) on the current build."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pfd_b200 import native as nv

dev = "cuda"
torch.manual_seed(0)
q = torch.randn(64, 4096, 40, device=dev).half()
k = torch.randn(64, 4096, 40, device=dev).half()
vt = torch.randn(64, 40, 4096, device=dev).half()
oa = torch.empty(8, 4096, 320, device=dev, dtype=torch.float16)
for _ in range(2):
    nv.flash_attn(q, k, vt, B=8, heads=8, Nq=4096, Nk=4096, scale=40 ** -0.5, out=oa)
torch.cuda.synchronize()
print("done")
