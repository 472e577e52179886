"""ncu target: the UNet's level-0 small-K Linear (bias + residual) on the current build."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pfd_b200 import native as nv

dev = "cuda"
x = torch.randn(32768, 320, device=dev).half()
w = (torch.randn(320, 320, device=dev) * 320 ** -0.5).half()
b = torch.randn(320, device=dev).half()
r = torch.randn(32768, 320, device=dev).half()
o = torch.empty(32768, 320, device=dev, dtype=torch.float16)
for _ in range(3):
    nv.linear(x, w, b, residual=r, out=o)
torch.cuda.synchronize()
print("done")
