"""Small driver for `ncu --set full`: launches the hot kernels on representative UNet shapes.
  ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_kernel|flash_attn|gn_' -c 12 \
      -o gpurun_out/prof python tools/ncu_targets.py
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pfd_b200 import native as nv, attention as att

dev = "cuda"
B = 8
# 1) small square linear (to_q/to_out at level 1): M=32768, N=K=320
x = torch.randn(32768, 320, device=dev).half()
w = (torch.randn(320, 320, device=dev) * 320 ** -0.5).half()
b = torch.randn(320, device=dev).half()
o = torch.empty(32768, 320, device=dev, dtype=torch.float16)
for _ in range(2):
    nv.linear(x, w, b, out=o)
# 2) ResBlock conv at level 1: 320->320 @64x64, batch 8
xc = torch.randn(B, 64, 64, 320, device=dev).half()
wc = (torch.randn(320, 9 * 320, device=dev) * (9 * 320) ** -0.5).half()
oc = torch.empty(B, 64, 64, 320, device=dev, dtype=torch.float16)
for _ in range(2):
    nv.conv3x3(xc, wc, b, out=oc)
# 3) big conv 1920->640 @32
xd = torch.randn(B, 32, 32, 1920, device=dev).half()
wd = (torch.randn(640, 9 * 1920, device=dev) * (9 * 1920) ** -0.5).half()
od = torch.empty(B, 32, 32, 640, device=dev, dtype=torch.float16)
for _ in range(2):
    nv.conv3x3(xd, wd, None, out=od)
# 4) flash attention level 1
q = torch.randn(B * 8, 4096, 40, device=dev).half()
k = torch.randn(B * 8, 4096, 40, device=dev).half()
vt = torch.randn(B * 8, 40, 4096, device=dev).half()
oa = torch.empty(B, 4096, 320, device=dev, dtype=torch.float16)
for _ in range(2):
    nv.flash_attn(q, k, vt, B=B, heads=8, Nq=4096, Nk=4096, scale=40 ** -0.5, out=oa)
# 5) GroupNorm level 1
g = torch.ones(320, device=dev).half()
for _ in range(2):
    nv.groupnorm(xc, g, g, 1e-5, silu=True)
torch.cuda.synchronize()
print("done")
