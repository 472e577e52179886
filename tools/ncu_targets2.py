"""ncu driver: small-K linear, GEGLU projection, 3x3 convs (32x32 1920->640, 64x64 320->320), split-K 8x8 conv,
flash attention (level-0 self-attention) on the current build; two launches each (the second is L2-warm)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pfd_b200 import native as nv

dev = "cuda"
x = torch.randn(32768, 320, device=dev).half()
w = (torch.randn(320, 320, device=dev) * 320 ** -0.5).half()
b = torch.randn(320, device=dev).half()
o = torch.empty(32768, 320, device=dev, dtype=torch.float16)
for _ in range(2):
    nv.linear(x, w, b, out=o)
wg = (torch.randn(2560, 320, device=dev) * 320 ** -0.5).half()
wgp, _, bn = nv.pack_geglu(wg, None)
og = torch.empty(32768, 1280, device=dev, dtype=torch.float16)
for _ in range(2):
    nv.linear(x, wgp, None, act=nv.ACT_GEGLU, out=og, bn_force=bn)
xa = torch.randn(8, 32, 32, 1920, device=dev).half()
wa = (torch.randn(640, 9 * 1920, device=dev) * (9 * 1920) ** -0.5).half()
oa2 = torch.empty(8, 32, 32, 640, device=dev, dtype=torch.float16)
for _ in range(2):
    nv.conv3x3(xa, wa, b.new_zeros(640), out=oa2)
xb = torch.randn(8, 64, 64, 320, device=dev).half()
wb = (torch.randn(320, 9 * 320, device=dev) * (9 * 320) ** -0.5).half()
ob2 = torch.empty(8, 64, 64, 320, device=dev, dtype=torch.float16)
for _ in range(2):
    nv.conv3x3(xb, wb, b, out=ob2)
xc = torch.randn(8, 8, 8, 1280, device=dev).half()
wc = (torch.randn(1280, 9 * 1280, device=dev) * (9 * 1280) ** -0.5).half()
oc = torch.empty(8, 8, 8, 1280, device=dev, dtype=torch.float16)
for _ in range(2):
    nv.conv3x3(xc, wc, None, out=oc)
q = torch.randn(64, 4096, 40, device=dev).half()
k = torch.randn(64, 4096, 40, device=dev).half()
vt = torch.randn(64, 40, 4096, device=dev).half()
oa = torch.empty(8, 4096, 320, device=dev, dtype=torch.float16)
for _ in range(2):
    nv.flash_attn(q, k, vt, B=8, heads=8, Nq=4096, Nk=4096, scale=40 ** -0.5, out=oa)
torch.cuda.synchronize()
print("done")
