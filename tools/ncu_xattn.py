"""ncu target: the level-0 CROSS-attention launch of BASELINE configs[1] (B*heads = 64, 4096 queries, 148 context
keys, d = 40) - persistent short-key kernel, then the generic flash kernel on the same operands; two launches each
(the second is L2-warm)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pfd_b200 import native as nv

dev = "cuda"
torch.manual_seed(0)
q = torch.randn(64, 4096, 40, device=dev).half()
k = torch.zeros(64, 152, 40, device=dev, dtype=torch.float16)
k[:, :148] = torch.randn(64, 148, 40, device=dev).half()
vt = torch.zeros(64, 40, 152, device=dev, dtype=torch.float16)
vt[:, :, :148] = torch.randn(64, 40, 148, device=dev).half()
oa = torch.empty(8, 4096, 320, device=dev, dtype=torch.float16)
for _ in range(2):
    nv.flash_attn(q, k, vt, B=8, heads=8, Nq=4096, Nk=148, scale=40 ** -0.5, out=oa)
nv.set_env_option("xattn_short", 0)
for _ in range(2):
    nv.flash_attn(q, k, vt, B=8, heads=8, Nq=4096, Nk=148, scale=40 ** -0.5, out=oa)
torch.cuda.synchronize()
print("done")
