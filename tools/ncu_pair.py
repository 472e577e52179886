"""ncu target: the largest UNet conv (32x32, 1920->640, 8 samples: 256 tiles = 1.73 waves) on the single-CTA kernel, the
CTA-pair kernel (gemm_pair = 1) and with the stream-K tail (gemm_streamk = 1); two launches each (the second L2-warm)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pfd_b200 import native as nv

dev = "cuda"
torch.manual_seed(0)
xa = torch.randn(8, 32, 32, 1920, device=dev).half()
wa = (torch.randn(640, 9 * 1920, device=dev) * (9 * 1920) ** -0.5).half()
b = torch.zeros(640, device=dev, dtype=torch.float16)
oa = torch.empty(8, 32, 32, 640, device=dev, dtype=torch.float16)
for opts in ({}, {"gemm_pair": 1}, {"gemm_streamk": 1}):
    nv.set_env_option(None, None)
    for k, v in opts.items():
        nv.set_env_option(k, v)
    for _ in range(2):
        nv.conv3x3(xa, wa, b, out=oa)
torch.cuda.synchronize()
print("done")
