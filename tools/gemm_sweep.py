"""Sweep the forced N-tile width for the small-K Linear shapes of the UNet (graph-timed device time)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pfd_b200 import native as nv
from tools.gemm_perf import timeit

dev = "cuda"
for (M, N, K) in [(32768, 320, 320), (8192, 640, 640), (2048, 1280, 1280), (32768, 960, 320), (8192, 1920, 640),
                  (2048, 3840, 1280), (32768, 320, 1280), (32768, 1280, 320)]:
    x = torch.randn(M, K, device=dev).half()
    w = (torch.randn(N, K, device=dev) * K ** -0.5).half()
    o = torch.empty(M, N, device=dev, dtype=torch.float16)
    for bn in (0, 64, 128, 160, 192, 256):
        try:
            ms = timeit(lambda: nv.linear(x, w, None, out=o, bn_force=bn), n=20)
            print(json.dumps(dict(M=M, N=N, K=K, bn=bn, us=ms * 1e3, tflops=2.0 * M * N * K / ms / 1e9)))
        except Exception as e:
            print(json.dumps(dict(M=M, N=N, K=K, bn=bn, err=str(e)[:80])))
