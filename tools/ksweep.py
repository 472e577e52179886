"""Fixed-overhead / epilogue-cost probe for the GEMM kernel: time vs K at fixed M, N (graph-timed)."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pfd_b200 import native as nv
from tools.gemm_perf import timeit

dev = "cuda"
for (M, N) in [(128, 320), (18944, 320), (32768, 320), (32768, 640), (8192, 640), (2048, 1280)]:
    for K in (64, 320, 640, 1280):
        for full in (0, 1):
            x = torch.randn(M, K, device=dev).half()
            w = (torch.randn(N, K, device=dev) * K ** -0.5).half()
            b = torch.randn(N, device=dev).half() if full else None
            r = torch.randn(M, N, device=dev).half() if full else None
            out = torch.empty(M, N, device=dev, dtype=torch.float16)
            ms = timeit(lambda: nv.linear(x, w, b, residual=r, out=out), n=20)
            print(json.dumps(dict(M=M, N=N, K=K, bias_res=full, us=round(ms * 1e3, 2),
                                  tflops=round(2.0 * M * N * K / ms / 1e9, 1))))
