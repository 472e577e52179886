"""Throughput of the UNet's small-K Linear shapes (graph-timed), for A/B runs of GEMM variants via env switches."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pfd_b200 import native as nv
from tools.gemm_perf import timeit


def main():
    dev = "cuda"
    for (M, N, K, act) in [(32768, 320, 320, 0), (32768, 640, 320, 0), (32768, 2560, 320, 4), (32768, 320, 1280, 0),
                           (8192, 640, 640, 0), (8192, 1280, 640, 0), (8192, 5120, 640, 4), (8192, 640, 2560, 0),
                           (2048, 1280, 1280, 0), (2048, 10240, 1280, 4)]:
        x = torch.randn(M, K, device=dev).half()
        w = (torch.randn(N, K, device=dev) * K ** -0.5).half()
        bn = 0
        if act == 4:
            w, _, bn = nv.pack_geglu(w, None)
        out = torch.empty(M, N // 2 if act == 4 else N, device=dev, dtype=torch.float16)
        ms = timeit(lambda: nv.linear(x, w, None, act=act, out=out, bn_force=bn))
        print(json.dumps(dict(op="linear" + ("+geglu" if act == 4 else ""), M=M, N=N, K=K, us=ms * 1e3,
                              tflops=2.0 * M * N * K / ms / 1e9)))


if __name__ == "__main__":
    main()
