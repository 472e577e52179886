"""A/B of a GEMM option on the UNet's 3x3 convs and large-K Linears at BASELINE configs[1] (8 CFG samples): by default
the CTA-pair kernel (cta_group::2, gemm_pair = 1 vs 0); --option gemm_streamk --on 1 --off 0 times the stream-K tail.
Graph-timed, interleaved rounds, L2-warm operands.

    python tools/gemm_pair_ab.py [--option gemm_pair --on 1 --off 0] [--rounds 5] [--reps 10]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pfd_b200 import native as nv


def graph_of(fn, reps):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    return g


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=5)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--option", default="gemm_pair")
    ap.add_argument("--on", type=int, default=1)
    ap.add_argument("--off", type=int, default=0)
    a = ap.parse_args()
    dev = "cuda"
    torch.manual_seed(0)
    cases = []
    for (H, C, N) in [(64, 320, 320), (64, 640, 320), (64, 960, 320), (32, 640, 640), (32, 1280, 640), (32, 1920, 640),
                      (16, 1280, 1280), (16, 2560, 1280), (32, 320, 640), (16, 640, 1280)]:
        x = torch.randn(8, H, H, C, device=dev).half()
        w = (torch.randn(N, 9 * C, device=dev) * (9 * C) ** -0.5).half()
        b = torch.randn(N, device=dev).half()
        ra = torch.randn(8, N, device=dev).half()
        o = torch.empty(8, H, H, N, device=dev, dtype=torch.float16)
        cases.append((f"conv3x3 {H}x{H} {C}->{N}", 2.0 * 8 * H * H * N * 9 * C,
                      (lambda x=x, w=w, b=b, ra=ra, o=o: nv.conv3x3(x, w, b, rowadd=ra, out=o))))
    for (M, N, K) in [(32768, 320, 1280), (8192, 640, 2560), (2048, 1280, 5120), (8192, 8192, 8192)]:
        x = torch.randn(M, K, device=dev).half()
        w = (torch.randn(N, K, device=dev) * K ** -0.5).half()
        b = torch.randn(N, device=dev).half()
        r = torch.randn(M, N, device=dev).half()
        o = torch.empty(M, N, device=dev, dtype=torch.float16)
        cases.append((f"linear {M}x{N}x{K} +res", 2.0 * M * N * K,
                      (lambda x=x, w=w, b=b, r=r, o=o: nv.linear(x, w, b, residual=r, out=o))))
    res_all = {}
    for name, flops, fn in cases:
        graphs = {}
        for vn, val in (("pair", a.on), ("single", a.off)):
            nv.set_env_option(None, None)
            nv.set_env_option(a.option, val)
            graphs[vn] = graph_of(fn, a.reps)
        nv.set_env_option(None, None)
        times = {vn: [] for vn in graphs}
        for _ in range(a.rounds):
            for vn, g in graphs.items():
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize()
                times[vn].append(e0.elapsed_time(e1) / a.reps * 1000.0)
        med = {vn: sorted(ts)[len(ts) // 2] for vn, ts in times.items()}
        res_all[name] = med
        print(f"{name:34s} on {med['pair']:8.2f} us ({flops / med['pair'] / 1e6:7.1f} TF/s)   off {med['single']:8.2f} us "
              f"({flops / med['single'] / 1e6:7.1f} TF/s)   speed-up {med['single'] / med['pair']:.2f}", flush=True)
    print("PAIR_AB_RESULT " + json.dumps(res_all))


if __name__ == "__main__":
    main()
