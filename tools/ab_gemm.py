"""Per-shape A/B of the lean GEMM epilogue variants (8 vs 16 epilogue warps; with / without producer-side GroupNorm
statistics) on the UNet's epilogue-bound shapes, interleaved inside one process (graph-timed, 20 launches per graph)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pfd_b200 import native as nv  # noqa: E402


def graph_of(fn, n=20):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    return g


def time_graph(g, n=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    dev = "cuda"
    shapes = [("linear", 32768, 320, 320, 0, True), ("linear", 32768, 320, 320, 0, False), ("geglu", 32768, 2560, 320, 4, False),
              ("linear", 32768, 320, 1280, 0, True), ("linear", 32768, 640, 320, 0, False),
              ("linear", 8192, 640, 640, 0, True), ("geglu", 8192, 5120, 640, 4, False), ("linear", 8192, 640, 2560, 0, True),
              ("linear", 2048, 1280, 1280, 0, True), ("geglu", 2048, 10240, 1280, 4, False), ("linear", 2048, 1280, 5120, 0, True),
              ("conv", 8, 64, 320, 320, True), ("conv", 8, 32, 640, 640, True), ("conv", 8, 64, 640, 320, False)]
    for sh in shapes:
        kind = sh[0]
        if kind == "conv":
            _, B, H, C, N, res = sh
            x = torch.randn(B, H, H, C, device=dev).half()
            w = (torch.randn(N, 9 * C, device=dev) * (9 * C) ** -0.5).half()
            b = torch.randn(N, device=dev).half()
            r = torch.randn(B, H, H, N, device=dev).half() if res else None
            out = torch.empty(B, H, H, N, device=dev, dtype=torch.float16)
            flops = 2.0 * B * H * H * N * 9 * C
            mk = lambda su: (lambda: (nv.gn_reset(), nv.conv3x3(x, w, b, residual=r, out=out, stats_unit=su)))
            label = f"conv3x3 {B}x{H}x{H}x{C}->{N} res={int(res)}"
        else:
            _, M, N, K, act, res = sh
            x = torch.randn(M, K, device=dev).half()
            w = (torch.randn(N, K, device=dev) * K ** -0.5).half()
            b = torch.randn(N, device=dev).half()
            bn = 0
            if act == 4:
                w, b, bn = nv.pack_geglu(w, b)
            n_out = N // 2 if act == 4 else N
            r = torch.randn(M, n_out, device=dev).half() if res else None
            out = torch.empty(M, n_out, device=dev, dtype=torch.float16)
            flops = 2.0 * M * N * K
            mk = lambda su: (lambda: (nv.gn_reset(), nv.linear(x, w, b, act=act, residual=r, out=out, bn_force=bn, stats_unit=su)))
            label = f"{kind} M={M} N={N} K={K} res={int(res)}"
        graphs = {}
        for name, ew, su in (("ew8", 8, 0), ("ew16", 16, 0), ("ew8+stats", 8, 10 if kind != "geglu" else 0)):
            if name == "ew8+stats" and su == 0:
                continue
            nv.set_option("gemm_epilogue_warps", ew)
            graphs[name] = graph_of(mk(su))
        nv.set_option("gemm_epilogue_warps", 0)
        graphs["memset only"] = graph_of(lambda: nv.gn_reset())
        ts = {k: [] for k in graphs}
        for _ in range(5):
            for k, g in graphs.items():
                ts[k].append(time_graph(g))
        base = sorted(ts.pop("memset only"))[2]
        row = {k: sorted(v)[2] - base for k, v in ts.items()}
        print(label + "  " + "  ".join(f"{k}: {v:7.2f} us ({flops / v / 1e6:6.0f} TF/s)" for k, v in row.items()), flush=True)
        print("ABG " + json.dumps({"shape": label, **row}))


if __name__ == "__main__":
    main()
