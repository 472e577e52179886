"""Cross-attention launch (Nk = 148 context tokens) of every UNet level at BASELINE configs[1] (batch 4 -> 8 CFG
samples): the persistent short-key kernel (xattn_short_kernel, d <= 48) against the generic flash kernel, and the
polynomial-exp2 share (flash_poly_mod) on both the cross- and the level-0 self-attention launch.  Graph-timed
(no host gaps), same operands reused (L2-warm), interleaved rounds.  MUFU floor = exps / (148 SMs x 16 / clk x f).

    python tools/xattn_perf.py [--rounds 5] [--reps 20]
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
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    dev = "cuda"
    torch.manual_seed(0)
    res = {}
    cases = [("cross L0 512^2", 8, 8, 4096, 148, 40), ("cross L0 768^2", 8, 8, 9216, 148, 40),
             ("cross L1 512^2", 8, 8, 1024, 148, 80), ("self  L0 512^2", 8, 8, 4096, 4096, 40)]
    for name, B, H, Nq, Nk, d in cases:
        Nkp = (Nk + 7) // 8 * 8
        q = torch.randn(B * H, Nq, d, device=dev).half()
        k = torch.zeros(B * H, Nkp, d, device=dev, dtype=torch.float16)
        k[:, :Nk] = torch.randn(B * H, Nk, d, device=dev).half()
        vt = torch.zeros(B * H, d, Nkp, device=dev, dtype=torch.float16)
        vt[:, :, :Nk] = torch.randn(B * H, d, Nk, device=dev).half()
        out = torch.empty(B, Nq, H * d, device=dev, dtype=torch.float16)
        run = lambda: nv.flash_attn(q, k, vt, B=B, heads=H, Nq=Nq, Nk=Nk, scale=d ** -0.5, out=out)
        variants = {"generic": {"xattn_short": 0}, "short": {}, "short poly/4": {"flash_poly_mod": 4}, "short f16x2": {"flash_poly_mod": 1},
                    "generic poly/4": {"xattn_short": 0, "flash_poly_mod": 4}}
        if not (Nk <= 160 and d <= 48):
            variants = {"generic": {}, "generic f16x2": {"flash_poly_mod": 1}, "generic poly/4": {"flash_poly_mod": 4},
                        "generic poly/3": {"flash_poly_mod": 3}}
        graphs = {}
        for vn, opts in variants.items():
            nv.set_env_option(None, None)
            for kk, vv in opts.items():
                nv.set_env_option(kk, vv)
            graphs[vn] = graph_of(run, a.reps)
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
        exps = B * H * Nq * Nk
        flops = 4.0 * B * H * Nq * Nk * d
        floor_us = exps / (148 * 16 * 1.9e9) * 1e6
        print(f"{name}: B*h={B * H} Nq={Nq} Nk={Nk} d={d}  ({flops / 1e9:.1f} GFLOP, {exps / 1e6:.1f} M exp, MUFU floor {floor_us:.1f} us @1.9 GHz)")
        res[name] = {}
        for vn, ts in times.items():
            ts = sorted(ts)
            med = ts[len(ts) // 2]
            res[name][vn] = med
            print(f"    {vn:16s} {med:8.2f} us   {flops / med / 1e6:7.1f} TF/s   MUFU-floor frac {floor_us / med:.2f}")
    print("XATTN_RESULT " + json.dumps(res))


if __name__ == "__main__":
    main()
