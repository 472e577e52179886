"""Join an ncu launch list (gpu__time_duration.sum, csv) of tools/unet_eval_profile.py with its PFD_GEMM_TRACE
lines: per-shape device time of the GEMM launches of ONE UNet evaluation (the last one in the run).
Usage: python tools/gemm_breakdown.py launches.csv trace.log"""
import csv
import re
import sys
from collections import OrderedDict


def main(csv_path, trace_path):
    with open(csv_path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = []
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = val * {"ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(unit, 1)
        rows.append((r["Kernel Name"], ns))
    gemm = [(n, ns) for n, ns in rows if "gemm_tc_kernel" in n]
    tr, seg = [], []
    for l in open(trace_path):
        if l.startswith("EVALMARK"):
            if seg:
                tr.append(seg)
            seg = []
        elif l.startswith("GEMMTRACE"):
            dd = dict(kv.split("=") for kv in l.split()[1:])
            if dd.get("tmae") == "1":
                dd["lean"] = "T"                      # lean epilogue with TMA stores / TMA-loaded residual
            seg.append(dd)
    last = tr[-1]
    times = gemm[-len(last):]          # the last evaluation's launches are the last GEMM rows of the list
    agg = OrderedDict()
    for d, (name, ns) in zip(last, times):
        key = tuple(d[k] for k in ("M", "N", "K", "taps", "stride", "act", "bias", "res", "rowadd", "BN", "lean", "splits", "grid", "batched", "plain"))
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(a[1] for a in agg.values())
    other = sum(ns for n, ns in rows[-(len(rows) // 2):] if "gemm_tc_kernel" not in n)
    print(f"GEMM launches in the evaluation: {len(last)}   summed device time {tot / 1e6:.3f} ms")
    print(f"{'M':>6} {'N':>6} {'K':>6} tap s act b r ra  BN lean spl grid bat pln {'cnt':>4} {'us each':>8} {'ms tot':>7} {'TF/s':>7} {'share':>6}")
    for key, (cnt, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        M, N, K = int(key[0]), int(key[1]), int(key[2])
        fl = 2.0 * M * N * K * int(key[3] if False else 1)
        print(f"{M:6d} {N:6d} {K:6d} {key[3]:>3} {key[4]} {key[5]:>3} {key[6]} {key[7]} {key[8]:>2} {key[9]:>4} {key[10]:>4} {key[11]:>3} {key[12]:>4} {key[13]:>3} {key[14]:>3} "
              f"{cnt:4d} {ns / cnt / 1e3:8.2f} {ns / 1e6:7.3f} {fl * cnt / ns / 1e3:7.1f} {100 * ns / tot:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
