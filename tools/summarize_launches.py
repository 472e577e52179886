"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = val * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "second": 1e9}.get(unit, 1)
        name = r["Kernel Name"]
        name = re.sub(r"\(.*$", "", name)
        rows.append((name, ns, r.get("Grid Size", "")))
    tot = sum(r[1] for r in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for n, ns, _ in rows:
        agg[n][0] += 1
        agg[n][1] += ns
    print(f"launches: {len(rows)}   total device time: {tot / 1e6:.3f} ms")
    print(f"{'kernel':70s} {'count':>7s} {'total ms':>10s} {'share':>7s} {'avg us':>9s}")
    for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n[:70]:70s} {c:7d} {ns / 1e6:10.3f} {100 * ns / tot:6.1f}% {ns / c / 1e3:9.2f}")


if __name__ == "__main__":
    main(sys.argv[1])
