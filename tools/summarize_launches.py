"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (count, total, share)."""
import csv
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((r["Kernel Name"].split("(")[0], float(r["Metric Value"].replace(",", ""))))
    agg = defaultdict(lambda: [0, 0.0])
    for k, v in rows:
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# {path}: {len(rows)} launches, {tot / 1e6:.3f} ms total (cold-cache, serialised: compare shares)")
    print(f"{'kernel':60s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:60s} {v[0]:8d} {v[1] / 1e3:12.1f} {v[1] / v[0] / 1e3:10.2f} {100 * v[1] / tot:6.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
