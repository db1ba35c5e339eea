#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total/mean duration, share.
usage: python profiles/summarize_launches.py gpurun_out/launches_c2.csv > profiles/r01_launches_c2_summary.md"""
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
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except Exception:
            continue
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "s": 1e6}.get(unit, 1e-3)
        name = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
        name = re.sub(r"^void\s+", "", name)
        rows.append((name, v * scale))
    agg = defaultdict(lambda: [0, 0.0])
    for n, us in rows:
        agg[n][0] += 1
        agg[n][1] += us
    tot = sum(v[1] for v in agg.values())
    print(f"# launch list summary: {path}\n")
    print(f"total kernels {len(rows)}, total device time {tot / 1e3:.2f} ms (serialised, cold-cache: compare SHARES)\n")
    print("| kernel | launches | total ms | mean us | share |")
    print("|---|---:|---:|---:|---:|")
    for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{n}` | {c} | {us / 1e3:.2f} | {us / c:.1f} | {100 * us / tot:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1])
