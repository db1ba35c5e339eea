#!/usr/bin/env python
"""ncu raw-page CSV (ncu -i X.ncu-rep --page raw --csv) -> per-kernel DRAM traffic per launch (median over the captured launches).
usage: python profiles/extract_traffic.py raw.csv out.json"""
import csv
import json
import statistics
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "nsecond": 1e-3,
         "msecond": 1e3}


def val(r, name):
    return float(r[col[name]].replace(",", "")) * SCALE.get(units[col[name]], 1.0)


out = {}
KEYS = {  # the GEMM is captured in the default fp8-corrected form (~0.83 ms) and as three fp16 passes (~1.16 ms): told apart by duration
        "gemm_tc_kernel": lambda n, r: "gemm_tc" in n and val(r, "gpu__time_duration.sum") < 1000.0,
        "gemm_tc_kernel_3xfp16": lambda n, r: "gemm_tc" in n and val(r, "gpu__time_duration.sum") >= 1000.0,
        "pips_corr": lambda n, r: "pips_corr" in n,
        # the two attention launches of tools/ncu_targets.py: the windowed one has 4000 items (grid 148), told apart by duration
        "attn_windowed": lambda n, r: "attn_ws" in n and val(r, "gpu__time_duration.sum") < 800.0,
        "attn_global": lambda n, r: "attn_ws" in n and val(r, "gpu__time_duration.sum") >= 800.0}
for key, pred in KEYS.items():
    rs = [r for r in data if pred(r[col["Kernel Name"]], r)]
    if not rs:
        continue
    rd = [val(r, "dram__bytes_read.sum") for r in rs]
    wr = [val(r, "dram__bytes_write.sum") for r in rs]
    us = [val(r, "gpu__time_duration.sum") for r in rs]
    entry = {"kernel": rs[0][col["Kernel Name"]][:80], "launches_captured": len(rs), "dram_read_bytes": statistics.median(rd),
             "dram_write_bytes": statistics.median(wr), "duration_us_under_ncu": statistics.median(us)}
    for extra in ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
                  "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
                  "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
                  "sm__cycles_active.avg", "smsp__cycles_active.avg", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
                  "smsp__cycles_elapsed.avg.per_second"):
        if extra in col:
            entry[extra] = statistics.median(float(r[col[extra]].replace(",", "")) for r in rs)
    out[key] = entry
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
