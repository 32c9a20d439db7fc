#!/usr/bin/env python
"""ncu --csv metric log (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum per launch) -> the small JSON
bench.py scales to report roofline.traffic:   traffic_from_ncu.py launches.csv PARTICLES NTAU > profiles/r2e_traffic.json"""
import csv
import json
import sys

path, particles, ntau = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
lines = open(path).read().splitlines()
i = [k for k, l in enumerate(lines) if l.startswith('"ID"')][0]
acc = {}
for r in csv.DictReader(lines[i:]):
    name = "k_onepass_a" if "k_onepass_a" in r["Kernel Name"] else "k_onepass_b" if "k_onepass_b" in r["Kernel Name"] else None
    if not name:
        continue
    val = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"].lower()
    scale = {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}.get(unit, 1.0)
    acc.setdefault(name, {}).setdefault(r["Metric Name"], []).append(val * scale)
out = {"particles": particles, "ntau": ntau, "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum, mean over the captured launches ({path})"}
for k, m in acc.items():
    mean = lambda key: sum(m[key]) / len(m[key])
    out[k] = {"dram_bytes_read": mean("dram__bytes_read.sum"), "dram_bytes_write": mean("dram__bytes_write.sum"), "seconds": mean("gpu__time_duration.sum"),
              "launches": len(m["gpu__time_duration.sum"])}
json.dump(out, sys.stdout, indent=1)
