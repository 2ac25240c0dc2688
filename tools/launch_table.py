#!/usr/bin/env python
"""Aggregates an `ncu --csv` launch list (gpu__time_duration.sum [+ dram__bytes_read.sum, dram__bytes_write.sum]) per kernel:
launches, total time, share, DRAM bytes and achieved DRAM GB/s. Usage: python tools/launch_table.py launches.csv [--md]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
per = collections.OrderedDict()
for row in r:
    d = dict(zip(hdr, row))
    key = d["ID"]
    k = per.setdefault(key, {"name": re.sub(r"\(.*", "", d["Kernel Name"]).replace("magpo::<unnamed>::", "").replace("magpo::", "").replace("void ", ""),
                             "grid": d["Grid Size"]})
    try:
        v = float(d["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    unit = d["Metric Unit"]
    scale = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
    k[d["Metric Name"]] = v * scale
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for k in per.values():
    a = agg[k["name"]]
    a[0] += 1
    a[1] += k.get("gpu__time_duration.sum", 0.0)
    a[2] += k.get("dram__bytes_read.sum", 0.0) + k.get("dram__bytes_write.sum", 0.0)
tot = sum(a[1] for a in agg.values())
md = "--md" in sys.argv
if "--json" in sys.argv:
    import json
    out = sys.argv[sys.argv.index("--json") + 1]
    json.dump({"source": path, "total_ms": tot / 1e6,
               "kernels": {n.split("<")[0] if n.startswith("gemm") else n: {"launches": a[0], "ms": a[1] / 1e6, "dram_bytes": a[2]}
                           for n, a in agg.items()}}, open(out, "w"), indent=1)
print(f"total {tot / 1e6:.3f} ms over {sum(a[0] for a in agg.values())} launches")
if md:
    print("| kernel | launches | ms | share | DRAM GB | GB/s |\n|---|---|---|---|---|---|")
for name, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    gbs = a[2] / a[1] if a[1] else 0.0
    if md:
        print(f"| `{name}` | {a[0]} | {a[1] / 1e6:.3f} | {100 * a[1] / tot:.1f}% | {a[2] / 1e9:.3f} | {gbs:.0f} |")
    else:
        print(f"{name:48s} {a[0]:5d} {a[1] / 1e6:9.3f} ms {100 * a[1] / tot:5.1f}%  {a[2] / 1e9:8.3f} GB {gbs:7.0f} GB/s")
