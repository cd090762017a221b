"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (this library's smb:: kernels and the rest).
usage: python tools/launch_list.py gpurun_out/launches.csv"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
# a capture filtered with -k regex:<this library's kernels> prints base names without the smb:: namespace: then every row is ours
filtered = not any("smb::" in r[ik] for r in rows[1:])
for r in rows[1:]:
    name = r[ik].split("<")[0].split("(")[0].replace("void ", "")
    if filtered:
        name = "smb::" + name
    if "smb::" not in name:
        name = "(not this library: torch fill / copy / random init)"
    us = float(r[iv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[iu], 1e-3)
    agg[name][0] += 1
    agg[name][1] += us
tot = sum(v[1] for k, v in agg.items() if "smb::" in k)
print("| kernel | launches | total us | avg us | share of smb:: time |")
print("|---|---|---|---|---|")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k} | {n} | {us:.1f} | {us / n:.2f} | {us / tot:.3f} |" if "smb::" in k else f"| {k} | {n} | {us:.1f} | {us / n:.2f} | - |")
