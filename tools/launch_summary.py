#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list:
per-kernel count / time / share / DRAM bytes for the LAST bench step in the file, written as text + JSON.

    python tools/launch_summary.py gpurun_out/launches.csv profiles/r02x  [launches_per_step]
"""
import csv, json, re, sys
from collections import OrderedDict

path, out = sys.argv[1], sys.argv[2]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
col = {h: i for i, h in enumerate(hdr)}
launch = OrderedDict()
for r in rd:
    if len(r) != len(hdr): continue
    lid = int(r[col["ID"]])
    e = launch.setdefault(lid, {"name": r[col["Kernel Name"]], "us": 0.0, "rd": 0.0, "wr": 0.0})
    m, v, u = r[col["Metric Name"]], float(r[col["Metric Value"]].replace(",", "")), r[col["Metric Unit"]]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    if m == "gpu__time_duration.sum": e["us"] = v * scale
    elif m == "dram__bytes_read.sum": e["rd"] = v * scale
    elif m == "dram__bytes_write.sum": e["wr"] = v * scale
ids = list(launch)
# one step = from the last 'bbox_kernel'-starting pyramid back to the end: find step boundaries by the first kernel of a step
first = launch[ids[0]]["name"]
starts = [i for i in ids if launch[i]["name"] == first]
n_per = int(sys.argv[3]) if len(sys.argv) > 3 else None
if n_per:
    sel = ids[-n_per:]
else:
    sel = ids
def short(n):
    n = re.sub(r"\(.*", "", n)
    return n.replace("void ", "").replace("gr::", "")[:70]
agg = OrderedDict()
for i in sel:
    e = launch[i]
    a = agg.setdefault(short(e["name"]), [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += e["us"]; a[2] += e["rd"]; a[3] += e["wr"]
tot = sum(a[1] for a in agg.values())
with open(out + "_by_kernel.txt", "w") as f:
    f.write(f"{len(sel)} launches, {tot:.1f} us summed kernel time (ncu: cold-cache, serialised; compare SHARES)\n")
    f.write("       us    n   avg us  share  dram rd MB  dram wr MB  kernel\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{a[1]:9.1f} {a[0]:4d} {a[1]/a[0]:8.1f} {100*a[1]/tot:5.1f}% {a[2]/1e6:10.1f} {a[3]/1e6:10.1f}  {k}\n")
json.dump({"source": path, "launches": len(sel), "sum_us": tot,
           "kernels": {k: {"n": a[0], "us": a[1], "dram_read_bytes": a[2], "dram_write_bytes": a[3]} for k, a in agg.items()}},
          open(out + "_traffic.json", "w"), indent=1)
print(open(out + "_by_kernel.txt").read())
