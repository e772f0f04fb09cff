#!/usr/bin/env python
"""Split an ncu gpu__time_duration launch list into bench steps (a step starts with the pyramid's first bbox_kernel after
a gap) and print the per-kernel table of step `which` (default: the 4th = first timed one).
    python tools/launch_steps.py gpurun_out/launches_warm.csv [which] [out.txt]"""
import csv, re, sys
from collections import OrderedDict
path = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 3
lines = [l for l in open(path) if l.startswith('"')]
rd = csv.reader(lines); hdr = next(rd); col = {h: i for i, h in enumerate(hdr)}
L = OrderedDict()
for r in rd:
    if len(r) != len(hdr) or r[col["Metric Name"]] != "gpu__time_duration.sum": continue
    u = r[col["Metric Unit"]]
    sc = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3}.get(u, 1.0)
    L[int(r[col["ID"]])] = (r[col["Kernel Name"]], r[col["Grid Size"]], float(r[col["Metric Value"]].replace(",", "")) * sc)
ids = sorted(L)
def short(n): return re.sub(r"\(.*", "", n).replace("void ", "").replace("gr::", "")[:58]
# step boundaries: the grid-subsample chain starts with bbox_kernel; a step has 9 bbox launches -> take every first of a run
starts = []
prev_bbox = -100
for k, i in enumerate(ids):
    if "bbox_kernel" in L[i][0]:
        if k - prev_bbox > 200: starts.append(k)
        prev_bbox = k
starts.append(len(ids))
print("steps found:", len(starts) - 1, [starts[j + 1] - starts[j] for j in range(len(starts) - 1)])
a, b = starts[which], starts[which + 1]
sel = ids[a:b]
agg = OrderedDict()
for i in sel:
    n, g, us = L[i]
    e = agg.setdefault(short(n), [0, 0.0]); e[0] += 1; e[1] += us
tot = sum(e[1] for e in agg.values())
out = [f"bench step {which} under ncu --cache-control none ({path}): {len(sel)} launches, {tot:.1f} us summed kernel time (warm caches, serialised)",
       "       us    n   avg us  share  kernel"]
for k, e in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{e[1]:9.1f} {e[0]:4d} {e[1]/e[0]:8.1f} {100*e[1]/tot:5.1f}%  {k}")
txt = "\n".join(out) + "\n"
if len(sys.argv) > 3: open(sys.argv[3], "w").write(txt)
print(txt)
