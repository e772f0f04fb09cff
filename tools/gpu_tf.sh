#!/bin/bash
mkdir -p gpurun_out
python tools/tf_bench.py
REPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none --csv --log-file gpurun_out/tf_launches.csv python tools/tf_bench.py > gpurun_out/tf_ncu.log 2>&1
tail -2 gpurun_out/tf_ncu.log
python - <<'PY'
import csv,re
lines=[l for l in open('gpurun_out/tf_launches.csv') if l.startswith('"')]
rd=csv.reader(lines); hdr=next(rd); col={h:i for i,h in enumerate(hdr)}
rows=[r for r in rd if len(r)==len(hdr)]
# last call = last ~110 launches; print them in order with grid and time
names=[(re.sub(r"\(.*","",r[col['Kernel Name']]).replace("void ","").replace("gr::","")[:44], r[col['Grid Size']], float(r[col['Metric Value']].replace(',',''))/ (1000.0 if r[col['Metric Unit']] in ('ns','nsecond') else 1.0)) for r in rows]
n=len(names)
# find the start of the last transformer call: last occurrence of a memcpy is not a kernel; take last 112 kernels
last=names[-112:]
tot=sum(x[2] for x in last)
print("last call: %d kernels, %.1f us summed (warm caches, serialised)"%(len(last),tot))
agg={}
for nm,g,us in last:
    a=agg.setdefault((nm,g),[0,0.0]); a[0]+=1; a[1]+=us
for (nm,g),a in sorted(agg.items(), key=lambda kv:-kv[1][1]): print("%8.1f us  n=%3d  avg %6.1f  %s %s"%(a[1],a[0],a[1]/a[0],nm,g))
PY
