#!/bin/bash
# ncu --set full of the first $2 launches matching $1 in a bench warm-up step; exports raw + sass pages
mkdir -p gpurun_out
name=ncu_one
env $3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$1" -c ${2:-1} -f -o gpurun_out/$name \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-throughput > gpurun_out/$name.log 2>&1
tail -1 gpurun_out/$name.log | cut -c1-120
ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
ncu -i gpurun_out/$name.ncu-rep --page source --csv > gpurun_out/${name}_sass.csv 2>/dev/null
ls -la gpurun_out/$name.ncu-rep
