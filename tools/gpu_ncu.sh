#!/bin/bash
# ncu --set full capture of kernels matching $1 (regex) during one bench step; optional env in $2; -c count in $3
mkdir -p gpurun_out
name=${4:-ncu_$(echo "$1" | tr -c 'a-zA-Z0-9' '_' | cut -c1-40)}
env $2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$1" -c ${3:-6} \
  -f -o gpurun_out/$name python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-throughput > gpurun_out/$name.log 2>&1
tail -3 gpurun_out/$name.log
ls -la gpurun_out/$name.ncu-rep
