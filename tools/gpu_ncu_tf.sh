#!/bin/bash
# ncu --set full of transformer kernels matching $1 in tools/tf_bench.py (REPS=1); $2 = count, $3 = env
mkdir -p gpurun_out
name=ncu_tf_$(echo "$1" | tr -c 'a-zA-Z0-9' '_' | cut -c1-30)
env REPS=1 $3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${4:-0} -c ${2:-6} \
  -f -o gpurun_out/$name python tools/tf_bench.py > gpurun_out/$name.log 2>&1
tail -2 gpurun_out/$name.log
ls -la gpurun_out/$name.ncu-rep
