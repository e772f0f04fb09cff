#!/bin/bash
# warm launch list (gpu__time_duration only, caches NOT flushed between kernels) of bench steps; $1 = env knobs
mkdir -p gpurun_out
env $1 timeout 900 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none --csv \
  --log-file gpurun_out/launches_warm.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-throughput > gpurun_out/launches.log 2>&1
tail -2 gpurun_out/launches.log | cut -c1-200
wc -l gpurun_out/launches_warm.csv
