#!/bin/bash
# launch list (time + DRAM bytes per launch) of bench steps under ncu; $1 = env knobs
mkdir -p gpurun_out
env $1 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-throughput > gpurun_out/launches.log 2>&1
tail -2 gpurun_out/launches.log | cut -c1-300
wc -l gpurun_out/launches.csv
