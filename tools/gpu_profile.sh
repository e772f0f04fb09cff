#!/bin/bash
# ncu launch list of one bench step + full captures of selected kernels.  Usage: tools/gpu_profile.sh 'regex' [count]
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2200 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-throughput > gpurun_out/ncu_bench.log 2>&1
wc -l gpurun_out/launches.csv
if [ -n "$1" ]; then
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${3:-0} -c ${2:-6} \
    -f -o gpurun_out/ncu_full_sel python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-throughput > gpurun_out/ncu_sel.log 2>&1
  ls -la gpurun_out/ncu_full_sel.ncu-rep
fi
