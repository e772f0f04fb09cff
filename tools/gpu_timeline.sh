#!/bin/bash
mkdir -p gpurun_out
for m in table tc; do
  GAUSSREG_T1=$m timeout 300 python tools/step_timeline.py > gpurun_out/timeline_$m.txt 2>&1
  head -4 gpurun_out/timeline_$m.txt
done
