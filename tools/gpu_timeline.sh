#!/bin/bash
# host/device timeline of one step (tools/step_timeline.py) for each env setting given as an argument ("" = defaults)
mkdir -p gpurun_out
i=0
for cfg in "${@:-}"; do
  i=$((i+1))
  env $cfg timeout 300 python tools/step_timeline.py > gpurun_out/timeline_$i.txt 2>&1
  echo "[$cfg]"; head -3 gpurun_out/timeline_$i.txt
done
