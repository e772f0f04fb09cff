#!/bin/bash
# A/B of env knobs on the bench: tools/gpu_ab.sh "NAME=VAL ..." "NAME=VAL ..." ...
mkdir -p gpurun_out
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-throughput > gpurun_out/ab_$i.json 2> gpurun_out/ab_$i.err
  python - "$cfg" gpurun_out/ab_$i.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print("[%s] ms/step %.3f pairs/s %.2f" % (sys.argv[1], d["ms_per_step"], d["value"]))
    print("   ", json.dumps(d["per_op_ms"]))
except Exception as e:
    print("[%s] failed: %s" % (sys.argv[1], e))
PY
done
