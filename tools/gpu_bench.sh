#!/bin/bash
# bench line(s) with env knobs: tools/gpu_bench.sh "ENV=.. ENV=.." [extra bench args]
mkdir -p gpurun_out
cfg="$1"; shift
env $cfg timeout 900 python bench.py --steps 20 --warmup 3 "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 600 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
    print("ms/step", d["ms_per_step"], "pairs/s", d["value"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
    print(json.dumps(d["per_op_ms"]))
    for k in ("gpu_torch_baseline", "cpu_baseline", "config3_128_pairs_1gpu", "config4_sharded_pairs", "config5_200k_pair"):
        print(k, json.dumps(d.get(k)))
    print("throughput", json.dumps(d.get("throughput_mode")))
except Exception as e:
    print("bench parse failed", e)
PY
