#!/bin/bash
# multi-GPU bench line exactly as the driver launches it: tools/gpu_bench_n.sh N
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 400 gpurun_out/bench_n$N.err
python - $N <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/bench_n{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("n_gpus", d["n_gpus"], "ms/step", d["ms_per_step"], "pairs/s", d["value"], "e2e", d["e2e"]["value"], "scaling", d["scaling"])
print("config4", json.dumps(d.get("config4_sharded_pairs")))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-300
