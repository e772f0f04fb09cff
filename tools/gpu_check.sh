#!/bin/bash
# One GPU-box pass: parity tests, bench line, optional ncu captures.  Usage: tools/gpu_check.sh [ncu]
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 400 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
    print("ms/step", d["ms_per_step"], "pairs/s", d["value"], "e2e", d["e2e"]["value"])
    print(json.dumps(d["per_op_ms"]))
except Exception as e:
    print("bench parse failed", e)
PY
if [ "$1" = "ncu" ]; then
  timeout 400 ncu --set full --clock-control none --import-source on \
    -k regex:'structure_embedding_tc_kernel|rpe_scores_softmax_v2_kernel|sinkhorn_kernel|hash_order_replay_kernel' -c 5 \
    -f -o gpurun_out/ncu_full_misc python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-throughput > gpurun_out/ncu_misc.log 2>&1
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:'gemm_tf32x3_kernel' -s 8 -c 4 \
    -f -o gpurun_out/ncu_full_gemm python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-throughput > gpurun_out/ncu_gemm.log 2>&1
  ls -la gpurun_out/*.ncu-rep
fi
