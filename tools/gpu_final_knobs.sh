#!/bin/bash
# Last GPU pass of the round: validate the three default-off kernels (single-launch GroupNorm for small inputs,
# exp2-domain Sinkhorn, fused AttentionOutput) -- parity tests with all of them on, then each alone if that fails,
# then the bench with each knob.
mkdir -p gpurun_out
ALL="GAUSSREG_GN_SMALL=1 GAUSSREG_SINKHORN_EXP2=1 GAUSSREG_TF_MLP=1"
env $ALL timeout 120 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_knobs_all.log 2>&1
echo "ALL: $(grep -E 'passed|failed|error' gpurun_out/pytest_knobs_all.log | tail -1)"
if ! grep -q " passed" gpurun_out/pytest_knobs_all.log || grep -q "failed" gpurun_out/pytest_knobs_all.log; then
  grep -E "^FAILED|Error" gpurun_out/pytest_knobs_all.log | head -5
  for k in GAUSSREG_GN_SMALL GAUSSREG_SINKHORN_EXP2 GAUSSREG_TF_MLP; do
    env $k=1 timeout 90 python -m pytest tests/test_network_gpu.py -m gpu -x -q > gpurun_out/pytest_knob_$k.log 2>&1
    echo "$k: $(grep -E 'passed|failed|error' gpurun_out/pytest_knob_$k.log | tail -1)"
  done
fi
i=0
for cfg in "GAUSSREG_NOOP=1" "$ALL" "GAUSSREG_GN_SMALL=1" "GAUSSREG_SINKHORN_EXP2=1" "GAUSSREG_TF_MLP=1"; do
  i=$((i+1))
  env $cfg timeout 60 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-throughput > gpurun_out/knob_$i.json 2> gpurun_out/knob_$i.err
  python - "$cfg" gpurun_out/knob_$i.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    p = d["per_op_ms"]
    print("[%s] ms/step %.3f | gn %.3f sinkhorn %.3f transformer %.3f" % (sys.argv[1], d["ms_per_step"], p.get("gr_group_norm", 0), p.get("gr_sinkhorn", 0), p.get("gr_conditional_transformer", 0)))
except Exception as e:
    print("[%s] failed: %s" % (sys.argv[1], e))
PY
done
