#!/bin/bash
# round-2 final evidence: full GPU suite, warm launch list of one bench step, ncu --set full of the tabulated T1 kernel and the
# replay kernel, full bench line
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/suite_pytest.log 2>&1
tail -3 gpurun_out/suite_pytest.log
tools/gpu_launches.sh ""
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"structure_embedding_table|hash_order_replay" -c 6 -f -o gpurun_out/r02_full_t1tab \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-throughput > gpurun_out/r02_full_t1tab.log 2>&1
ncu -i gpurun_out/r02_full_t1tab.ncu-rep --page raw --csv > gpurun_out/r02_full_t1tab_raw.csv 2>/dev/null
ls -la gpurun_out/r02_full_t1tab.ncu-rep
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -2 gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("value %.2f e2e %.2f ms %.3f launches %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["gpu_launches"]))
print("config3", json.dumps(d["config3_128_pairs_1gpu"])[:400])
print("config5", d["config5_200k_pair"]["ms_per_pair"])
print("roofline", d["roofline"]["frac"], d["roofline"].get("tf32_peak_measured"))
print(json.dumps(d["per_op_ms"]))
PY
