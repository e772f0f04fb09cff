#!/bin/bash
# full GPU parity suite, then the default bench line (as the driver runs it)
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/suite_pytest.log 2>&1
tail -4 gpurun_out/suite_pytest.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print("value %.2f e2e %.2f ms %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
print("roofline", d["roofline"]["kernel"][:60], d["roofline"]["frac"])
for k in ("config3_128_pairs_1gpu", "config5_200k_pair", "cpu_baseline", "gpu_torch_baseline"):
    print(k, json.dumps(d.get(k))[:700])
print(json.dumps(d["per_op_ms"]))
PY
