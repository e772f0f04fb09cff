#!/bin/bash
# round-2 pass A: parity of the KPConv aggregation variants + per-layer timing
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_network_gpu.py -m gpu -x -q -k "kpconv") > gpurun_out/r2a_pytest.log 2>&1
tail -5 gpurun_out/r2a_pytest.log
tools/gpu_aggbench.sh "$@"
