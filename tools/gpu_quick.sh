#!/bin/bash
# network + config parity tests, then short bench line(s) per env setting ("" = defaults)
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_network_gpu.py tests/test_configs_gpu.py -m gpu -x -q) > gpurun_out/quick_pytest.log 2>&1
tail -5 gpurun_out/quick_pytest.log
tools/gpu_ab.sh "${@:-GAUSSREG_X=0}"
