#!/bin/bash
# neighbour + network + config parity tests, short bench line(s) per env setting ("" = defaults), host/device timeline
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_neighbors_gpu.py tests/test_network_gpu.py tests/test_configs_gpu.py -m gpu -x -q) > gpurun_out/quick_pytest.log 2>&1
tail -4 gpurun_out/quick_pytest.log
tools/gpu_ab.sh "${@:-GAUSSREG_X=0}"
tools/gpu_timeline.sh "GAUSSREG_X=0" > /dev/null
sed -n 17,30p gpurun_out/timeline_1.txt
