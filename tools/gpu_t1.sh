#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_network_gpu.py tests/test_configs_gpu.py -m gpu -x -q) > gpurun_out/t1_pytest.log 2>&1
tail -5 gpurun_out/t1_pytest.log
timeout 300 python tools/t1_bench.py 479 4200 2>&1 | tee gpurun_out/t1_bench.txt
tools/gpu_bench.sh ""
