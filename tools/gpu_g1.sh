#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_neighbors_gpu.py -m gpu -x -q) > gpurun_out/g1_pytest.log 2>&1
tail -4 gpurun_out/g1_pytest.log
tools/gpu_ab.sh "GAUSSREG_REPLAY_SCAN=1" "GAUSSREG_REPLAY_SCAN=0"
