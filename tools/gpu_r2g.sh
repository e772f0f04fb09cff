#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_network_gpu.py -m gpu -x -q -k "gemm or unary or residual or full_forward") > gpurun_out/r2g_pytest.log 2>&1
tail -4 gpurun_out/r2g_pytest.log
tools/gpu_gemmbench.sh "$@"
