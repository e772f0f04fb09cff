#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_network_gpu.py -m gpu -x -q -k "structure_embedding or transformer or full_forward") > gpurun_out/t1_pytest.log 2>&1
tail -15 gpurun_out/t1_pytest.log
timeout 300 python tools/t1_bench.py 479 1024 4200 2>&1 | tee gpurun_out/t1_bench.txt
tools/gpu_ab.sh "GAUSSREG_T1=table" "GAUSSREG_T1=tc"
