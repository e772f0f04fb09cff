#!/bin/bash
mkdir -p gpurun_out
sel="$1"; shift
(timeout 900 python -m pytest $sel -m gpu -x -q) > gpurun_out/r2e_pytest.log 2>&1
tail -4 gpurun_out/r2e_pytest.log
python tools/tf_bench.py
tools/gpu_ab.sh "$@"
