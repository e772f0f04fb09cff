#!/bin/bash
# parity subset + A/B of env knobs: tools/gpu_r2d.sh "<pytest args>" "ENV.." "ENV.." ...
mkdir -p gpurun_out
sel="$1"; shift
(timeout 900 python -m pytest $sel -m gpu -x -q) > gpurun_out/r2d_pytest.log 2>&1
tail -6 gpurun_out/r2d_pytest.log
tools/gpu_ab.sh "$@"
