#!/bin/bash
# full suite, then focused sanitizer re-runs, then the bench line
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r2f_pytest.log 2>&1
tail -5 gpurun_out/r2f_pytest.log
timeout 600 compute-sanitizer --tool initcheck --error-exitcode 99 --print-limit 5 python -m pytest tests/test_neighbors_gpu.py -m gpu -x -q -k "golden or ragged" > gpurun_out/sanitize_initcheck_rerun.log 2>&1
echo "[initcheck rerun] $(grep -E 'ERROR SUMMARY|passed|failed' gpurun_out/sanitize_initcheck_rerun.log | tail -2 | tr '\n' ' ')"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 99 --print-limit 5 python -m pytest tests/test_network_gpu.py -m gpu -x -q -k "unary_block and 5003" > gpurun_out/sanitize_racecheck_rerun.log 2>&1
echo "[racecheck rerun] $(grep -E 'RACECHECK SUMMARY|passed|failed' gpurun_out/sanitize_racecheck_rerun.log | tail -2 | tr '\n' ' ')"
tools/gpu_bench.sh ""
