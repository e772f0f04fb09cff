#!/bin/bash
# full GPU parity suite, then one short bench line (per-op times on stderr-free JSON)
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/suite_pytest.log 2>&1
tail -6 gpurun_out/suite_pytest.log
tools/gpu_ab.sh "$@"
