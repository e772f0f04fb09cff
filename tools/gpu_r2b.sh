#!/bin/bash
# round-2 pass B: full GPU parity suite (fail fast) + bench A/B of the native backbone
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q "$@") > gpurun_out/r2b_pytest.log 2>&1
tail -15 gpurun_out/r2b_pytest.log
