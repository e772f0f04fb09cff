#!/bin/bash
# compute-sanitizer passes over the GPU parity suite (race detection / memory checking).
#   tools/sanitize.sh [memcheck|racecheck|initcheck|synccheck ...]   (default: memcheck racecheck)
# Run on a GPU box (gpurun -- tools/sanitize.sh); reports go to gpurun_out/sanitize_<tool>.log and a one-line summary each.
# The selection below covers every C-ABI entry point once at small sizes (the sanitizer slows kernels 10-100x).
mkdir -p gpurun_out
tools=${@:-memcheck racecheck}
sel="tests/test_neighbors_gpu.py tests/test_network_gpu.py tests/test_ransac.py tests/test_gaussians_gpu.py"
filt="not full_size and not config5 and not 60000 and not 100000 and not test_gemm_tensor_core and not config3"
for t in $tools; do
  GAUSSREG_SANITIZE=1 timeout 3000 compute-sanitizer --tool $t --error-exitcode 99 --print-limit 20 \
    python -m pytest $sel -m gpu -x -q -k "$filt" > gpurun_out/sanitize_$t.log 2>&1
  rc=$?
  echo "[$t] exit=$rc  $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/sanitize_$t.log | tail -3 | tr '\n' ' ')"
done
