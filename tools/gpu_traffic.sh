#!/bin/bash
# ncu --set full of the hot kernels of one bench step (warm-up step of `bench.py --steps 1 --warmup 1`), three passes:
#   A: the dominant kernel (T1) and the other single-launch heavyweights   B: backbone GEMMs + KPConv gather (first stages)
#   C: the wide one-tile GEMM of the later stages.  Raw pages are exported next to the reports;
# tools/traffic_json.py turns them into profiles/r02_traffic.json
mkdir -p gpurun_out
run() {  # name regex count [skip]
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$2" -s ${4:-0} -c $3 -f -o gpurun_out/$1 \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-throughput > gpurun_out/$1.log 2>&1
  tail -1 gpurun_out/$1.log | cut -c1-160
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ls -la gpurun_out/$1.ncu-rep
}
run r02_full_a "structure_embedding_f16|sinkhorn_scaling128|rpe_scores_softmax_v2|cross_attention|lgr_refine|lgr_correspondence" ${1:-12}
run r02_full_b "gemm_tf32x3_persist|kpconv_aggregate_cp" ${2:-12}
run r02_full_c "gemm_tf32x3_tma_kernel<256>|hash_order_replay" ${3:-8} 2
