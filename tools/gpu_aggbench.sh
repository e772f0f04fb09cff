#!/bin/bash
mkdir -p gpurun_out
for cfg in "$@"; do echo "== $cfg"; env $cfg timeout 300 python tools/agg_bench.py 2>&1 | tail -12; done
