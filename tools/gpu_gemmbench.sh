#!/bin/bash
for cfg in "$@"; do echo "== $cfg"; env $cfg timeout 300 python tools/gemm_bench.py 2>&1 | tail -16; done
