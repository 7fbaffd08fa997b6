#!/bin/bash
# Attention-only iteration: parity tests that touch the attention kernel + micro-benchmark (per POLY setting).
mkdir -p gpurun_out
for P in ${POLYS:-3}; do
  echo "== ICD_ATTN_POLY=$P"
  ICD_ATTN_POLY=$P timeout -k 10 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider -k "attention or attn" 2>&1 | tail -4
  ICD_ATTN_POLY=$P timeout 120 python tools/attn_bench.py 2>&1 | tee gpurun_out/mb_attn_poly$P.log
done
