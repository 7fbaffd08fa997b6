#!/bin/bash
# Sweep of the attention tuning knobs (query tiles per CTA x polynomial share) on the d=40 / d=64 shapes.
mkdir -p gpurun_out
: > gpurun_out/mb_attn_sweep.log
for MT in 1 2; do for P in 0 2 3; do
  echo "== ICD_ATTN_MT=$MT ICD_ATTN_POLY=$P" | tee -a gpurun_out/mb_attn_sweep.log
  for i in 0 3 6 7 8; do ICD_ATTN_MT=$MT ICD_ATTN_POLY=$P timeout 60 python tools/attn_bench.py --only $i 2>&1 | tail -1 | tee -a gpurun_out/mb_attn_sweep.log; done
done; done
timeout -k 10 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -p no:cacheprovider -k "attention or attn" 2>&1 | tail -3
