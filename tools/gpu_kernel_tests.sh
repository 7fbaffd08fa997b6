#!/bin/bash
# Runs the kernel parity tests in separate processes (a hung kernel must not block the other groups).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for grp in "linear" "conv or downsample or upsample" "groupnorm or layernorm or softmax or embeddings" "explicit" "fused_self" "fused_cross"; do
  name=$(echo "$grp" | tr ' ' '_')
  timeout -k 10 420 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "$grp" -p no:cacheprovider --timeout 200 -rA 2>&1 | tail -60 > gpurun_out/kt_${name}.log
  echo "== $grp: exit $?" ; tail -15 gpurun_out/kt_${name}.log
done
