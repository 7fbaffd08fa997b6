#!/bin/bash
# short-context attention kernel + hot-swappable adapters: parity tests, micro-benchmark old vs new, step bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -p no:cacheprovider -k "attention" 2>&1 | tail -5 | cut -c1-300
timeout 900 python -m pytest tests/test_adapters_gpu.py -x -q -p no:cacheprovider -s 2>&1 | tail -12 | cut -c1-300
for C in "" "--no-capture"; do
  echo "== old kernel $C"; ICD_ATTN_SMALLKV=0 timeout 300 python tools/attn_bench.py --cross-only $C 2>&1 | tail -5
  echo "== new kernel $C"; timeout 300 python tools/attn_bench.py --cross-only $C 2>&1 | tail -5
  for T in 1 2 4 8; do echo "== new kernel TPC=$T $C"; ICD_ATTN_SMALLKV_TPC=$T timeout 300 python tools/attn_bench.py --cross-only $C 2>&1 | tail -5; done
done
