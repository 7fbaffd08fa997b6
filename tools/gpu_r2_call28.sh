#!/bin/bash
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -p no:cacheprovider -k "layernorm" 2>&1 | tail -1
for K in 1000 8 4 2; do
echo "== ICD_LN_CPS=$K"
ICD_LN_CPS=$K timeout 300 python tools/gn_bench.py 2>&1 | grep "LN "
done
