#!/bin/bash
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -p no:cacheprovider -k "groupnorm" 2>&1 | tail -2
for K in 0 1 2 3; do
echo "== ICD_GN_BULK=$K"
ICD_GN_BULK=$K ICD_LIB_PATH=$PWD/invertible_cd_b200/libicd_b200_gnprof.so timeout 300 python tools/gn_prof.py 2>&1 | grep -v "320+320" | tail -3
ICD_GN_BULK=$K timeout 300 python tools/gn_bench.py --all 2>&1 | grep -E "GN  B=8|GN  B=4"
done
