#!/bin/bash
for T in 640 960; do
echo "== GN_THREADS=$T"
ICD_LIB_PATH=$PWD/invertible_cd_b200/libicd_b200_gn$T.so timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -p no:cacheprovider -k "groupnorm" 2>&1 | tail -1
ICD_LIB_PATH=$PWD/invertible_cd_b200/libicd_b200_gnprof$T.so timeout 300 python tools/gn_prof.py 2>&1 | grep -v "320+320" | tail -3
ICD_LIB_PATH=$PWD/invertible_cd_b200/libicd_b200_gn$T.so timeout 300 python tools/gn_bench.py --all 2>&1 | grep GN
done
