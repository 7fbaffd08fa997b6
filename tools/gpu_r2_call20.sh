#!/bin/bash
# attention softmax phase profile (instrumented build) with and without speculation
for S in 0 1; do
  echo "== SPEC=$S"; ICD_ATTN_SPEC=$S ICD_LIB_PATH=$PWD/invertible_cd_b200/libicd_b200_prof.so timeout 300 python tools/attn_prof.py 2>&1 | tail -8
done
timeout 300 python tools/attn_bench.py --iters 10 2>&1 | grep -v "Nk=77"
