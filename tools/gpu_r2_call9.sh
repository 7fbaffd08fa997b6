#!/bin/bash
mkdir -p gpurun_out
for D in 0 7 15 23 39 63 8 16 32; do
  echo "#### ICD_EPI_DEBUG=$D (1 no TMA store, 2 no tcgen05.ld, 4 no bias loads, 8 no fence, 16 no st.shared, 32 no bulk_wait/syncwarp)"
  ICD_EPI_DEBUG=$D ICD_LIB_PATH=$PWD/invertible_cd_b200/libicd_b200_gprof.so timeout 600 python tools/gemm_prof.py 2>&1 | grep -A1 "320 plain\|1280 bias" | cut -c1-420
done | tee gpurun_out/gemm_prof_dbg2.log
