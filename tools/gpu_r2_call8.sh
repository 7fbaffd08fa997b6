#!/bin/bash
mkdir -p gpurun_out
for W in 1 0; do
  echo "== ICD_GEMM_WARP_EPI=$W"
  ICD_GEMM_WARP_EPI=$W timeout 600 python tools/gemm_epi_study.py 2>&1 | tee gpurun_out/gemm_epi_study_warp$W.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 1 -s 2 -o gpurun_out/r2_gemm_k320_warp -f python tools/gemm_epi_study.py --one > gpurun_out/ncu_k320.log 2>&1
tail -3 gpurun_out/ncu_k320.log
python tools/ncu_report_summary.py gpurun_out/r2_gemm_k320_warp.ncu-rep gpurun_out/r2_gemm_k320_warp_ncu_full.txt; cat gpurun_out/r2_gemm_k320_warp_ncu_full.txt | head -70
timeout 900 python -m pytest tests/test_text_gpu.py -m gpu -q -s -p no:cacheprovider 2>&1 | grep -v "^Endpoints" | tail -25 | cut -c1-400
