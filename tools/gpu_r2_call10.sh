#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -m gpu -q -x -p no:cacheprovider > gpurun_out/pytest_kern_resmma.log 2>&1
echo "kernels+unet rc=$?"; tail -15 gpurun_out/pytest_kern_resmma.log | cut -c1-300
ICD_LIB_PATH=$PWD/invertible_cd_b200/libicd_b200_gprof.so timeout 600 python tools/gemm_prof.py 2>&1 | cut -c1-420 | tee gpurun_out/gemm_prof_resmma.log
timeout 600 python tools/gemm_epi_study.py 2>&1 | grep "bm=  0\|copy" | tee gpurun_out/gemm_epi_study_resmma.log
for W in 1 0; do
  for wl in sd15 sdxl; do
    ICD_GEMM_WARP_EPI=$W timeout 600 python bench.py --workload $wl --no-cpu-baseline --no-eager-gpu --steps 10 2>/dev/null | tail -1 > gpurun_out/bench_${wl}_warp$W.json
    python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${wl}_warp$W.json").read())
r=d["roofline"]
print("$wl warp_epi=$W", "ms", round(d["ms_per_step"],2), "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"],
      "| gemm frac", round(r["frac"],3), "ms", round(r["ms_per_step_in_kernel"],2), "| attn", round(r["attention"]["frac"],3), round(r["attention"]["ms_per_step_in_kernel"],2),
      "| gn", round(r["groupnorm"]["frac"],3), round(r["groupnorm"]["ms_per_step_in_kernel"],2), "| ln", round(r["layernorm"]["frac"],3), round(r["layernorm"]["ms_per_step_in_kernel"],2),
      "| other", round(r["other_kernels_ms"],2), "| default_store", d.get("default_store",{}).get("ms_per_step"))
PY
  done
done
