#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
ICD_GEMM_WARP_EPI=0 timeout 1800 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_all_warp0.log 2>&1
echo "full suite (warp epi off) rc=$?"; tail -4 gpurun_out/pytest_all_warp0.log | cut -c1-300
cp gpurun_out/parity_report.jsonl gpurun_out/parity_report_warp0.jsonl 2>/dev/null
ICD_GEMM_WARP_EPI=1 timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_kern_warp1.log 2>&1
echo "kernels+unet (warp epi on) rc=$?"; tail -12 gpurun_out/pytest_kern_warp1.log | cut -c1-300
for W in 0 1; do
  for wl in sd15 sdxl; do
    ICD_GEMM_WARP_EPI=$W timeout 600 python bench.py --workload $wl --no-cpu-baseline --no-eager-gpu --steps 10 2>/dev/null | tail -1 > gpurun_out/bench_${wl}_warp$W.json
    python - <<PY
import json
d=json.loads(open("gpurun_out/bench_${wl}_warp$W.json").read())
r=d["roofline"]
print("$wl warp_epi=$W", "ms", round(d["ms_per_step"],2), "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "clk", d["clocks"],
      "| gemm frac", round(r["frac"],3), "ms", round(r["ms_per_step_in_kernel"],2), "| attn", round(r["attention"]["frac"],3), round(r["attention"]["ms_per_step_in_kernel"],2),
      "| gn", round(r["groupnorm"]["frac"],3), round(r["groupnorm"]["ms_per_step_in_kernel"],2), "| ln", round(r["layernorm"]["frac"],3), round(r["layernorm"]["ms_per_step_in_kernel"],2),
      "| other", round(r["other_kernels_ms"],2), "| default_store", d.get("default_store",{}).get("ms_per_step"))
PY
  done
done
