#!/bin/bash
# ncu --set full captures of the kernels VERDICT r1 asked for (one launch each), summarised into gpurun_out/*.txt
mkdir -p gpurun_out
cap() {  # name, kernel regex, skip, command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -c 1 -s $skip -o gpurun_out/$name -f "$@" > gpurun_out/ncu_$name.log 2>&1
  python tools/ncu_report_summary.py gpurun_out/$name.ncu-rep gpurun_out/${name}_ncu_full.txt > /dev/null 2>&1
  python - <<PY
t=open("gpurun_out/${name}_ncu_full.txt").read()
h=len(t)//2
print("==== $name"); print(t[:h] if t[:h].strip()==t[h:].strip() else t)
PY
}
cap r2_gn_fused_64x64x320 gn_fused 3 python tools/gn_bench.py
cap r2_attn_self40 attention_tc 2 python tools/attn_bench.py --only 0
cap r2_attn_cross64 attention_tc 2 python tools/attn_bench.py --only 8
cap r2_gemm_k320_after gemm_tc_kernel 2 python tools/gemm_epi_study.py --one
cap r2_layernorm layernorm 3 python tools/gn_bench.py
