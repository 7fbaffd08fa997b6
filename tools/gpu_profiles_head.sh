#!/bin/bash
# HEAD (r2_c) profiles: ncu launch lists with DRAM bytes for both workloads, --set full captures of the new / changed kernels
mkdir -p gpurun_out
TAG=r2_c
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
for W in sd15 sdxl; do
  timeout 900 ncu --nvtx --nvtx-include "icd_step/" --metrics $M --clock-control none --csv \
      --log-file gpurun_out/${TAG}_${W}_launches.csv python bench.py --workload $W --profile-step > gpurun_out/prof_$W.log 2>&1
  python tools/ncu_launch_summary.py gpurun_out/${TAG}_${W}_launches.csv gpurun_out/step_shapes_$W.json $W gpurun_out/${TAG}_traffic.json > gpurun_out/${TAG}_${W}_launch_summary.txt 2>&1
  head -30 gpurun_out/${TAG}_${W}_launch_summary.txt | cut -c1-200
done
cap() {  # name, kernel regex, skip, command...
  local name=$1 rx=$2 skip=$3; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -c 1 -s $skip -o gpurun_out/$name -f "$@" > gpurun_out/ncu_$name.log 2>&1
  python tools/ncu_report_summary.py gpurun_out/$name.ncu-rep gpurun_out/${name}_ncu_full.txt > /dev/null 2>&1
  echo "==== $name"; head -32 gpurun_out/${name}_ncu_full.txt | cut -c1-220
}
cap r2_c_attn_self40_spec attention_tc 2 python tools/attn_bench.py --only 0
cap r2_c_attn_self64_spec attention_tc 2 python tools/attn_bench.py --only 6
cap r2_c_attn_smallkv_cross64 attention_smallkv 2 python tools/attn_bench.py --only 8 --no-capture
cap r2_c_attn_smallkv_cross40_capture attention_smallkv 2 python tools/attn_bench.py --only 3
rm -f gpurun_out/*.ncu-rep
