#!/bin/bash
# ncu launch lists (duration + DRAM bytes per launch) of one eager iCD step, SD1.5 and SDXL, joined with shapes.
mkdir -p gpurun_out
M="${METRICS:-gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum}"
for W in sd15 sdxl; do
  timeout 900 ncu --nvtx --nvtx-include "icd_step/" --metrics $M --clock-control none --csv \
      --log-file gpurun_out/${TAG}_${W}_launches.csv python bench.py --workload $W --profile-step > gpurun_out/prof_$W.log 2>&1
  tail -2 gpurun_out/prof_$W.log
  python tools/ncu_launch_summary.py gpurun_out/${TAG}_${W}_launches.csv gpurun_out/step_shapes_$W.json > gpurun_out/${TAG}_${W}_launch_summary.txt 2>&1
  head -60 gpurun_out/${TAG}_${W}_launch_summary.txt
done
