#!/bin/bash
# Round-2 GPU call 1: probes, full parity suite (numbers -> gpurun_out/parity_report.jsonl), bench line, ncu launch lists.
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv >> gpurun_out/host.txt
./tools/probes/coop_probe > gpurun_out/coop_probe.log 2>&1
cat gpurun_out/coop_probe.log
timeout 1500 python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
grep -h "rel-L2\|l2 \|\[parity\]\|\[smoke\]" gpurun_out/pytest_gpu.log | head -60
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_all.json 2> gpurun_out/bench_all.err
echo "bench rc=$?"; tail -c 6000 gpurun_out/bench_all.json; tail -5 gpurun_out/bench_all.err
TAG=r2_a
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
for W in sd15 sdxl; do
  timeout 900 ncu --nvtx --nvtx-include "icd_step/" --metrics $M --clock-control none --csv \
      --log-file gpurun_out/${TAG}_${W}_launches.csv python bench.py --workload $W --profile-step > gpurun_out/prof_$W.log 2>&1
  tail -2 gpurun_out/prof_$W.log
  python tools/ncu_launch_summary.py gpurun_out/${TAG}_${W}_launches.csv gpurun_out/step_shapes_$W.json $W gpurun_out/${TAG}_traffic.json > gpurun_out/${TAG}_${W}_launch_summary.txt 2>&1
  head -75 gpurun_out/${TAG}_${W}_launch_summary.txt
done
