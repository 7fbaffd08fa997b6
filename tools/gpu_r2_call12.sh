#!/bin/bash
# full suite in ONE process (as the driver runs it), smoke, the default bench line, ncu launch lists + traffic for HEAD
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 1800 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1
echo "full suite rc=$?"; tail -4 gpurun_out/pytest_all.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_all.out 2> gpurun_out/bench_all.err
echo "bench rc=$?"; tail -1 gpurun_out/bench_all.out > gpurun_out/bench_all.json; tail -3 gpurun_out/bench_all.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_all.json").read())
def brief(name, r):
    ro=r.get("roofline",{})
    print(name, "ms", round(r["ms_per_step"],2), "value", round(r["value"],2), "e2e", r.get("e2e",{}).get("value"), "clk", (r.get("clocks") or {}).get("sm_mhz"), (r.get("clocks") or {}).get("reasons"))
    if ro:
        print("   gemm", ro.get("frac"), ro.get("ms_per_step_in_kernel"), "attn", ro["attention"]["frac"], ro["attention"]["ms_per_step_in_kernel"], "gn", ro["groupnorm"]["frac"], ro["groupnorm"]["ms_per_step_in_kernel"], "ln", ro["layernorm"]["frac"], ro["layernorm"]["ms_per_step_in_kernel"], "other", ro["other_kernels_ms"], "whole", ro["whole_step_frac"])
    for k in ("default_store","eager_gpu","cpu_baseline"):
        if k in r: print("   ", k, json.dumps(r[k])[:400])
brief("sd15", d); brief("sdxl", d["sdxl"]); brief("sdxl_cfg3", d["sdxl_cfg3"]); print(json.dumps(d.get("small_batch"))[:1500])
PY
TAG=r2_b
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
for W in sd15 sdxl; do
  timeout 900 ncu --nvtx --nvtx-include "icd_step/" --metrics $M --clock-control none --csv \
      --log-file gpurun_out/${TAG}_${W}_launches.csv python bench.py --workload $W --profile-step > gpurun_out/prof_$W.log 2>&1
  python tools/ncu_launch_summary.py gpurun_out/${TAG}_${W}_launches.csv gpurun_out/step_shapes_$W.json $W gpurun_out/${TAG}_traffic.json > gpurun_out/${TAG}_${W}_launch_summary.txt 2>&1
  head -48 gpurun_out/${TAG}_${W}_launch_summary.txt | cut -c1-200
done
