#!/bin/bash
# full suite in one process + smoke + default bench at HEAD
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
S=$(date +%s)
timeout 1700 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/pytest_all.log 2>&1
echo "full suite rc=$? wall=$(( $(date +%s) - S ))s"; tail -4 gpurun_out/pytest_all.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
S=$(date +%s)
timeout 1200 python bench.py > gpurun_out/bench_default.out 2> gpurun_out/bench_default.err
echo "bench rc=$? wall=$(( $(date +%s) - S ))s"; tail -1 gpurun_out/bench_default.out > gpurun_out/bench_default.json; tail -2 gpurun_out/bench_default.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_default.json").read())
def brief(name, r):
    ro=r.get("roofline",{})
    print(name, "ms", round(r["ms_per_step"],2), "value", round(r["value"],2), "e2e", r.get("e2e",{}).get("value"), "clk", (r.get("clocks") or {}).get("sm_mhz"), (r.get("clocks") or {}).get("reasons"))
    if ro:
        print("   gemm", ro.get("frac"), ro.get("ms_per_step_in_kernel"), "attn", ro["attention"]["frac"], ro["attention"]["ms_per_step_in_kernel"], "gn", ro["groupnorm"]["frac"], ro["groupnorm"]["ms_per_step_in_kernel"], "ln", ro["layernorm"]["frac"], ro["layernorm"]["ms_per_step_in_kernel"], "other", ro["other_kernels_ms"], "whole", ro["whole_step_frac"])
    for k in ("default_store","eager_gpu","cpu_baseline"):
        if k in r: print("   ", k, json.dumps(r[k])[:300])
brief("sd15", d); brief("sdxl", d["sdxl"]); brief("sdxl_cfg3", d["sdxl_cfg3"]); print(json.dumps(d.get("small_batch"))[:1200])
PY
