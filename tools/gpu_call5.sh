#!/bin/bash
# Baseline of the committed state: parity tests, smoke, SD1.5 + SDXL bench lines, micro-benchmarks.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout -k 10 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
echo "== pytest"; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout -k 10 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_sd15.json 2> gpurun_out/bench_sd15.err
echo "== bench sd15"; python - <<'PY'
import json
for l in open('gpurun_out/bench_sd15.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e'], d['clocks'], d['gpu_launches_per_step']); print(d['kernels']); print(d['roofline']); print(d['cpu_baseline'])
PY
tail -3 gpurun_out/bench_sd15.err
timeout -k 10 400 python bench.py --workload sdxl --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sdxl.json 2> gpurun_out/bench_sdxl.err
echo "== bench sdxl"; python - <<'PY'
import json
for l in open('gpurun_out/bench_sdxl.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e'], d['gpu_launches_per_step']); print(d['kernels']); print(d['roofline'])
PY
tail -3 gpurun_out/bench_sdxl.err
timeout 120 python tools/attn_bench.py > gpurun_out/mb_attn.log 2>&1; cat gpurun_out/mb_attn.log
timeout 120 python tools/gemm_bench.py > gpurun_out/mb_gemm.log 2>&1; cat gpurun_out/mb_gemm.log
timeout 120 python tools/gn_bench.py > gpurun_out/mb_gn.log 2>&1; cat gpurun_out/mb_gn.log
timeout 120 ./tools/probes/mma_probe > gpurun_out/mma_probe.log 2>&1; tail -40 gpurun_out/mma_probe.log
