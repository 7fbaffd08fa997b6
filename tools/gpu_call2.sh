#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
echo "== pytest"; tail -5 gpurun_out/pytest_gpu.log
timeout 120 ./tools/probes/mma_probe > gpurun_out/mma_probe.log 2>&1; cat gpurun_out/mma_probe.log
timeout 120 python tools/attn_bench.py > gpurun_out/mb_attn.log 2>&1; cat gpurun_out/mb_attn.log
ICD_PDL=0 timeout -k 10 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sd15_nopdl.json 2> gpurun_out/bench_sd15_nopdl.err
echo "== bench no pdl"; python -c "
import json
for l in open('gpurun_out/bench_sd15_nopdl.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
timeout -k 10 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_sd15.json 2> gpurun_out/bench_sd15.err
echo "== bench pdl"; python -c "
import json
for l in open('gpurun_out/bench_sd15.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
tail -3 gpurun_out/bench_sd15.err
