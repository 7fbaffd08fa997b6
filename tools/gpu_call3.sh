#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
echo "== pytest"; tail -5 gpurun_out/pytest_gpu.log
timeout 120 ./tools/probes/mma_probe > gpurun_out/mma_probe.log 2>&1; cat gpurun_out/mma_probe.log
timeout 120 python tools/attn_bench.py > gpurun_out/mb_attn.log 2>&1; cat gpurun_out/mb_attn.log
timeout 120 python tools/gemm_bench.py > gpurun_out/mb_gemm.log 2>&1; cat gpurun_out/mb_gemm.log
timeout -k 10 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_sd15.json 2> gpurun_out/bench_sd15.err
echo "== bench"; python -c "
import json
for l in open('gpurun_out/bench_sd15.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['kernels'])"
tail -3 gpurun_out/bench_sd15.err
