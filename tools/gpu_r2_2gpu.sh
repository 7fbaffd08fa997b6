#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.out 2> gpurun_out/bench_2gpu.err
echo "rc=$?"; tail -1 gpurun_out/bench_2gpu.out > gpurun_out/bench_2gpu.json; tail -5 gpurun_out/bench_2gpu.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_2gpu.json").read())
print("sd15 N=2", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["clocks"])
print("sdxl_cfg3", {k:d["sdxl_cfg3"][k] for k in ("per_gpu_batch","global_batch","value","ms_per_step","scaling")})
print([k for k in d.keys()])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
