#!/bin/bash
# 2-GPU check of the bench contract at HEAD (torchrun, NCCL), both arms
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus ${NG:-2} --steps 5 --warmup 3 --no-eager-gpu > gpurun_out/bench_2gpu.out 2> gpurun_out/bench_2gpu.err
echo "rc=$?"; tail -1 gpurun_out/bench_2gpu.out | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('n_gpus', d['n_gpus'], 'value', round(d['value'],2), 'ms', round(d['ms_per_step'],2), 'e2e', d['e2e']['value'], 'scaling', d['scaling'])
print('sdxl_cfg3', json.dumps(d.get('sdxl_cfg3'))[:500])"
tail -3 gpurun_out/bench_2gpu.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus ${NG:-2} --steps 1 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
