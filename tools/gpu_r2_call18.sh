#!/bin/bash
# speculative softmax (exps before the tile maximum): parity + A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -p no:cacheprovider -k "attention" 2>&1 | tail -3 | cut -c1-300
for S in 0 1; do
  echo "== ICD_ATTN_SPEC=$S"; ICD_ATTN_SPEC=$S timeout 300 python tools/attn_bench.py --iters 10 2>&1 | grep -v "Nk=77"
done
for S in 0 1; do
  echo "== bench sd15 ICD_ATTN_SPEC=$S"
  ICD_ATTN_SPEC=$S timeout 900 python bench.py --workload sd15 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-gpu 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d.get('roofline',{})
print('ms', round(d['ms_per_step'],3), 'value', round(d['value'],2), 'attn', r.get('attention',{}).get('ms_per_step_in_kernel'), r.get('attention',{}).get('frac'), 'clk', d.get('clocks'))"
done
