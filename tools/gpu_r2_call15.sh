#!/bin/bash
# short-context attention: 2 vs 3 resident CTAs; LoRA fuse precision study; adapter tests; step benches
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -p no:cacheprovider -k "attention" 2>&1 | tail -3 | cut -c1-300
for N in 2 3; do for C in "" "--no-capture"; do
  echo "== smallkv CTAS=$N $C"; ICD_ATTN_SMALLKV_CTAS=$N timeout 300 python tools/attn_bench.py --cross-only $C 2>&1 | tail -5
done; done
timeout 300 python tools/lora_fuse_study.py 2>&1 | tail -3
timeout 900 python -m pytest tests/test_adapters_gpu.py -q -p no:cacheprovider -s 2>&1 | grep -E "rel-L2|ulps|passed|failed|Error|assert" | cut -c1-300
for W in sd15 sdxl; do
  for E in 0 1; do
    echo "== bench $W ICD_ATTN_SMALLKV=$E"
    ICD_ATTN_SMALLKV=$E timeout 900 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline --no-eager-gpu 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d.get('roofline',{})
print('ms', round(d['ms_per_step'],3), 'value', round(d['value'],2), 'e2e', d.get('e2e',{}).get('value'), 'attn', r.get('attention',{}).get('ms_per_step_in_kernel'), r.get('attention',{}).get('frac'), 'clk', d.get('clocks'))"
  done
done
