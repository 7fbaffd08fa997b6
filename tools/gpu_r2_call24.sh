#!/bin/bash
# mbarrier suspend-time hint: A/B (hint = product build, nohint = variant build) on the same box, alternating
timeout 900 python -m pytest tests/test_kernels_gpu.py -x -q -p no:cacheprovider 2>&1 | tail -2 | cut -c1-200
run() { # label, workload, libpath-or-empty
  local out
  out=$(ICD_LIB_PATH=$3 timeout 900 python bench.py --workload $2 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-gpu 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms', round(d['ms_per_step'],3), 'value', round(d['value'],2), 'clk', d.get('clocks'))")
  echo "$1 $2: $out"
}
NH=$PWD/invertible_cd_b200/libicd_b200_nohint.so
for W in sd15 sdxl; do
  run hint $W ""; run nohint $W $NH; run hint $W ""; run nohint $W $NH
done
