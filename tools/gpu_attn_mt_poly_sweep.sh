#!/bin/bash
# MT x POLY sweep with the speculative softmax on
for MT in 1 2; do for P in 0 2 3; do
  echo "== SPEC=1 MT=$MT POLY=$P"; ICD_ATTN_MT=$MT ICD_ATTN_POLY=$P timeout 300 python tools/attn_bench.py --iters 10 2>&1 | grep -v "Nk=77"
done; done
echo "== SPEC=0 defaults"; ICD_ATTN_SPEC=0 timeout 300 python tools/attn_bench.py --iters 10 2>&1 | grep -v "Nk=77"
