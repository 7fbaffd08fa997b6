#!/bin/bash
for K in "4,4" "2,4" "1,4" "4,2" "4,1" "2,2" "1,1" "1,2" "8,8"; do
echo "== ICD_GN2_CPS=$K"
ICD_GN2_CPS=$K timeout 300 python tools/gn_bench.py --all 2>&1 | grep -E "B=8 HW= 4096 C=640\+320|B=8 HW= 4096 C=320\+320|B=4 HW=16384|B=8 HW= 1024 C=1280\+640"
done
