#!/bin/bash
mkdir -p gpurun_out
( timeout 100 python tools/gemm_bench.py
for bm in 128 256; do for bn in 64 128 160 256; do echo "== bm=$bm bn=$bn"; for i in 0 1 2 3 6; do timeout 60 python tools/gemm_bench.py --only $i --bm $bm --bn $bn 2>&1 | tail -1; done; done; done ) > gpurun_out/mb_gemm_sweep.log 2>&1
cat gpurun_out/mb_gemm_sweep.log
timeout 120 python tools/gn_bench.py > gpurun_out/mb_gn.log 2>&1; cat gpurun_out/mb_gn.log
