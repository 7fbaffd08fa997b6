#!/bin/bash
# One GPU call: parity tests, bench, micro-benchmarks, ncu --set full captures of the three dominant kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout -k 10 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
echo "== pytest"; tail -5 gpurun_out/pytest_gpu.log
timeout -k 10 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_sd15.json 2> gpurun_out/bench_sd15.err
echo "== bench"; cut -c1-400 gpurun_out/bench_sd15.json
timeout 120 python tools/attn_bench.py > gpurun_out/mb_attn.log 2>&1; cat gpurun_out/mb_attn.log
timeout 120 python tools/gemm_bench.py > gpurun_out/mb_gemm.log 2>&1; cat gpurun_out/mb_gemm.log
timeout 120 python tools/gn_bench.py > gpurun_out/mb_gn.log 2>&1; cat gpurun_out/mb_gn.log
NCU="ncu --set full --clock-control none --import-source on --launch-count 1 -f"
timeout 300 $NCU -k regex:attention_tc --launch-skip 2 -o gpurun_out/r1f_attn_self64 python tools/attn_bench.py --only 0 --iters 2 > gpurun_out/ncu_attn.log 2>&1
timeout 300 $NCU -k regex:attention_tc --launch-skip 2 -o gpurun_out/r1f_attn_sdxl64 python tools/attn_bench.py --only 4 --iters 2 >> gpurun_out/ncu_attn.log 2>&1
timeout 300 $NCU -k regex:gemm_tc --launch-skip 3 -o gpurun_out/r1f_gemm_k320 python tools/gemm_bench.py --only 0 --iters 2 > gpurun_out/ncu_gemm.log 2>&1
timeout 300 $NCU -k regex:gemm_tc --launch-skip 3 -o gpurun_out/r1f_gemm_conv640 python tools/gemm_bench.py --only 8 --iters 2 >> gpurun_out/ncu_gemm.log 2>&1
timeout 300 $NCU -k regex:gn_apply --launch-skip 3 -o gpurun_out/r1f_gn_apply python tools/gn_bench.py > gpurun_out/ncu_gn.log 2>&1
ls -la gpurun_out/*.ncu-rep
