#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/gemm_epi_study.py 2>&1 | tee gpurun_out/gemm_tile_study_r2.log | cut -c1-200
