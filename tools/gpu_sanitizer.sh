#!/bin/bash
# compute-sanitizer: memcheck over ALL kernel parity tests + the fp32 path (kernels and one whole forward), racecheck over
# the shared-memory-heavy kernels; only kernels of namespace icd are instrumented
mkdir -p gpurun_out
# every tensor its own cudaMalloc: inside the caching allocator's 2 MB+ segments an overrun is invisible to memcheck
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
S="/usr/local/cuda/bin/compute-sanitizer --kernel-regex kns=icd"   # our kernels only: cuBLAS SIMT kernels of the torch references trip memcheck
timeout 1500 $S --tool memcheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider > gpurun_out/san_memcheck_kernels.log 2>&1
echo "memcheck kernels rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san_memcheck_kernels.log | tail -3
timeout 1500 $S --tool memcheck --error-exitcode 9 python -m pytest tests/test_f32_gpu.py -x -q -p no:cacheprovider -k "sgemm or attention or norms or embeddings or small_sd15" > gpurun_out/san_memcheck_f32.log 2>&1
echo "memcheck f32 rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san_memcheck_f32.log | tail -3
timeout 1200 $S --tool racecheck --error-exitcode 9 python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -k "short_context or cross_with_capture or geglu or groupnorm" > gpurun_out/san_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|passed|failed" gpurun_out/san_racecheck.log | tail -3
export ICD_CUDA_GRAPHS=0   # executor-level tests below: no allocation is possible inside a capture without the caching allocator
S="/usr/local/cuda/bin/compute-sanitizer --kernel-regex kns=icd"
for T in tests/test_unet_gpu.py tests/test_adapters_gpu.py tests/test_vae_gpu.py tests/test_text_gpu.py; do
  N=$(basename $T .py)
  timeout 1500 $S --tool memcheck --error-exitcode 9 python -m pytest $T -q -p no:cacheprovider -k "not graph" > gpurun_out/san_memcheck_$N.log 2>&1
  echo "$N rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san_memcheck_$N.log | tail -2; grep -c "Invalid __" gpurun_out/san_memcheck_$N.log
done
timeout 2400 $S --tool memcheck --error-exitcode 9 python -m pytest tests/test_fullsize_gpu.py -q -p no:cacheprovider -k "cfg0 or sdxl_full_row_forward or cfg1" > gpurun_out/san_memcheck_fullsize.log 2>&1
echo "fullsize rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san_memcheck_fullsize.log | tail -2; grep -c "Invalid __" gpurun_out/san_memcheck_fullsize.log
