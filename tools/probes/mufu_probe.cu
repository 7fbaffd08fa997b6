// Micro-probe: MUFU.EX2 issue rate per SM sub-partition for f32 and packed f16x2 operands, with 1..4 warps per
// sub-partition issuing independent ops (does ex2.approx.f16x2 deliver two exponentials per MUFU slot?).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/probes/mufu_probe.cu -o tools/probes/mufu_probe
#include <cstdio>
#include <cuda_fp16.h>

template <int MODE>   // 0: ex2.approx.ftz.f32, 1: ex2.approx.ftz.f16x2, 2: ex2.approx.f16x2
__global__ void probe(int reps, float seed, long long* out, float* sink) {
  float x[8];
  unsigned h[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = seed * (i + 1) + threadIdx.x * 1e-3f; h[i] = 0x38003800u + i + threadIdx.x; }
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
    }
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += x[i] + __uint_as_float(h[i]);
  if (acc == 1.2345f) sink[0] = acc;
  if ((threadIdx.x & 31) == 0) out[blockIdx.x * 32 + (threadIdx.x >> 5)] = t1 - t0;
}

template <int MODE>
static void run(const char* name, int warps, long long* out, float* sink) {
  const int reps = 4096;
  probe<MODE><<<148, warps * 32>>>(reps, 0.01f, out, sink);
  cudaDeviceSynchronize();
  long long mx = 0;
  for (int i = 0; i < 148 * 32; ++i) mx = out[i] > mx ? out[i] : mx;
  const double cyc_per_op = double(mx) / (reps * 8);
  printf("%-22s warps/SM %2d (%d per sub-partition): %6.2f cyc per warp-op, %6.2f cyc per op per sub-partition\n", name,
         warps, (warps + 3) / 4, cyc_per_op, cyc_per_op / ((warps + 3) / 4));
}

int main() {
  long long* out;
  float* sink;
  cudaMallocManaged(&out, 148 * 32 * sizeof(long long));
  cudaMallocManaged(&sink, 16);
  for (int w : {4, 8, 16}) {
    for (int i = 0; i < 148 * 32; ++i) out[i] = 0;
    run<0>("ex2.approx.ftz.f32", w, out, sink);
    for (int i = 0; i < 148 * 32; ++i) out[i] = 0;
    run<1>("ex2.approx.f16x2", w, out, sink);
    for (int i = 0; i < 148 * 32; ++i) out[i] = 0;
    run<2>("ex2.approx.ftz.bf16x2", w, out, sink);
  }
  return 0;
}
