// Micro-probe: tensor-memory load bandwidth (tcgen05.ld 32x32b.x32 = 4 KB per warp instruction) as a function of
// the number of warps issuing loads, and how much a concurrent tcgen05.mma stream (TS mode, N=64: the attention
// kernel's MMAs) slows the loads / is slowed by them.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I invertible_cd_b200/csrc tools/probes/tmem_probe.cu -o tools/probes/tmem_probe
#include <cstdio>
#include <cstdlib>
#include "icd_ptx.cuh"
using namespace icd;

// warps 0..NW-1 load; warp 8 issues MMAs when with_mma
__global__ void __launch_bounds__(288) probe(int nw, int reps, int with_mma, int mma_reps, int batch, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 0 && lane == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 1) tmem_alloc<512>(&tmem_ptr);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_ptr;
  long long dt = 0;
  if (warp < nw) {
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    float acc = 0.f;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      float v[32], w[32];
      tmem_ld32(tb + lane_off + ((r & 1) * 64) + (warp >> 2) * 128, v);
      if (batch == 2) tmem_ld32(tb + lane_off + ((r & 1) * 64) + 32 + (warp >> 2) * 128, w);
      tmem_ld_wait();
      acc += v[0] + v[31];
      if (batch == 2) acc += w[0] + w[31];
    }
    dt = clock64() - t0;
    if (acc == 123.456f) out[1023] = 1;
    if (lane == 0) out[blockIdx.x * 16 + warp] = dt;
  } else if (warp == 8 && with_mma) {
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_f16(128, 64, false, false);
      const uint64_t hi = static_cast<uint64_t>((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;
      const uint32_t b_lo = ((smem_u32(smem) >> 4) & 0x3FFFu) | (1u << 16);
      const long long t0 = clock64();
      for (int r = 0; r < mma_reps; ++r) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ts(tb + 256 + (r & 1) * 64, tb + 480 + k * 8, hi | (b_lo + k * 2u), idesc, 1u);
      }
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      out[blockIdx.x * 16 + 8] = clock64() - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc<512>(tb); }
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 4096 * sizeof(long long));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int reps = 2048, mma_reps = 1024;
  printf("%6s %8s | %12s %14s | %12s\n", "warps", "mma", "cyc/LDTM.x32", "B/clk/SM (ld)", "cyc/MMA(N64)");
  for (int batch = 1; batch <= 2; ++batch)
  for (int with_mma = 0; with_mma <= 1; ++with_mma)
    for (int nw : {0, 1, 2, 4, 8}) {
      if (nw == 0 && !with_mma) continue;
      for (int i = 0; i < 4096; ++i) out[i] = 0;
      probe<<<148, 288, 64 * 1024>>>(nw, reps, with_mma, mma_reps, batch, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
      long long mx = 0, mm = 0;
      for (int b = 0; b < 148; ++b) {
        for (int w = 0; w < nw; ++w) mx = out[b * 16 + w] > mx ? out[b * 16 + w] : mx;
        mm = out[b * 16 + 8] > mm ? out[b * 16 + 8] : mm;
      }
      const double cyc = nw ? double(mx) / (reps * batch) : 0.0;
      printf("%6d %5s b%d | %12.1f %14.1f | %12.1f\n", nw, with_mma ? "yes" : "no", batch, cyc, nw ? nw * 4096.0 / cyc : 0.0,
             with_mma ? double(mm) / (mma_reps * 4) : 0.0);
    }
  return 0;
}
