// Micro-probe: issue cost of tcgen05.mma (kind::f16, M=128, cta_group::1) as a function of N, of the number of
// independent accumulators the issue stream round-robins over, of co-resident CTAs per SM, and of HOW the issuing
// thread is selected: `lane == 0` (divergent branch: ptxas wraps every UTCHMMA in an ELECT/BRA.U.ANY waterfall loop)
// versus `elect.sync` (uniform datapath, no loop).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I invertible_cd_b200/csrc tools/probes/mma_probe.cu -o mma_probe
#include <cstdio>
#include <cstdlib>
#include "icd_ptx.cuh"
using namespace icd;

// MODE 0: A and B from shared memory (SS). MODE 1: A from tensor memory (TS), B K-major. MODE 2: TS, B MN-major.
// MODE 3: the attention kernel's per-key-tile sequence: 4 x (TS N=64) + 4 x (TS N=N MN-major B + TS N=16).
template <int N_ACC, bool ELECT, int MODE = 0>
__global__ void __launch_bounds__(128) probe(int N, int reps, int tmem_cols, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 0 && lane == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 1) {
    if (tmem_cols == 512) tmem_alloc<512>(&tmem_ptr); else tmem_alloc<256>(&tmem_ptr);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_ptr;
  if (warp == 0) {
    const bool me = ELECT ? elect_one() : (lane == 0);
    if (me) {
      const uint32_t idesc = umma_idesc_f16(128, N, false, MODE >= 2);
      const uint32_t idesc64 = umma_idesc_f16(128, 64, false, false), idesc16 = umma_idesc_f16(128, 16, false, false);
      const uint32_t ta = tb + tmem_cols - 32;   // A operand columns (packed half2), garbage values: timing only
      const uint64_t hi = static_cast<uint64_t>((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;
      const uint32_t a_lo = ((smem_u32(smem) >> 4) & 0x3FFFu) | (1u << 16);
      const uint32_t b_lo = ((smem_u32(smem + 16384) >> 4) & 0x3FFFu) | (1u << 16);
      const int stride = (N + 31) / 32 * 32;
      long long best = 1LL << 60;
      for (int trial = 0; trial < 3; ++trial) {
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
          if (MODE == 3) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ts(tb, ta + k * 8, hi | (b_lo + k * 2u), idesc64, 1u);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_f16_ts(tb + 64, ta + k * 8, hi | (a_lo + k * 128u), idesc, 1u);
              umma_f16_ts(tb + 64 + 96, ta + k * 8, hi | (b_lo + k * 2u), idesc16, 1u);
            }
            continue;
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (MODE == 0)
              umma_f16_ss(tb + ((r * 4 + k) & (N_ACC - 1)) * stride, hi | (a_lo + k * 2u), hi | (b_lo + k * 2u), idesc, 1u);
            else
              umma_f16_ts(tb + ((r * 4 + k) & (N_ACC - 1)) * stride, ta + k * 8,
                          hi | (b_lo + (MODE == 2 ? k * 128u : k * 2u)), idesc, 1u);
          }
        }
        umma_commit(&bar);
        mbar_wait(&bar, trial & 1);
        const long long t1 = clock64();
        if (t1 - t0 < best) best = t1 - t0;
      }
      out[blockIdx.x] = best;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (tmem_cols == 512) tmem_dealloc<512>(tb); else tmem_dealloc<256>(tb);
  }
}

template <int N_ACC, bool ELECT, int MODE = 0>
static void run(int N, int ctas, long long* out) {
  const int reps = 256;
  const int stride = (N + 31) / 32 * 32;
  const int cols = ctas == 1 ? 512 : 256;
  if (N_ACC * stride > cols - 32) return;
  cudaFuncSetAttribute(probe<N_ACC, ELECT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int grid = 148 * ctas;   // 100 KB dynamic smem: at most 2 CTAs per SM
  probe<N_ACC, ELECT, MODE><<<grid, 128, 100 * 1024>>>(N, reps, cols, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); exit(1); }
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = out[i] > mx ? out[i] : mx;
  const char* modes[4] = {"SS", "TS", "TS-MN", "attn-tile"};
  printf("%5d %5d %5d %6s %9s | %10.1f %10.1f\n", N, N_ACC, ctas, ELECT ? "elect" : "lane0", modes[MODE],
         double(mx) / (reps * (MODE == 3 ? 1 : 4)), 128.0 * N / 256.0);
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 1024 * sizeof(long long));
  printf("%5s %5s %5s %6s | %10s %10s\n", "N", "nacc", "ctas", "issue", "cyc/MMA", "work(cyc)");
  for (int ctas = 1; ctas <= 2; ++ctas)
    for (int N : {16, 48, 64, 128, 160, 256}) {
      run<1, false>(N, ctas, out);
      run<1, true>(N, ctas, out);
      run<2, true>(N, ctas, out);
      run<4, true>(N, ctas, out);
    }
  printf("-- A operand from tensor memory (cyc/MMA), and the attention per-key-tile MMA sequence (cyc/tile)\n");
  for (int ctas = 1; ctas <= 2; ++ctas) {
    for (int N : {16, 48, 64, 80, 128}) {
      run<1, true, 1>(N, ctas, out);
      run<1, true, 2>(N, ctas, out);
    }
    run<1, true, 3>(48, ctas, out);
    run<1, true, 3>(64, ctas, out);
  }
  return 0;
}
