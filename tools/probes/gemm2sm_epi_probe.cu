// Probe for the next round: 2-SM tcgen05 GEMM (tcgen05.mma.cta_group::2) — D[M,N] = A[M,K] . B[N,K]^T, fp16 in, fp32
// accumulation in TMEM, fp16 out. One 256x256 output tile per CTA PAIR: each CTA loads its 128 rows of A and HALF of
// the B tile (128 of the 256 N rows) per 64-wide K block, i.e. 32 KB per SM per K block for 512 cycles of MMA, where a
// single-CTA 128x256 tile needs 48 KB (per-SM TMA ingest tops out at ~70 B/clk: tools/probes/tma_mc_probe.cu). The
// 256x256 accumulator is 256 TMEM columns per CTA, so it is double-buffered (a single-CTA 256x256 tile is not).
//   warp 0: TMA producer (both CTAs; transaction bytes of both land on the LEADER's full barrier)
//   warp 1: MMA issuer (leader CTA only) + TMEM allocation (both CTAs, cta_group::2)
//   warps 2..9: epilogue (each CTA drains its own 128 accumulator rows through smem staging + TMA stores)
// Every barrier wait is bounded and traps instead of hanging.
// This variant (gemm2sm_epi_probe) replaces the naive epilogue by the library's staged one: 8 epilogue warps, TMEM ->
// registers -> (+bias | GEGLU h*gelu(g)) -> 64B-swizzled smem staging (double-buffered 128x64 units) -> TMA stores.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I invertible_cd_b200/csrc tools/probes/gemm2sm_epi_probe.cu -o tools/probes/gemm2sm_epi_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "icd_ptx.cuh"
using namespace icd;

constexpr int BK = 64, STAGES = 6;
constexpr int A_BYTES = 128 * BK * 2, B_BYTES = 128 * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address: the pair's leader

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0;; ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
    if (spin > (1u << 22)) asm volatile("trap;");
  }
}
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  // executed by both CTAs of the pair; the transaction bytes are credited to the LEADER's barrier
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16_ss_2sm(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(
          tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {   // arrives on `bar` of BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & PEER_MASK) : "memory");
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}

constexpr int C_BYTES = 2 * 16384;   // staging: 2 x [128 rows x 64 cols] fp16 as 32-column atoms, 64B swizzle

template <bool GEGLU>
__global__ void __launch_bounds__(320, 1)
gemm2sm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmOut, const float* __restrict__ bias, int M, int N, int K) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_BYTES;
  uint8_t* smem_c = smem + STAGES * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_c + C_BYTES);
  uint64_t* full_bar = bars;                   // [STAGES]  (leader's is the one that counts)
  uint64_t* empty_bar = bars + STAGES;         // [STAGES]  own, released by the leader's multicast commit
  uint64_t* tmem_full = bars + 2 * STAGES;     // [2] own, multicast commit
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;// [2] leader's: 2 x 256 epilogue threads arrive
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 512); }
    fence_mbar_init();
  }
  if (warp == 1) {   // same warp id in both CTAs
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int m_tiles = M / 256, n_tiles = N / 256, num_kb = K / BK;
  const int total = m_tiles * n_tiles;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < total; tile += n_clusters) {
        const int mt = tile / n_tiles, nt = tile - mt * n_tiles;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait_bounded(&empty_bar[stage], phase ^ 1);
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);   // bytes of both CTAs
          tma_load_2d_2sm(smem_a + stage * A_BYTES, &tmA, &full_bar[stage], kb * BK, mt * 256 + rank * 128);
          tma_load_2d_2sm(smem_b + stage * B_BYTES, &tmB, &full_bar[stage], kb * BK, nt * 256 + rank * 128);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      const uint32_t idesc = umma_idesc_f16(256, 256, false, false);
      const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t a_lo0 = ((smem_u32(smem_a) >> 4) & 0x3FFFu) | (1u << 16);
      const uint32_t b_lo0 = ((smem_u32(smem_b) >> 4) & 0x3FFFu) | (1u << 16);
      int stage = 0, iter = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < total; tile += n_clusters, ++iter) {
        const int acc = iter & 1;
        mbar_wait_bounded(&tmem_empty[acc], ((iter >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait_bounded(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + stage * (A_BYTES >> 4), b_lo = b_lo0 + stage * (B_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ss_2sm(d_tmem, (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + k * 2u),
                            (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + k * 2u), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_2sm(&empty_bar[stage]);
          if (kb == num_kb - 1) umma_commit_2sm(&tmem_full[acc]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // staged epilogue (same scheme as invertible_cd_b200/csrc/gemm_tc.cuh): 64-column units through a double-buffered
    // staging area, the two warps of a TMEM lane quadrant split every unit (32 columns each)
    const int quad = warp & 3;
    const int part = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const bool st_leader = (warp == 2) && elect_one();
    const uint32_t stg = smem_u32(smem_c);
    const uint32_t sw = (row >> 1) & 3;           // 64B swizzle: 16B-chunk index ^= (row / 2) % 4
    constexpr int outw = GEGLU ? 128 : 256;       // output columns per tile
    constexpr int units = outw / 64;
    uint32_t unit = 0;
    int iter = 0;
    for (int tile = cluster_id; tile < total; tile += n_clusters, ++iter) {
      const int mt = tile / n_tiles, nt = tile - mt * n_tiles;
      const int acc = iter & 1;
      mbar_wait_bounded(&tmem_full[acc], (iter >> 1) & 1);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + acc * 256 + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
      for (int u = 0; u < units; ++u, ++unit) {
        const uint32_t buf = stg + (unit & 1) * 16384;
        if (st_leader) bulk_wait_read1();          // the store that last read this staging buffer has drained it
        named_bar_sync(1, 256);
        if constexpr (GEGLU) {
#pragma unroll 1
          for (int c16 = part * 32; c16 < part * 32 + 32; c16 += 16) {
            const int col_t = u * 64 + c16;
            float hv[16], gv[16];
            tmem_ld16(t_addr + col_t, hv);
            tmem_ld16(t_addr + 128 + col_t, gv);
            tmem_ld_wait();
            const float4* bh = reinterpret_cast<const float4*>(bias + nt * 256 + col_t);
            const float4* bg = reinterpret_cast<const float4*>(bias + nt * 256 + 128 + col_t);
            uint32_t o[8];
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              const float4 b0 = __ldg(bh + j4), b1 = __ldg(bg + j4);
              const float r0 = (hv[j4 * 4 + 0] + b0.x) * gelu_erf(gv[j4 * 4 + 0] + b1.x);
              const float r1 = (hv[j4 * 4 + 1] + b0.y) * gelu_erf(gv[j4 * 4 + 1] + b1.y);
              const float r2 = (hv[j4 * 4 + 2] + b0.z) * gelu_erf(gv[j4 * 4 + 2] + b1.z);
              const float r3 = (hv[j4 * 4 + 3] + b0.w) * gelu_erf(gv[j4 * 4 + 3] + b1.w);
              const __half2 h01 = __floats2half2_rn(r0, r1), h23 = __floats2half2_rn(r2, r3);
              o[j4 * 2] = *reinterpret_cast<const uint32_t*>(&h01);
              o[j4 * 2 + 1] = *reinterpret_cast<const uint32_t*>(&h23);
            }
            const uint32_t atom = buf + (c16 >> 5) * 8192 + row * 64;
            const uint32_t ch = (c16 & 16) >> 3;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(atom + ((ch ^ sw) << 4)), "r"(o[0]), "r"(o[1]),
                         "r"(o[2]), "r"(o[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(atom + (((ch + 1) ^ sw) << 4)), "r"(o[4]),
                         "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
          }
        } else {
          const int col_t = u * 64 + part * 32;
          float v[32];
          tmem_ld32(t_addr + col_t, v);
          tmem_ld_wait();
          const float4* bp = reinterpret_cast<const float4*>(bias + nt * 256 + col_t);
          const uint32_t atom = buf + part * 8192 + row * 64;
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const float4 b0 = __ldg(bp + 2 * cc), b1 = __ldg(bp + 2 * cc + 1);
            const __half2 h0 = __floats2half2_rn(v[cc * 8 + 0] + b0.x, v[cc * 8 + 1] + b0.y);
            const __half2 h1 = __floats2half2_rn(v[cc * 8 + 2] + b0.z, v[cc * 8 + 3] + b0.w);
            const __half2 h2 = __floats2half2_rn(v[cc * 8 + 4] + b1.x, v[cc * 8 + 5] + b1.y);
            const __half2 h3 = __floats2half2_rn(v[cc * 8 + 6] + b1.z, v[cc * 8 + 7] + b1.w);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(atom + ((cc ^ sw) << 4)),
                         "r"(*reinterpret_cast<const uint32_t*>(&h0)), "r"(*reinterpret_cast<const uint32_t*>(&h1)),
                         "r"(*reinterpret_cast<const uint32_t*>(&h2)), "r"(*reinterpret_cast<const uint32_t*>(&h3))
                         : "memory");
          }
        }
        if (u == units - 1) {                       // all TMEM reads of this accumulator are done
          tc_fence_before();
          mbar_arrive_leader(&tmem_empty[acc]);
        }
        fence_proxy_async_smem();
        named_bar_sync(2, 256);
        if (st_leader) {
#pragma unroll
          for (int h32 = 0; h32 < 2; ++h32)
            tma_store_2d(&tmOut, smem_c + (unit & 1) * 16384 + h32 * 8192, nt * outw + u * 64 + h32 * 32,
                         mt * 256 + static_cast<int>(rank) * 128);
          bulk_commit();
        }
      }
    }
    if (st_leader) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

__global__ void ref_kernel(const __half* A, const __half* B, float* C, int M, int N, int K) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
  if (n >= N) return;
  float s = 0.f;
  for (int k = 0; k < K; ++k) s += __half2float(A[(long long)m * K + k]) * __half2float(B[(long long)n * K + k]);
  C[(long long)m * N + n] = s;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void make_map(EncodeFn enc, CUtensorMap* m, void* p, int rows, int K) {
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {BK, 128}, es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("tensor map encode failed %d\n", (int)r); exit(1); }
}

static void make_out_map(EncodeFn enc, CUtensorMap* m, void* p, int rows, int cols) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {32, 128}, es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, p, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("out tensor map encode failed %d\n", (int)r); exit(1); }
}

static int run(EncodeFn enc, int M, int N, int K, bool check, bool geglu) {
  const int n_out = geglu ? N / 2 : N;
  __half *A, *B, *C;
  float* bias;
  cudaMalloc(&A, (size_t)M * K * 2);
  cudaMalloc(&B, (size_t)N * K * 2);
  cudaMalloc(&C, (size_t)M * n_out * 2);
  cudaMalloc(&bias, (size_t)N * 4);
  std::vector<__half> ha((size_t)M * K), hb((size_t)N * K);
  std::vector<float> hbias(N);
  unsigned s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 9) & 0xFFFF) / 65536.0f - 0.5f; };
  for (auto& x : ha) x = __float2half(rnd());
  for (auto& x : hb) x = __float2half(rnd() * 0.25f);
  for (auto& x : hbias) x = rnd();
  cudaMemcpy(A, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(B, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(bias, hbias.data(), hbias.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(C, 0, (size_t)M * n_out * 2);
  CUtensorMap ta, tb, to;
  make_map(enc, &ta, A, M, K);
  make_map(enc, &tb, B, N, K);
  make_out_map(enc, &to, C, M, n_out);
  const int smem = STAGES * STAGE_BYTES + C_BYTES + 1024 + 256;
  cudaFuncSetAttribute(gemm2sm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(gemm2sm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  const int tiles = (M / 256) * (N / 256);
  int clusters = tiles < 74 ? tiles : 74;
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(320);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  const int reps = check ? 1 : 10;   // back-to-back launches per measurement: amortises the launch latency
  for (int it = 0; it < (check ? 1 : 4); ++it) {
    cudaEventRecord(e0);
    cudaError_t le = cudaSuccess;
    for (int r = 0; r < reps && le == cudaSuccess; ++r)
      le = geglu ? cudaLaunchKernelEx(&cfg, gemm2sm_kernel<true>, ta, tb, to, (const float*)bias, M, N, K)
                 : cudaLaunchKernelEx(&cfg, gemm2sm_kernel<false>, ta, tb, to, (const float*)bias, M, N, K);
    cudaEventRecord(e1);
    cudaError_t err = cudaEventSynchronize(e1);
    if (le != cudaSuccess || err != cudaSuccess) {
      printf("M=%d N=%d K=%d: launch %s / sync %s\n", M, N, K, cudaGetErrorString(le), cudaGetErrorString(err));
      return 1;
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    if (it > 0 || check) best = ms < best ? ms : best;
  }
  if (check) {
    float* R;
    cudaMalloc(&R, (size_t)M * N * 4);
    ref_kernel<<<dim3((N + 127) / 128, M), 128>>>(A, B, R, M, N, K);
    std::vector<float> hr((size_t)M * N);
    std::vector<__half> hc((size_t)M * n_out);
    cudaMemcpy(hr.data(), R, hr.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hc.data(), C, hc.size() * 2, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0;
    long long bad = 0;
    for (int m = 0; m < M; ++m)
      for (int j = 0; j < n_out; ++j) {
        double ref;
        if (geglu) {   // B rows are packed per 256-wide tile as [128 hidden | 128 gate]
          const int t = j / 128, jj = j % 128;
          const double h = hr[(size_t)m * N + t * 256 + jj] + hbias[t * 256 + jj];
          const double g = hr[(size_t)m * N + t * 256 + 128 + jj] + hbias[t * 256 + 128 + jj];
          ref = h * 0.5 * g * (1.0 + erf(g * 0.70710678118654752440));
        } else {
          ref = hr[(size_t)m * N + j] + hbias[j];
        }
        const double e = fabs((double)__half2float(hc[(size_t)m * n_out + j]) - ref);
        if (e > maxerr) maxerr = e;
        if (fabs(ref) > maxref) maxref = fabs(ref);
        if (e > 2e-2 + 2e-3 * fabs(ref)) ++bad;
      }
    printf("check %s M=%d N=%d K=%d: max err %.4g (ref max %.3g), %lld out of tolerance  -> %s\n", geglu ? "geglu" : "bias ",
           M, N, K, maxerr, maxref, bad, bad == 0 ? "OK" : "MISMATCH");
    cudaFree(R);
  } else {
    printf("time  %s M=%6d N=%5d K=%5d: %8.1f us  %8.1f TFLOP/s\n", geglu ? "geglu" : "bias ", M, N, K, best * 1e3,
           2.0 * M * N * K / (best * 1e-3) / 1e12);
  }
  cudaFree(A); cudaFree(B); cudaFree(C); cudaFree(bias);
  return 0;
}

int main() {
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", reinterpret_cast<void**>(&enc), cudaEnableDefault, &q);
  if (!enc) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  if (run(enc, 512, 512, 256, true, false)) return 1;
  if (run(enc, 1024, 768, 1280, true, false)) return 1;
  if (run(enc, 512, 1024, 320, true, true)) return 1;
  run(enc, 4096, 3840, 1280, false, false);   // in-library 128x256: 38.8 us
  run(enc, 4096, 1280, 1280, false, false);   // in-library 256x160: 19.5 us
  run(enc, 4096, 4096, 4096, false, false);   // in-library 256x256: 107 us; naive-epilogue 2-SM probe: 98.6 us
  run(enc, 4096, 10240, 1280, false, true);   // SDXL GEGLU projection: in-library 93.2 us
  run(enc, 16384, 5120, 640, false, true);    // SDXL GEGLU projection, 64^2 level
  run(enc, 8192, 5120, 640, false, true);     // SD1.5 GEGLU projection: in-library 54-55 us
  run(enc, 32768, 2560, 320, false, true);    // SD1.5 GEGLU projection, 64^2 level: in-library 85.8 us
  return 0;
}
