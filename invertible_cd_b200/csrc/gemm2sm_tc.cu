// 2-SM tcgen05 GEMM (tcgen05.mma.cta_group::2) for the large plain / GEGLU linears.
//
//   out[M, N'] = epilogue( A[M,K] . B[N,K]^T + bias )      fp16 operands, fp32 accumulation in TMEM, fp16 out
//
// One 256x256 output tile per CTA PAIR (thread-block cluster of 2, two SMs of one TPC): each CTA loads its 128 rows of
// A and HALF of the B tile (128 of the 256 N rows) per 64-wide K block, i.e. 32 KB per SM per K block for 512 cycles
// of MMA, where the single-CTA 128x256 tile of gemm_tc.cuh needs 48 KB (per-SM TMA ingest tops out at ~70 B/clk,
// tools/probes/tma_mc_probe.cu), and the 256-column accumulator is double-buffered (a single-CTA 256x256 tile fills
// the TMEM). Measured (tools/probes/gemm2sm_epi_probe.cu, profiles/r1_i_gemm2sm_epi_probe.txt): GEGLU projection
// 4096x10240x1280 76.6 us (1401 TFLOP/s) vs 93.2 us with the single-CTA kernel; 4096x3840x1280 35.4 vs 38.8 us.
// icd_gemm (gemm_tc.cu) routes a launch here only for the shapes where that measurement says it pays.
//
//   warp 0      TMA producer (both CTAs; the transaction bytes of both are credited to the LEADER's full barrier)
//   warp 1      MMA issuer (leader CTA only) and TMEM allocation (both CTAs, cta_group::2)
//   warps 2..9  epilogue: each CTA drains its own 128 accumulator rows: TMEM -> registers -> (+bias | GEGLU
//               h*gelu(g)) -> 64B-swizzled smem staging (double-buffered 128x64 units) -> TMA stores
// Every mbarrier wait is bounded and traps instead of hanging.
#include <cuda_fp16.h>

#include <cstdlib>
#include <string>

#include "host_util.h"
#include "icd_ptx.cuh"

namespace icd {

struct Gemm2smParams {
  int M, N;              // N counts packed B rows (before GEGLU halving)
  int num_kb;            // 64-wide K blocks
  const float* bias;     // [N] (packed like B rows when geglu) or null
};

namespace {
constexpr int BK2 = 64, STAGES2 = 6;
constexpr int A2_BYTES = 128 * BK2 * 2, B2_BYTES = 128 * BK2 * 2, STAGE2_BYTES = A2_BYTES + B2_BYTES;
constexpr int C2_BYTES = 2 * 16384;          // staging: 2 x [128 rows x 64 cols] fp16 as 32-column atoms, 64B swizzle
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address: the pair's leader
constexpr int SMEM2_BYTES = STAGES2 * STAGE2_BYTES + C2_BYTES + 1024 /*align*/ + 256 /*barriers*/;

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
    if (spin > (1u << 26)) asm volatile("trap;");   // seconds: a protocol error, fail loudly instead of hanging
  }
}
// executed by both CTAs of the pair; the transaction bytes are credited to the LEADER's barrier
__device__ __forceinline__ void tma_load_4d_2sm(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, "
      "%5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1), "r"(0)
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

template <bool GEGLU>
__global__ void __launch_bounds__(320, 1)
gemm2sm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmOut, const Gemm2smParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES2 * A2_BYTES;
  uint8_t* smem_c = smem + STAGES2 * STAGE2_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_c + C2_BYTES);
  uint64_t* full_bar = bars;                     // [STAGES2]  (the leader's is the one that counts)
  uint64_t* empty_bar = bars + STAGES2;          // [STAGES2]  own, released by the leader's multicast commit
  uint64_t* tmem_full = bars + 2 * STAGES2;      // [2] own, multicast commit
  uint64_t* tmem_empty = bars + 2 * STAGES2 + 2; // [2] leader's: 2 x 256 epilogue threads arrive
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES2 + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    for (int i = 0; i < STAGES2; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 512); }
    fence_mbar_init();
  }
  if (warp == 1) {   // the same warp id in both CTAs
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();   // everything above ran under the previous kernel's tail; global memory is touched below

  const int n_tiles = p.N / 256;
  const int total = (p.M / 256) * n_tiles;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < total; tile += n_clusters) {
        const int mt = tile / n_tiles, nt = tile - mt * n_tiles;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait_bounded(&empty_bar[stage], phase ^ 1);
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * STAGE2_BYTES);   // bytes of both CTAs
          tma_load_4d_2sm(smem_a + stage * A2_BYTES, &tmA, &full_bar[stage], kb * BK2, mt * 256 + rank * 128);
          tma_load_4d_2sm(smem_b + stage * B2_BYTES, &tmB, &full_bar[stage], kb * BK2, nt * 256 + rank * 128);
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
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
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait_bounded(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + stage * (A2_BYTES >> 4), b_lo = b_lo0 + stage * (B2_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_ss_2sm(d_tmem, (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + k * 2u),
                            (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + k * 2u), idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_2sm(&empty_bar[stage]);
          if (kb == p.num_kb - 1) umma_commit_2sm(&tmem_full[acc]);
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // staged epilogue (same scheme as gemm_tc.cuh): 64-column units through a double-buffered staging area, the two
    // warps of a TMEM lane quadrant split every unit (32 columns each)
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
            uint32_t o[8];
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
              if (p.bias != nullptr) {
                b0 = __ldg(reinterpret_cast<const float4*>(p.bias + nt * 256 + col_t) + j4);
                b1 = __ldg(reinterpret_cast<const float4*>(p.bias + nt * 256 + 128 + col_t) + j4);
              }
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
          const uint32_t atom = buf + part * 8192 + row * 64;
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
            if (p.bias != nullptr) {
              const float4* bp = reinterpret_cast<const float4*>(p.bias + nt * 256 + col_t);
              b0 = __ldg(bp + 2 * cc);
              b1 = __ldg(bp + 2 * cc + 1);
            }
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
            tma_store_4d(&tmOut, smem_c + (unit & 1) * 16384 + h32 * 8192, nt * outw + u * 64 + h32 * 32,
                         mt * 256 + static_cast<int>(rank) * 128, 0, 0);
          bulk_commit();
        }
      }
    }
    if (st_leader) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // neither CTA may free its TMEM / exit while the pair can still touch it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

template <bool GEGLU>
int launch_2sm(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& o, const Gemm2smParams& p, cudaStream_t st) {
  static PerDeviceFlag configured;
  if (!configured.cur()) {
    cudaError_t e = cudaFuncSetAttribute(gemm2sm_tc_kernel<GEGLU>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES);
    if (e != cudaSuccess) return set_error(std::string("gemm2sm cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    configured.cur() = true;
  }
  const int tiles = (p.M / 256) * (p.N / 256);
  const int pairs = sm_count() / 2;
  const int clusters = tiles < pairs ? tiles : pairs;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(320);
  cfg.dynamicSmemBytes = SMEM2_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm2sm_tc_kernel<GEGLU>, a, b, o, p);
  if (e != cudaSuccess) return set_error(std::string("gemm2sm launch: ") + cudaGetErrorString(e));
  return check_launch("gemm2sm_tc");
}
}  // namespace

// Shapes for which the pair kernel measured faster than the single-CTA one (profiles/r1_i_gemm2sm_epi_probe.txt):
// GEGLU projections with K >= 1024, and plain fat-N projections (N >= 3840, K >= 1024, M >= 4096).
bool gemm2sm_wanted(int M, int N, int K, bool geglu) {
  static const bool enabled = [] { const char* e = getenv("ICD_GEMM_2SM"); return e == nullptr || atoi(e) != 0; }();
  if (!enabled || (M % 256) != 0 || (N % 256) != 0 || (K % 64) != 0 || sm_count() < 2) return false;
  if (geglu) return K >= 1024 && M >= 2048;
  return N >= 3840 && K >= 1024 && M >= 4096;
}

int launch_gemm2sm(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& o, int M, int N, int num_kb,
                   const float* bias, bool geglu, cudaStream_t st) {
  Gemm2smParams p{M, N, num_kb, bias};
  return geglu ? launch_2sm<true>(a, b, o, p, st) : launch_2sm<false>(a, b, o, p, st);
}

}  // namespace icd
