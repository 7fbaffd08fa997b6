// Attention with a short key/value sequence (N_kv <= 80: the 77-token text context of every cross-attention layer,
// utils/p2p.py:331-338 with is_cross == True), sm_100a.
//
// The flash-style kernel (attention_tc.cu) gives each 128-query tile its own CTA. With 77 keys a tile is ~0.1 us of
// tensor work behind ~5 us of serial start-up (barrier init, TMEM allocation, K/V and Q fetch, Q -> TMEM), so the
// launch ran at 50-110 TFLOP/s with the tensor pipe 8 % busy (profiles/r2_attn_cross64_ncu_full.txt). Here a CTA
// owns one (batch, head) and a RUN of query tiles:
//   * K and V are fetched once per CTA and stay in shared memory;
//   * Q tiles stream through a TMA ring, fetched ahead of use;
//   * the whole score row (<= 80 keys) fits in one accumulator, so the softmax is exact in one pass (no online
//     rescaling): the probabilities are normalised in registers, rounded to fp16 once, and those same values feed
//     P.V (TS-mode MMA, A = P in tensor memory) and the optional AttentionStore capture (utils/p2p.py:145-149);
//   * O overlays the dead tail of the score accumulator (columns 48..): 128 TMEM columns per CTA for d <= 80.
//   warp 0  TMA producer | warp 1  MMA issuer (one elected thread) | warps 2..5  softmax + epilogue (thread == row)
#include <cuda_fp16.h>

#include <cstdlib>
#include <string>

#include "../../include/icd_b200.h"
#include "host_util.h"
#include "icd_ptx.cuh"

namespace icd {

int make_tmap_4d(CUtensorMap* out, const void* ptr, const uint64_t dims[4], const uint64_t strides_b[3],
                 const uint32_t box[4], int swizzle_bytes, int elem_bytes);

struct SmallKvParams {
  int B, H, Nq, Nk;
  int tiles_per_cta;  // consecutive 128-query tiles per CTA
  int q_groups;       // CTAs per (batch, head)
  float scale_log2e;
  __half* out;
  long long out_ld;
  __half* probs;      // optional [B*H][Nq][probs_ld], probs_ld % 8 == 0, 16-byte aligned
  long long probs_ld;
  float* stats;       // optional [B*H][Nq][2]
};

template <int D>
struct SmallKvCfg {
  static constexpr int KW = 80;                        // key slots (MMA N of Q.K^T, MMA K of P.V)
  static constexpr int DP = (D + 15) / 16 * 16;
  static constexpr int DATOMS = (D + 63) / 64;
  static constexpr int OFF_O = 48;                     // P (packed half2) occupies columns [0, 40); O follows
  static constexpr int TMEM_NEED = (OFF_O + DP > KW) ? OFF_O + DP : KW;
  static constexpr int TMEM_COLS = TMEM_NEED <= 128 ? 128 : 256;
  static constexpr int Q_TILE_BYTES = DATOMS * 16384;  // 128 rows x 128 B per 64-wide atom
  static constexpr int KV_BYTES = DATOMS * KW * 128;   // 80 rows x 128 B per atom
  static constexpr int Q_STAGES = DATOMS == 1 ? 3 : 2;
  // + 256: barriers and the TMEM pointer; + 1024: slack for the 1024-byte alignment of the dynamic segment when
  // something (a tool, a future static array) puts static shared memory in front of it
  static constexpr int SMEM_BYTES = Q_STAGES * Q_TILE_BYTES + 2 * KV_BYTES + 256 + 1024;
  static constexpr int THREADS = 192;
  static constexpr int MIN_CTAS = (3 * (SMEM_BYTES + 1024) <= 228 * 1024 && 3 * TMEM_COLS <= 512)   ? 3
                                  : (2 * (SMEM_BYTES + 1024) <= 228 * 1024 && 2 * TMEM_COLS <= 512) ? 2
                                                                                                     : 1;
};

// NCTA = resident CTAs per SM the register allocation is bounded for. 2: 134 registers, no spills. 3 (d <= 64 only,
// ICD_ATTN_SMALLKV_CTAS=3): 96 registers and ~19 spilled words per thread; measured slower on every shape
// (4096x77 d=40: 28.5 vs 21.6 us, 1024x77 d=64: 10.6 vs 9.2 us), kept for re-measurement.
template <int D, int NCTA>
__global__ void __launch_bounds__(SmallKvCfg<D>::THREADS, NCTA)
attention_smallkv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                         const __grid_constant__ CUtensorMap tmV, const SmallKvParams p) {
  using Cfg = SmallKvCfg<D>;
  constexpr int KW = Cfg::KW, DP = Cfg::DP, DATOMS = Cfg::DATOMS, QS = Cfg::Q_STAGES, OFF_O = Cfg::OFF_O;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + QS * Cfg::Q_TILE_BYTES;
  uint8_t* sV = sK + Cfg::KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + Cfg::KV_BYTES);
  uint64_t* kv_full = bars;             // 1
  uint64_t* q_full = bars + 1;          // QS
  uint64_t* q_empty = q_full + QS;      // QS   (Q.K^T of the tile completed: the stage may be refilled)
  uint64_t* s_full = q_empty + QS;      // 1    (scores landed in TMEM)
  uint64_t* p_full = s_full + 1;        // 1    (128 softmax threads wrote P over the head of S)
  uint64_t* pv_done = p_full + 1;       // 1    (O complete)
  uint64_t* o_free = pv_done + 1;       // 1    (128 threads read O: the accumulator may be overwritten)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_free + 1);
  static_assert((5 + 2 * QS) * 8 + 4 <= 256, "barrier block");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_tiles = (p.Nq + 127) / 128;
  const int grp = blockIdx.x % p.q_groups;
  const int bh = blockIdx.x / p.q_groups;
  const int h = bh % p.H, b = bh / p.H;
  const int tile0 = grp * p.tiles_per_cta;
  const int n_tiles = min(p.tiles_per_cta, q_tiles - tile0);

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(kv_full, 1);
    for (int i = 0; i < QS; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(pv_done, 1);
    mbar_init(o_free, 128);
    fence_mbar_init();
  } else if (warp == 1) {
    tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr_smem);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();

  if (warp == 0) {
    if (elect_one() && n_tiles > 0) {
      mbar_expect_tx(kv_full, 2 * Cfg::KV_BYTES);
#pragma unroll
      for (int a = 0; a < DATOMS; ++a) {   // key rows beyond N_kv and head-dim columns beyond D are zero-filled
        tma_load_4d(sK + a * KW * 128, &tmK, kv_full, a * 64, h, 0, b);
        tma_load_4d(sV + a * KW * 128, &tmV, kv_full, a * 64, h, 0, b);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < n_tiles; ++i) {
        mbar_wait(&q_empty[stage], phase ^ 1);
        mbar_expect_tx(&q_full[stage], Cfg::Q_TILE_BYTES);
#pragma unroll
        for (int a = 0; a < DATOMS; ++a)
          tma_load_4d(sQ + stage * Cfg::Q_TILE_BYTES + a * 16384, &tmQ, &q_full[stage], a * 64, h,
                      (tile0 + i) * 128, b);
        if (++stage == QS) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one() && n_tiles > 0) {
      constexpr uint32_t idesc_s = umma_idesc_f16(128, KW, false, false);
      constexpr uint32_t idesc_o = umma_idesc_f16(128, DP, false, true);
      const uint64_t desc_hi = static_cast<uint64_t>((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;
      const uint32_t q_lo0 = ((smem_u32(sQ) >> 4) & 0x3FFFu) | (1u << 16);
      const uint32_t k_lo0 = ((smem_u32(sK) >> 4) & 0x3FFFu) | (1u << 16);
      // V is consumed MN-major (head dim contiguous): LBO = distance between 64-wide head-dim atoms
      const uint32_t v_lo0 = ((smem_u32(sV) >> 4) & 0x3FFFu) | ((static_cast<uint32_t>(KW * 128) >> 4) << 16);
      mbar_wait(kv_full, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < n_tiles; ++i) {
        mbar_wait(&q_full[stage], phase);
        if (i > 0) mbar_wait(o_free, (i - 1) & 1);
        tc_fence_after();
        const uint32_t q_lo = q_lo0 + stage * (Cfg::Q_TILE_BYTES >> 4);
#pragma unroll
        for (int a = 0; a < DATOMS; ++a) {
          const int kk_n = (DP - a * 64) >= 64 ? 4 : (DP - a * 64) / 16;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            if (kk < kk_n)
              umma_f16_ss(tmem_base, desc_hi | (q_lo + a * (16384u >> 4) + kk * 2u),
                          desc_hi | (k_lo0 + a * (static_cast<uint32_t>(KW * 128) >> 4) + kk * 2u), idesc_s,
                          (a | kk) != 0);
          }
        }
        umma_commit(&q_empty[stage]);
        umma_commit(s_full);
        mbar_wait(p_full, i & 1);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < KW / 16; ++ks)
          umma_f16_ts(tmem_base + OFF_O, tmem_base + ks * 8, desc_hi | (v_lo0 + ks * (2048u >> 4)), idesc_o,
                      ks != 0);
        umma_commit(pv_done);
        if (++stage == QS) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t t0 = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const float sl2e = p.scale_log2e;
    const int nchunk = p.probs != nullptr ? static_cast<int>(p.probs_ld >> 3) : 0;
    for (int i = 0; i < n_tiles; ++i) {
      const int q = (tile0 + i) * 128 + row;
      mbar_wait(s_full, i & 1);
      tc_fence_after();
      float s[KW];
#pragma unroll
      for (int c = 0; c < KW; c += 16) tmem_ld16(t0 + c, s + c);
      tmem_ld_wait();
#pragma unroll
      for (int c = 64; c < KW; ++c)            // only the last 16 slots can be padding when N_kv > 64 ...
        if (c >= p.Nk) s[c] = -INFINITY;
      if (p.Nk < 64) {                         // ... shorter contexts: warp-uniform slow path
#pragma unroll
        for (int c = 0; c < 64; ++c)
          if (c >= p.Nk) s[c] = -INFINITY;
      }
      float mx[4] = {fmaxf(s[0], s[1]), fmaxf(s[2], s[3]), fmaxf(s[4], s[5]), fmaxf(s[6], s[7])};
#pragma unroll
      for (int c = 8; c < KW; c += 8) {
#pragma unroll
        for (int u = 0; u < 4; ++u) mx[u] = fmaxf(mx[u], fmaxf(s[c + 2 * u], s[c + 2 * u + 1]));
      }
      const float m = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
      const float m_scaled = m * sl2e;
      float sum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < KW; ++c) {
        float e;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(s[c], sl2e, -m_scaled)));
        s[c] = e;
        sum[c & 3] += e;
      }
      const float inv = 1.0f / ((sum[0] + sum[1]) + (sum[2] + sum[3]));
      uint32_t pk[KW / 2];
#pragma unroll
      for (int c = 0; c < KW / 2; ++c) {
        const __half2 e = __floats2half2_rn(s[2 * c] * inv, s[2 * c + 1] * inv);
        pk[c] = *reinterpret_cast<const uint32_t*>(&e);
      }
      tmem_st16_u32(t0, pk);
      tmem_st16_u32(t0 + 16, pk + 16);
      tmem_st8_u32(t0 + 32, pk + 32);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full);
      if (q < p.Nq) {
        if (p.stats != nullptr)
          *reinterpret_cast<float2*>(p.stats + (static_cast<long long>(bh) * p.Nq + q) * 2) = make_float2(m_scaled, inv);
        if (nchunk > 0) {
          // normalised probabilities, the whole padded row [0, probs_ld): padding slots hold exp2(-inf) = 0
          uint4* dst = reinterpret_cast<uint4*>(p.probs + (static_cast<long long>(bh) * p.Nq + q) * p.probs_ld);
#pragma unroll
          for (int c = 0; c < KW / 8; ++c)
            if (c < nchunk) dst[c] = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
          for (int c = KW / 8; c < nchunk; ++c) dst[c] = make_uint4(0, 0, 0, 0);
        }
      }
      // epilogue of this tile: O -> fp16 -> global (already normalised)
      mbar_wait(pv_done, i & 1);
      tc_fence_after();
      __half* orow = p.out + (static_cast<long long>(b) * p.Nq + q) * p.out_ld + h * D;
#pragma unroll
      for (int c0 = 0; c0 < DP; c0 += 16) {
        float o[16];
        tmem_ld16(t0 + OFF_O + c0, o);
        tmem_ld_wait();
        if (c0 + 16 >= DP) {                  // last read of the accumulator: release it to the next tile's Q.K^T
          tc_fence_before();
          mbar_arrive(o_free);
        }
        if (q < p.Nq) {
          __half hv[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) hv[u] = __float2half_rn(o[u]);
          if (c0 + 16 <= D) {
            reinterpret_cast<uint4*>(orow + c0)[0] = reinterpret_cast<const uint4*>(hv)[0];
            reinterpret_cast<uint4*>(orow + c0)[1] = reinterpret_cast<const uint4*>(hv)[1];
          } else if (c0 + 8 <= D) {
            reinterpret_cast<uint4*>(orow + c0)[0] = reinterpret_cast<const uint4*>(hv)[0];
            for (int u = 8; u < 16; ++u)
              if (c0 + u < D) orow[c0 + u] = hv[u];
          } else {
            for (int u = 0; u < 16; ++u)
              if (c0 + u < D) orow[c0 + u] = hv[u];
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int D, int NCTA>
static int launch_smallkv(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, SmallKvParams p,
                          int tpc_override, cudaStream_t st) {
  using Cfg = SmallKvCfg<D>;
  static PerDeviceFlag configured;
  if (!configured.cur()) {
    cudaError_t e = cudaFuncSetAttribute(attention_smallkv_kernel<D, NCTA>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess)
      return set_error(std::string("attention_smallkv cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    configured.cur() = true;
  }
  // One wave: every CTA resident at once, each walking a run of query tiles of its (batch, head): the shortest run
  // length whose grid B*H*ceil(q_tiles / run) still fits the resident slots (measured: a second wave costs more than
  // longer runs, tools/attn_bench.py --cross-only).
  const int q_tiles = (p.Nq + 127) / 128;
  const int ctas_per_sm = NCTA;
  const long long slots = static_cast<long long>(sm_count()) * ctas_per_sm;
  const long long bh = static_cast<long long>(p.B) * p.H;
  int tpc = 1;
  while (tpc < q_tiles && bh * ((q_tiles + tpc - 1) / tpc) > slots) ++tpc;
  if (tpc_override > 0) tpc = tpc_override < q_tiles ? tpc_override : q_tiles;
  p.tiles_per_cta = tpc;
  p.q_groups = (q_tiles + tpc - 1) / tpc;
  const int grid = p.B * p.H * p.q_groups;
  launch_k(attention_smallkv_kernel<D, NCTA>, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, st, tq, tk, tv, p);
  return check_launch("attention_smallkv");
}

// Called by icd_attention_ex (attention_tc.cu) for N_kv <= 80. Returns -1 when the problem is not eligible
// (the caller falls through to the flash-style kernel), 0 on success, 1 on error.
int attention_smallkv_dispatch(const void* q, const void* k, const void* v, void* out, int B, int H, int Nq, int Nk,
                               int D, long long q_ld, long long k_ld, long long v_ld, long long out_ld, float scale,
                               void* probs_out, long long probs_ld, float* stats_out, cudaStream_t st) {
  static const int enabled = [] { const char* e = getenv("ICD_ATTN_SMALLKV"); return e ? atoi(e) : 1; }();
  static const int tpc_env = [] { const char* e = getenv("ICD_ATTN_SMALLKV_TPC"); return e ? atoi(e) : 0; }();
  if (!enabled || Nk > 80) return -1;
  if (D != 40 && D != 64 && D != 80 && D != 160) return -1;
  if (probs_out != nullptr && ((probs_ld & 7) != 0 || (reinterpret_cast<uintptr_t>(probs_out) & 15) != 0)) return -1;
  CUtensorMap tq, tk, tv;
  const uint32_t box[4] = {64, 1, 128, 1};
  const uint32_t boxkv[4] = {64, 1, 80, 1};
  {
    const uint64_t dims[4] = {(uint64_t)D, (uint64_t)H, (uint64_t)Nq, (uint64_t)B};
    const uint64_t str[3] = {(uint64_t)D * 2, (uint64_t)q_ld * 2, (uint64_t)q_ld * Nq * 2};
    if (make_tmap_4d(&tq, q, dims, str, box, 128, 2)) return 1;
  }
  {
    const uint64_t dims[4] = {(uint64_t)D, (uint64_t)H, (uint64_t)Nk, (uint64_t)B};
    const uint64_t strk[3] = {(uint64_t)D * 2, (uint64_t)k_ld * 2, (uint64_t)k_ld * Nk * 2};
    if (make_tmap_4d(&tk, k, dims, strk, boxkv, 128, 2)) return 1;
    const uint64_t strv[3] = {(uint64_t)D * 2, (uint64_t)v_ld * 2, (uint64_t)v_ld * Nk * 2};
    if (make_tmap_4d(&tv, v, dims, strv, boxkv, 128, 2)) return 1;
  }
  SmallKvParams p;
  p.B = B; p.H = H; p.Nq = Nq; p.Nk = Nk;
  p.tiles_per_cta = 1; p.q_groups = 1;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__half*>(out);
  p.out_ld = out_ld;
  p.probs = reinterpret_cast<__half*>(probs_out);
  p.probs_ld = probs_ld;
  p.stats = stats_out;
  static const int ncta_env = [] { const char* e = getenv("ICD_ATTN_SMALLKV_CTAS"); return e ? atoi(e) : 0; }();
  static_assert(SmallKvCfg<40>::MIN_CTAS == 3 && SmallKvCfg<64>::MIN_CTAS == 3 && SmallKvCfg<80>::MIN_CTAS == 2 &&
                    SmallKvCfg<160>::MIN_CTAS == 1, "resident CTAs per SM by shared / tensor memory");
  const bool three = ncta_env >= 3;
  switch (D) {
    case 40: return three ? launch_smallkv<40, 3>(tq, tk, tv, p, tpc_env, st) : launch_smallkv<40, 2>(tq, tk, tv, p, tpc_env, st);
    case 64: return three ? launch_smallkv<64, 3>(tq, tk, tv, p, tpc_env, st) : launch_smallkv<64, 2>(tq, tk, tv, p, tpc_env, st);
    case 80: return launch_smallkv<80, 2>(tq, tk, tv, p, tpc_env, st);
    default: return launch_smallkv<160, 1>(tq, tk, tv, p, tpc_env, st);
  }
}

}  // namespace icd
