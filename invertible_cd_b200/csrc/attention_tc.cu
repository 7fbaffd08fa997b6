// Fused attention core on tcgen05 tensor cores (flash-style online softmax), sm_100a.
//
//   O[b, q, h, :] = softmax_k( scale * Q[b, q, h, :] . K[b, k, h, :] ) . V[b, k, h, :]
//
// Replaces F.scaled_dot_product_attention (un-patched diffusers Attention: forward_cons_model and every SDXL
// model) and the baddbmm -> softmax -> bmm sequence of the p2p-patched forward (utils/p2p.py:335-338) whenever the
// controller only *reads* the probabilities; for cross-attention (N_kv <= 128) the normalised probabilities are
// written out in the same pass (AttentionStore capture, utils/p2p.py:145-149) instead of being materialised by a
// separate GEMM + softmax.
//
// One CTA = one (batch, head, 128-query tile); K/V stream through a TMA ring in 64-key tiles.
//   warp 0      TMA producer
//   warp 1      MMA issuer:  S = Q.K^T  (M128 x N64 x D)   ->  TMEM S[j&1]       (A = Q from TENSOR MEMORY)
//                            O += P.V   (M128 x D x K64)    ->  TMEM O            (A = P from TENSOR MEMORY,
//                                                                                  V consumed MN-major from smem)
//                            L += P.1   (M128 x 16 x K64)   ->  TMEM L            (row sums on the tensor core)
//   warps 2..5  softmax: one thread per query row == one TMEM lane; S read once from TMEM, P = exp2(..) packed to
//               half2 and written back over the first 32 columns of the S buffer it came from; O rescaled in TMEM
//               only when a row maximum moved by more than 2^8.
// Both A operands live in tensor memory: an SS-mode M128 MMA re-reads its 4 KB A slice from shared memory every K16
// step, i.e. 32 cycles of the 128 B/clk port, so with N <= 64 it is smem-bound (measured, tools/probes/mma_probe.cu:
// 39/44/48 cycles for N = 16/48/64 where the tensor pipe needs 8/24/32).
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>
#include <string>

#include "../../include/icd_b200.h"
#include "host_util.h"
#include "icd_ptx.cuh"

namespace icd {

int make_tmap_4d(CUtensorMap* out, const void* ptr, const uint64_t dims[4], const uint64_t strides_b[3],
                 const uint32_t box[4], int swizzle_bytes, int elem_bytes);
// attention_smallkv.cu: N_kv <= 80 (text context). -1 = not eligible, 0 = launched, 1 = error.
int attention_smallkv_dispatch(const void* q, const void* k, const void* v, void* out, int B, int H, int Nq, int Nk,
                               int D, long long q_ld, long long k_ld, long long v_ld, long long out_ld, float scale,
                               void* probs_out, long long probs_ld, float* stats_out, cudaStream_t st);

struct AttnParams {
  int B, H, Nq, Nk;
  float scale_log2e;  // scale * log2(e)
  __half* out;
  long long out_ld;
  __half* probs;      // optional [B*H][Nq][probs_ld]
  long long probs_ld;
  float* stats;       // optional [B*H][Nq][2]: (m_run * scale_log2e, 1 / l) of the online softmax
  int speculate;      // exponentials of a tile start before its maximum is known (see the softmax loop)
#ifdef ICD_ATTN_PROFILE
  long long* prof;    // [MT][8] phase cycle counters of one softmax warp per query tile (debug builds only)
#endif
};

// Phase timing of one softmax warp (debug builds: make PROF=1): cycles spent per key tile in
//   0 wait S | 1 tcgen05.ld (first 32 scores) | 2 max + rescale decision | 3 second load / speculation redo |
//   4 exponentials (speculative tiles: incl. the wait for the second 32 scores) | 5 tcgen05.st + arrive
#ifdef ICD_ATTN_PROFILE
#define ICD_PROF_DECL long long pt_[7] = {0, 0, 0, 0, 0, 0, 0}, pc_ = clock64();
#define ICD_PROF_MARK(k) { const long long n_ = clock64(); pt_[k] += n_ - pc_; pc_ = n_; }
#else
#define ICD_PROF_DECL
#define ICD_PROF_MARK(k)
#endif

// 2^x on the FMA/ALU pipes (no MUFU): round-to-nearest split x = n + f, f in [-0.5, 0.5], degree-3 minimax polynomial
// for 2^f (max relative error 7.5e-5, below half an fp16 ulp), exponent added with an integer shift-add.
// The softmax is MUFU.EX2-bound (16 lanes/clk/SM); computing a fixed fraction of every row's exponentials this way
// moves work to pipes that are otherwise idle during the exponential phase.
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -120.0f);                        // keeps the biased exponent positive; 2^-120 rounds to 0 in fp16
  const float t = x + 12582912.0f;              // 1.5 * 2^23: round(x) lands in the low mantissa bits
  const float f = x - (t - 12582912.0f);
  const float pl = fmaf(fmaf(fmaf(0.0551716648f, f, 0.2426111251f), f, 0.6932609677f), f, 0.9999280572f);
  return __int_as_float(__float_as_int(pl) + (__float_as_int(t) << 23));
}
__device__ __forceinline__ float exp2_mufu(float x) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x));
  return e;
}

template <int D, int MT_>
struct AttnCfg {
  static constexpr int DP = (D + 15) / 16 * 16;      // MMA extent along the head dim
  static constexpr int DATOMS = (D + 63) / 64;       // 64-wide (128 B) swizzle atoms along the head dim
  static constexpr int BKV = 64;                     // keys per K/V tile
  // TMEM columns per 128-row query tile: S0/P0 (64) | S1/P1 (64) | O (DP) | L (16) | Q (32 per head-dim atom)
  static constexpr int TMEM_PER_TILE = 128 + DP + 16 + DATOMS * 32;
  // MT = query tiles per CTA. MT = 2 (needs d <= 64 for the 512 TMEM columns) fetches every K/V tile once per 256
  // query rows: it pays when the K/V stream is the limiter (d = 40: 80-byte head rows straddle 128-byte lines, the
  // TMA moves 160 B per row and the kernel sat at ~6.5 TB/s of L2->SM traffic whatever the softmax did). MT = 1
  // runs two independent CTAs per SM instead, which balances the tail wave better (d = 64).
  static constexpr int MT = MT_;
  static_assert(MT == 1 || 2 * TMEM_PER_TILE <= 512, "two query tiles need 2 x TMEM_PER_TILE <= 512 columns");
  static constexpr int TILE_COLS = 256;              // TMEM column stride between the two query tiles
  static constexpr int TMEM_COLS = (MT == 2 || TMEM_PER_TILE > 256) ? 512 : 256;
  static_assert(TMEM_PER_TILE <= (MT == 2 ? 256 : 512), "TMEM budget");
  static constexpr int KV_STAGES = DATOMS == 1 ? 6 : (DATOMS == 2 ? 4 : 3);
  static_assert(KV_STAGES >= 3, "Q.K^T runs two tiles ahead of P.V");
  static constexpr int Q_TILE_BYTES = DATOMS * 16384;   // 128 query rows
  static constexpr int Q_BYTES = MT * Q_TILE_BYTES;
  static constexpr int KV_BYTES = DATOMS * 8192;        // 64 key rows
  static constexpr int KV_RING_BYTES = 2 * KV_STAGES * KV_BYTES;
  static constexpr int SMEM_BYTES = Q_BYTES + KV_RING_BYTES + 512;   // + barriers, TMEM pointer, ones tile
  static_assert(KV_RING_BYTES >= MT * 2 * 16384, "K/V ring doubles as the probability staging buffer");
  static constexpr int THREADS = 32 * (1 + MT + 4 * MT);   // TMA warp; per query tile: 1 MMA warp + 4 softmax warps
  static constexpr int MIN_CTAS = (2 * SMEM_BYTES + 2048 <= 227 * 1024 && 2 * TMEM_COLS <= 512) ? 2 : 1;
};

// Pipeline (one CTA = one (batch, head) x MT 128-query tiles; K/V stream in 64-key tiles through a TMA ring):
//   per query tile, S is double-buffered in TMEM and P_j overwrites the head of S_j, so the tensor core computes
//   S_{j+1} = Q.K_{j+1}^T while the softmax warps are still working on tile j, and P_j.V_j runs while they start on
//   tile j+1. tcgen05.mma ops of one thread execute in issue order, which is what makes the aliasing safe:
//   Q.K_{j+2}^T (overwrites S[j&1] = P_j) is issued right after P_j.V_j.
//   warp 0: TMA producer | warp 1+m: MMA issuer of query tile m | warps 1+MT+4m .. 4+MT+4m: softmax of query tile m
//   (one thread per query row, S read ONCE from TMEM). The query tiles share nothing but the K/V ring: each has its
//   own issuer so that neither waits behind the other's softmax.
// POLY = how many of every 8 exponentials are evaluated by exp2_poly instead of MUFU.EX2
template <int D, int MT_, int POLY>
__global__ void __launch_bounds__(AttnCfg<D, MT_>::THREADS, AttnCfg<D, MT_>::MIN_CTAS)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using Cfg = AttnCfg<D, MT_>;
  constexpr int DP = Cfg::DP, DATOMS = Cfg::DATOMS, ST = Cfg::KV_STAGES, BKV = Cfg::BKV, MT = Cfg::MT;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::Q_BYTES;
  uint8_t* sV = sK + ST * Cfg::KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + ST * Cfg::KV_BYTES);
  uint64_t* q_full = bars;                 // 1
  uint64_t* k_full = bars + 1;             // ST
  uint64_t* v_full = k_full + ST;          // ST
  uint64_t* kv_empty = v_full + ST;        // ST      (P.V of every query tile consumed K and V: MT arrivals)
  uint64_t* s_full = kv_empty + ST;        // 2 x MT  (Q.K^T landed in S[m][b])
  uint64_t* p_full = s_full + 2 * MT;      // 2 x MT  (128 softmax threads replaced S[m][b] by P[m][b] in TMEM)
  uint64_t* pv_done = p_full + 2 * MT;     // 2 x MT  (P[m][b].V finished: O/L of tile m quiescent)
  uint64_t* q_ready = pv_done + 2 * MT;    // MT      (softmax threads moved their Q rows from smem to TMEM)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(q_ready + MT);
  uint8_t* s_ones = reinterpret_cast<uint8_t*>(bars) + 384;   // one 8x8 fp16 core matrix of 1.0 (128 B)
  static_assert((1 + 3 * ST + 7 * MT) * 8 + 4 <= 384, "barrier block overflows into the ones tile");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef ICD_ATTN_PROFILE
  const long long cta_t0 = clock64();   // CTA timeline (slots 16..): set-up | Q landed | first S | key loop | epilogue
#define ICD_CTA_STAMP(k) if (p.prof != nullptr && blockIdx.x == gridDim.x / 2 && warp == 1 + MT && lane == 0) p.prof[16 + (k)] = clock64() - cta_t0;
#else
#define ICD_CTA_STAMP(k)
#endif
  const int q_tiles = (p.Nq + 128 * MT - 1) / (128 * MT);
  const int qt = blockIdx.x % q_tiles;
  const int bh = blockIdx.x / q_tiles;
  const int h = bh % p.H, b = bh / p.H;
  const int q0 = qt * 128 * MT;
  const int n_kv = (p.Nk + BKV - 1) / BKV;

  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < ST; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&kv_empty[i], MT);
    }
    for (int i = 0; i < 2 * MT; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
      mbar_init(&pv_done[i], 1);
    }
    for (int i = 0; i < MT; ++i) mbar_init(&q_ready[i], 128);
    fence_mbar_init();
  } else if (warp == 1) {
    tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr_smem);
  } else if (warp == 1 + MT) {
    reinterpret_cast<uint32_t*>(s_ones)[lane] = 0x3C003C00u;   // 64 halves of 1.0
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  ICD_CTA_STAMP(0)
  pdl_wait();   // prologue above overlaps the previous kernel's tail; Q/K/V are read below
  // per query tile m (column offset m * TILE_COLS): S0 | S1 | O | L | Q
  constexpr uint32_t OFF_O = 128, OFF_L = 128 + DP, OFF_Q = 128 + DP + 16;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(q_full, Cfg::Q_BYTES);
#pragma unroll
      for (int m = 0; m < MT; ++m) {
#pragma unroll
        for (int a = 0; a < DATOMS; ++a)   // rows beyond N_q (possibly the whole second tile) are zero-filled
          tma_load_4d(sQ + m * Cfg::Q_TILE_BYTES + a * 16384, &tmQ, q_full, a * 64, h, q0 + m * 128, b);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&kv_empty[stage], phase ^ 1);
        mbar_expect_tx(&k_full[stage], Cfg::KV_BYTES);
#pragma unroll
        for (int a = 0; a < DATOMS; ++a)
          tma_load_4d(sK + stage * Cfg::KV_BYTES + a * 8192, &tmK, &k_full[stage], a * 64, h, j * BKV, b);
        mbar_expect_tx(&v_full[stage], Cfg::KV_BYTES);
#pragma unroll
        for (int a = 0; a < DATOMS; ++a)
          tma_load_4d(sV + stage * Cfg::KV_BYTES + a * 8192, &tmV, &v_full[stage], a * 64, h, j * BKV, b);
        if (++stage == ST) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp <= MT) {
    const int m = warp - 1;                 // query tile of this issuer
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_f16(128, BKV, false, false);
      constexpr uint32_t idesc_o = umma_idesc_f16(128, DP, false, true);
      constexpr uint32_t idesc_l = umma_idesc_f16(128, 16, false, false);
      // ones operand = a single no-swizzle 8x8 core matrix reused for every (row group, k chunk): LBO=SBO=0
      const uint64_t ones_desc =
          (static_cast<uint64_t>((smem_u32(s_ones) >> 4) & 0x3FFF)) | (static_cast<uint64_t>(1) << 46);
      // smem descriptors are formed by adding compile-time constants to two precomputed 32-bit low words
      const uint64_t desc_hi = static_cast<uint64_t>((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;
      const uint32_t k_lo0 = ((smem_u32(sK) >> 4) & 0x3FFFu) | (1u << 16);
      const uint32_t v_lo0 = ((smem_u32(sV) >> 4) & 0x3FFFu) | ((8192u >> 4) << 16);   // MN-major: LBO = 8 KB
      // The issue loop is the serial link between "softmax of tile j done" and "S of tile j+2 ready", so everything
      // in it is straight-line with compile-time stage / buffer indices (unrolled over U key tiles), and a key tile
      // is always processed with all four K16 steps: keys beyond N_kv have P == 0 and zero-filled V rows.
      const uint32_t t0 = tmem_base + m * Cfg::TILE_COLS;
      uint64_t* s_full_m = s_full + 2 * m;
      uint64_t* p_full_m = p_full + 2 * m;
      uint64_t* pv_done_m = pv_done + 2 * m;
      auto issue_qk = [&](int st, int bsel) {   // S[m][bsel] = Q_m . K^T for the K tile in ring stage st
        const uint32_t k_lo = k_lo0 + st * (Cfg::KV_BYTES >> 4);
#pragma unroll
        for (int a = 0; a < DATOMS; ++a) {
          const int kk_n = (DP - a * 64) >= 64 ? 4 : (DP - a * 64) / 16;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            if (kk < kk_n)
              umma_f16_ts(t0 + bsel * BKV, t0 + OFF_Q + a * 32 + kk * 8,
                          desc_hi | (k_lo + a * (8192u >> 4) + kk * 2u), idesc_s, (a | kk) != 0);
          }
        }
        umma_commit(&s_full_m[bsel]);
      };
      constexpr int U = (ST % 2 == 0) ? ST : 2 * ST;
      mbar_wait(&q_ready[m], 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_qk(0, 0);
      if (n_kv > 1) {
        mbar_wait(&k_full[1], 0);
        tc_fence_after();
        issue_qk(1, 1);
      }
      for (int jb = 0; jb < n_kv; jb += U) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int j = jb + u;
          if (j < n_kv) {
            const int st = u % ST, bsel = u & 1, st2 = (u + 2) % ST;
            const bool has2 = j + 2 < n_kv;
            mbar_wait(&v_full[st], (jb / ST + u / ST) & 1);
            // K_{j+2} sits in a stage that was released by P.V of tile j+2-ST <= j-1 (already issued)
            if (has2) mbar_wait(&k_full[st2], (jb / ST + (u + 2) / ST) & 1);
            const uint32_t v_lo = v_lo0 + st * (Cfg::KV_BYTES >> 4);
            mbar_wait(&p_full_m[bsel], ((jb >> 1) + (u >> 1)) & 1);
            tc_fence_after();
            // O += P_j . V_j and L += P_j . 1 (A = P from tensor memory: the head of S[m][bsel])
            const uint32_t tmem_P = t0 + bsel * BKV;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t acc = (ks != 0) ? 1u : static_cast<uint32_t>(j != 0);
              umma_f16_ts(t0 + OFF_O, tmem_P + ks * 8, desc_hi | (v_lo + ks * (2048u >> 4)), idesc_o, acc);
              // row sums ride on the tensor core: L[128x16] += P[128xK] . ones[Kx16] (every column = sum_k P)
              umma_f16_ts(t0 + OFF_L, tmem_P + ks * 8, ones_desc, idesc_l, acc);
            }
            umma_commit(&kv_empty[st]);
            umma_commit(&pv_done_m[bsel]);
            // S[m][bsel] = Q_m . K_{j+2}^T, queued right behind P_j . V_j (which reads P_j from the same columns)
            if (has2) issue_qk(st2, bsel);
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue
    const int m = (warp - 1 - MT) >> 2;     // query tile of this warp
    const int quad = warp & 3;              // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;       // query row within the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t t0 = tmem_base + m * Cfg::TILE_COLS + lane_off;
    uint64_t* s_full_m = s_full + 2 * m;
    uint64_t* p_full_m = p_full + 2 * m;
    uint64_t* pv_done_m = pv_done + 2 * m;
    const int q = q0 + m * 128 + row;
    float m_run = -INFINITY;
    float alpha_after0 = 1.0f;              // rescale applied after tile 0's P was written (probability capture)
    const int sw = row & 7;
    // move this thread's Q row from the 128B-swizzled smem tile into tensor memory (A operand of S = Q.K^T);
    // rows beyond N_q were zero-filled by TMA
    mbar_wait(q_full, 0);
    ICD_CTA_STAMP(1)
#pragma unroll
    for (int a = 0; a < DATOMS; ++a) {
      uint32_t qr[32];
      const uint32_t q_row = smem_u32(sQ) + m * Cfg::Q_TILE_BYTES + a * 16384 + row * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(qr[4 * c]), "=r"(qr[4 * c + 1]), "=r"(qr[4 * c + 2]), "=r"(qr[4 * c + 3])
                     : "r"(q_row + ((c ^ sw) << 4)));
      tmem_st16_u32(t0 + OFF_Q + a * 32, qr);
      tmem_st16_u32(t0 + OFF_Q + a * 32 + 16, qr + 16);
    }
    tmem_st_wait();
    tc_fence_before();
    mbar_arrive(&q_ready[m]);
    ICD_PROF_DECL
    for (int j = 0; j < n_kv; ++j) {
      const int bsel = j & 1;
      ICD_PROF_MARK(6)
      mbar_wait(&s_full_m[bsel], (j >> 1) & 1);
      tc_fence_after();
      if (j == 0) { ICD_CTA_STAMP(2) }
      ICD_PROF_MARK(0)
      const int valid = min(BKV, p.Nk - j * BKV);
      float s[64];
      auto use_poly = [](int idx) constexpr {
        const int r = idx & 7;
        return (POLY >= 1 && r == 6) || (POLY >= 2 && r == 3) || (POLY >= 3 && r == 1) || (POLY >= 4 && r == 4);
      };
      tmem_ld32(t0 + bsel * BKV, s);
      tmem_ld_wait();
      ICD_PROF_MARK(1)
      if (p.speculate && j > 0 && valid == BKV) {
        // Speculative tile (every full tile but the first): the lazily-updated reference maximum m_run almost never
        // moves, so the exponentials start from the first 32 scores right away with the CURRENT m_run — no wait for
        // the tile maximum, and the load of the other 32 scores is in flight meanwhile. The maximum is tracked on the
        // side; if any row of the warp did move (warp-uniform check), the tile is redone on the regular path below
        // from the scores still in registers. Identical arithmetic, hence identical results.
        tmem_ld32(t0 + bsel * BKV + 32, s + 32);
        const float m_sc = m_run * p.scale_log2e;
        uint32_t pk[32];
        float mxa = fmaxf(s[0], s[1]), mxb = fmaxf(s[2], s[3]);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float x0 = fmaf(s[2 * i], p.scale_log2e, -m_sc), x1 = fmaf(s[2 * i + 1], p.scale_log2e, -m_sc);
          const float e0 = use_poly(2 * i) ? exp2_poly(x0) : exp2_mufu(x0);
          const float e1 = use_poly(2 * i + 1) ? exp2_poly(x1) : exp2_mufu(x1);
          const __half2 e = __floats2half2_rn(e0, e1);
          pk[i] = *reinterpret_cast<const uint32_t*>(&e);
          if (i >= 2) {
            if (i & 1) mxb = fmaxf(mxb, fmaxf(s[2 * i], s[2 * i + 1]));
            else mxa = fmaxf(mxa, fmaxf(s[2 * i], s[2 * i + 1]));
          }
        }
        tmem_ld_wait();
#pragma unroll
        for (int i = 16; i < 32; ++i) {
          const float x0 = fmaf(s[2 * i], p.scale_log2e, -m_sc), x1 = fmaf(s[2 * i + 1], p.scale_log2e, -m_sc);
          const float e0 = use_poly(2 * i) ? exp2_poly(x0) : exp2_mufu(x0);
          const float e1 = use_poly(2 * i + 1) ? exp2_poly(x1) : exp2_mufu(x1);
          const __half2 e = __floats2half2_rn(e0, e1);
          pk[i] = *reinterpret_cast<const uint32_t*>(&e);
          if (i & 1) mxb = fmaxf(mxb, fmaxf(s[2 * i], s[2 * i + 1]));
          else mxa = fmaxf(mxa, fmaxf(s[2 * i], s[2 * i + 1]));
        }
        const bool moved_spec = (fmaxf(mxa, mxb) - m_run) * p.scale_log2e > 8.0f;
        if (!__any_sync(0xffffffffu, moved_spec)) {
          tmem_st16_u32(t0 + bsel * BKV, pk);
          tmem_st16_u32(t0 + bsel * BKV + 16, pk + 16);
          ICD_PROF_MARK(4)
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&p_full_m[bsel]);
          ICD_PROF_MARK(5)
          continue;
        }
      } else {
        tmem_ld32(t0 + bsel * BKV + 32, s + 32);
        tmem_ld_wait();
      }
      ICD_PROF_MARK(3)
      if (valid < BKV) {                    // warp-uniform: only the last K/V tile can be partial
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i >= valid) s[i] = -INFINITY;
      }
      float mx[4] = {fmaxf(s[0], s[1]), fmaxf(s[2], s[3]), fmaxf(s[4], s[5]), fmaxf(s[6], s[7])};
#pragma unroll
      for (int i = 8; i < 64; i += 8) {
#pragma unroll
        for (int c = 0; c < 4; ++c) mx[c] = fmaxf(mx[c], fmaxf(s[i + 2 * c], s[i + 2 * c + 1]));
      }
      const float m_tile = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
      // Lazy rescale: the reference maximum only moves when the tile maximum exceeds it by more than 2^8 (exp2
      // domain); until then P <= 256 (fine in fp16 with fp32 accumulation) and the TMEM round trip is skipped.
      // Common case (no row of the warp moves): one subtract, one compare, one vote.
      const bool moved = (m_tile - m_run) * p.scale_log2e > 8.0f;
      if (__any_sync(0xffffffffu, moved)) {
        float alpha = 1.0f;
        if (moved) {
          alpha = exp2f((m_run - m_tile) * p.scale_log2e);   // 0 on the first tile (m_run = -inf)
          m_run = m_tile;
        }
        if (j > 0) {
          // O and L must be quiescent: P_{j-1}.V_{j-1} finished (P_j.V_j cannot start before our p_full arrive)
          mbar_wait(&pv_done_m[(j - 1) & 1], ((j - 1) >> 1) & 1);
          tc_fence_after();
#pragma unroll 1
          for (int c0 = 0; c0 < DP + 16; c0 += 16) {
            float o[16];
            tmem_ld16(t0 + OFF_O + c0, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] *= alpha;
            tmem_st16_u32(t0 + OFF_O + c0, reinterpret_cast<const uint32_t*>(o));
          }
          tmem_st_wait();
          if (j == 1) alpha_after0 = alpha;
        }
      }
      const float m_scaled = m_run * p.scale_log2e;
      ICD_PROF_MARK(2)
      // p = exp2(s*scale*log2e - m), packed to half2 and written over the first 32 columns of S[bsel] (this thread's
      // lane only; all 64 scores are already in registers). POLY of every 8 exponentials take the FMA-pipe
      // polynomial, spread out so that both instruction streams interleave.
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float x0 = fmaf(s[2 * i], p.scale_log2e, -m_scaled), x1 = fmaf(s[2 * i + 1], p.scale_log2e, -m_scaled);
        const float e0 = use_poly(2 * i) ? exp2_poly(x0) : exp2_mufu(x0);
        const float e1 = use_poly(2 * i + 1) ? exp2_poly(x1) : exp2_mufu(x1);
        const __half2 e = __floats2half2_rn(e0, e1);
        pk[i] = *reinterpret_cast<const uint32_t*>(&e);
      }
      tmem_st16_u32(t0 + bsel * BKV, pk);
      tmem_st16_u32(t0 + bsel * BKV + 16, pk + 16);
      ICD_PROF_MARK(4)
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full_m[bsel]);
      ICD_PROF_MARK(5)
    }
#ifdef ICD_ATTN_PROFILE
    if (p.prof != nullptr && blockIdx.x == gridDim.x / 2 && quad == 0 && lane == 0)
      for (int k = 0; k < 7; ++k) p.prof[m * 8 + k] = pt_[k];
#endif
    ICD_CTA_STAMP(3)
    // epilogue: O / l -> fp16 -> global   (l = row sum accumulated by the ones-MMA)
    mbar_wait(&pv_done_m[(n_kv - 1) & 1], ((n_kv - 1) >> 1) & 1);
    tc_fence_after();
    float inv;
    {
      float l16[16];
      tmem_ld16(t0 + OFF_L, l16);
      tmem_ld_wait();
      inv = 1.f / l16[0];
    }
    if (p.stats != nullptr && q < p.Nq)
      *reinterpret_cast<float2*>(p.stats + (static_cast<long long>(bh) * p.Nq + q) * 2) =
          make_float2(m_run * p.scale_log2e, inv);
    if (p.probs != nullptr && n_kv <= 2) {
      // <= 128 keys: emit the normalised probabilities (AttentionStore capture). The un-normalised P tiles are still
      // in tensor memory (nothing overwrote S0/S1). They are staged through the idle K/V ring in a swizzled row
      // layout (32 KB per query tile) so that the write-out below is coalesced. The ring is idle for THIS query tile
      // only once every MMA of the CTA has completed, which the last pv_done of the LAST query tile implies.
      if (MT > 1) {
        mbar_wait(&pv_done[2 * (MT - 1) + ((n_kv - 1) & 1)], ((n_kv - 1) >> 1) & 1);
        tc_fence_after();
      }
      const float sc0 = (n_kv == 2) ? inv * alpha_after0 : inv;   // tile 0 was written before the last rescale
      const uint32_t stage_base = smem_u32(sK) + m * 32768;
      for (int t = 0; t < n_kv; ++t) {
        uint32_t pk[32];
        tmem_ld32(t0 + t * BKV, reinterpret_cast<float*>(pk));
        tmem_ld_wait();
        const uint32_t dst = stage_base + row * 128 + t * 16384;
#pragma unroll
        for (int cc = 0; cc < 8; ++cc)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + ((cc ^ sw) << 4)), "r"(pk[4 * cc]),
                       "r"(pk[4 * cc + 1]), "r"(pk[4 * cc + 2]), "r"(pk[4 * cc + 3])
                       : "memory");
      }
      // Each warp owns 32 consecutive rows whose P was staged by its own lanes, so a warp-level sync suffices.
      __syncwarp();
      if ((p.probs_ld & 7) == 0 && (reinterpret_cast<uintptr_t>(p.probs) & 15) == 0) {
        // coalesced: consecutive lanes write consecutive 16-byte chunks; the whole padded row [0, probs_ld) is
        // written (masked keys have P == 0, chunks beyond the last tile are zero-filled)
        const int nchunk = static_cast<int>(p.probs_ld >> 3);
        const int row0 = quad * 32;
        const int qrow0 = q0 + m * 128 + row0;
        uint4* dst = reinterpret_cast<uint4*>(p.probs + (static_cast<long long>(bh) * p.Nq + qrow0) * p.probs_ld);
        const int rows_valid = p.Nq - qrow0;
        for (int idx = lane; idx < 32 * nchunk; idx += 32) {
          const int r = idx / nchunk, ch = idx - r * nchunk;
          const float s0 = __shfl_sync(0xffffffffu, sc0, r), s1 = __shfl_sync(0xffffffffu, inv, r);
          const int t = ch >> 3, cc = ch & 7;
          uint4 val = make_uint4(0, 0, 0, 0);
          if (t < n_kv) {
            const int rr = row0 + r;
            const uint32_t addr = stage_base + rr * 128 + t * 16384 + ((cc ^ (rr & 7)) << 4);
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                         : "r"(addr));
            const float sc = t == 0 ? s0 : s1;
            uint32_t* w = reinterpret_cast<uint32_t*>(&val);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
              const __half2 h2 = __floats2half2_rn(f.x * sc, f.y * sc);
              w[i] = *reinterpret_cast<const uint32_t*>(&h2);
            }
          }
          if (r < rows_valid) dst[idx] = val;
        }
      } else if (q < p.Nq) {
        __half* pr = p.probs + (static_cast<long long>(bh) * p.Nq + q) * p.probs_ld;
        for (int c = 0; c < p.Nk; ++c) {
          const int t = c >> 6, cc = c & 63;
          const uint32_t addr = stage_base + row * 128 + t * 16384 + (((cc >> 3) ^ sw) << 4) + (cc & 7) * 2;
          unsigned short u;
          asm volatile("ld.shared.u16 %0, [%1];" : "=h"(u) : "r"(addr));
          pr[c] = __float2half_rn(__half2float(__ushort_as_half(u)) * (t == 0 ? sc0 : inv));
        }
      }
    }
    __half* orow = p.out + (static_cast<long long>(b) * p.Nq + q) * p.out_ld + h * D;
#pragma unroll 1
    for (int c0 = 0; c0 < DP; c0 += 16) {
      float o[16];
      tmem_ld16(t0 + OFF_O + c0, o);
      tmem_ld_wait();
      if (q < p.Nq) {
        __half hv[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) hv[i] = __float2half_rn(o[i] * inv);
        if (c0 + 16 <= D) {
          reinterpret_cast<uint4*>(orow + c0)[0] = reinterpret_cast<const uint4*>(hv)[0];
          reinterpret_cast<uint4*>(orow + c0)[1] = reinterpret_cast<const uint4*>(hv)[1];
        } else if (c0 + 8 <= D) {
          reinterpret_cast<uint4*>(orow + c0)[0] = reinterpret_cast<const uint4*>(hv)[0];
          for (int i = 8; i < 16; ++i)
            if (c0 + i < D) orow[c0 + i] = hv[i];
        } else {
          for (int i = 0; i < 16; ++i)
            if (c0 + i < D) orow[c0 + i] = hv[i];
        }
      }
    }
  }
  ICD_CTA_STAMP(4)
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int D, int MT, int POLY>
static int launch_attention(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttnParams& p,
                            cudaStream_t st) {
  using Cfg = AttnCfg<D, MT>;
  static PerDeviceFlag configured;
  if (!configured.cur()) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel<D, MT, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(std::string("attention cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    configured.cur() = true;
  }
  const int grid = p.B * p.H * ((p.Nq + 128 * Cfg::MT - 1) / (128 * Cfg::MT));
  launch_k(attention_tc_kernel<D, MT, POLY>, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, st, tq, tk, tv, p);
  return check_launch("attention_tc");
}

}  // namespace icd

using namespace icd;

#ifdef ICD_ATTN_PROFILE
static long long* g_prof_buf_last = nullptr;
extern "C" void icd_attention_prof_dump(int n_kv) {
  cudaDeviceSynchronize();
  if (g_prof_buf_last == nullptr) return;
  const char* names[7] = {"wait S", "ld first half", "max+decide", "ld rest/redo", "exponentials", "st+arrive", "loop"};
  for (int m = 0; m < 2; ++m) {
    printf("  tile %d cycles/key-tile:", m);
    for (int k = 0; k < 7; ++k) printf("  %s %.0f", names[k], double(g_prof_buf_last[m * 8 + k]) / n_kv);
    printf("\n");
  }
  const long long* c = g_prof_buf_last + 16;
  printf("  CTA timeline (cycles since CTA start, softmax warp 0 of tile 0): set-up done %lld | Q landed %lld | first S %lld | "
         "key loop done %lld | epilogue done %lld\n", c[0], c[1], c[2], c[3], c[4]);
}
#endif

extern "C" int icd_attention(const void* q, const void* k, const void* v, void* out, int B, int H, int Nq, int Nk,
                             int D, long long q_ld, long long k_ld, long long v_ld, long long out_ld, float scale,
                             void* probs_out, long long probs_ld, void* stream) {
  return icd_attention_ex(q, k, v, out, B, H, Nq, Nk, D, q_ld, k_ld, v_ld, out_ld, scale, probs_out, probs_ld, nullptr,
                          stream);
}

extern "C" int icd_attention_ex(const void* q, const void* k, const void* v, void* out, int B, int H, int Nq, int Nk,
                                int D, long long q_ld, long long k_ld, long long v_ld, long long out_ld, float scale,
                                void* probs_out, long long probs_ld, float* stats_out, void* stream) {
  if (q == nullptr || k == nullptr || v == nullptr || out == nullptr) return set_error("icd_attention: null operand");
  if (B <= 0 || H <= 0 || Nq <= 0 || Nk <= 0) return set_error("icd_attention: empty problem");
  if (probs_out != nullptr && Nk > 128)
    return set_error("icd_attention: probability capture is fused only for N_kv <= 128 (use the explicit path)");
  if ((D % 8) != 0) return set_error("icd_attention: head dim must be a multiple of 8");
  if (Nk <= 80) {   // cross-attention over the text context: one CTA walks a run of query tiles (attention_smallkv.cu)
    const int r = attention_smallkv_dispatch(q, k, v, out, B, H, Nq, Nk, D, q_ld, k_ld, v_ld, out_ld, scale, probs_out,
                                             probs_ld, stats_out, reinterpret_cast<cudaStream_t>(stream));
    if (r >= 0) return r;
  }
  CUtensorMap tq, tk, tv;
  // tensor maps over [B][N][H][D] as dims (d, head, token, batch): strides grow monotonically
  const uint32_t box[4] = {64, 1, 128, 1};      // Q: 128 query rows
  const uint32_t boxkv[4] = {64, 1, 64, 1};     // K, V: 64 keys per tile
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
  AttnParams p;
  p.B = B; p.H = H; p.Nq = Nq; p.Nk = Nk;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__half*>(out);
  p.out_ld = out_ld;
  p.probs = reinterpret_cast<__half*>(probs_out);
  p.probs_ld = probs_ld;
  p.stats = stats_out;
  static const int spec_env = [] { const char* e = getenv("ICD_ATTN_SPEC"); return e ? atoi(e) : 1; }();
  p.speculate = spec_env;
#ifdef ICD_ATTN_PROFILE
  static long long* prof_buf = nullptr;
  if (prof_buf == nullptr) cudaMallocManaged(&prof_buf, 32 * sizeof(long long));
  p.prof = prof_buf;
  g_prof_buf_last = prof_buf;
  for (int i = 0; i < 32; ++i) prof_buf[i] = 0;
#endif
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // Tuning knobs, fixed per head dim from the sweeps in profiles/README.md (r1h; r2_c with the speculative softmax,
  // where the MUFU-only variant wins for every head dim); the environment overrides are for re-running the sweeps: ICD_ATTN_POLY = exponentials per 8 on the FMA pipe (0, 2, 3), ICD_ATTN_MT = query tiles
  // per CTA (1, 2; 2 only for d <= 64).
  static const int poly_env = [] { const char* e = getenv("ICD_ATTN_POLY"); return e ? atoi(e) : -1; }();
  static const int mt_env = [] { const char* e = getenv("ICD_ATTN_MT"); return e ? atoi(e) : -1; }();
#define ICD_ATTN_POLY_SWITCH(DD, MM, DEF)                                          \
  switch (poly_env >= 0 ? poly_env : DEF) {                                        \
    case 0: return launch_attention<DD, MM, 0>(tq, tk, tv, p, st);                 \
    case 2: return launch_attention<DD, MM, 2>(tq, tk, tv, p, st);                 \
    default: return launch_attention<DD, MM, 3>(tq, tk, tv, p, st);                \
  }
  switch (D) {
    case 40:   // self-attention (long K/V stream): two query tiles per CTA; cross-attention: more, smaller CTAs
      if ((mt_env >= 0 ? mt_env : (Nk > 128 ? 2 : 1)) == 2) { ICD_ATTN_POLY_SWITCH(40, 2, 0) }
      ICD_ATTN_POLY_SWITCH(40, 1, 0)
    case 64:
      if ((mt_env >= 0 ? mt_env : 1) == 2) { ICD_ATTN_POLY_SWITCH(64, 2, 0) }
      ICD_ATTN_POLY_SWITCH(64, 1, 0)
    case 80: ICD_ATTN_POLY_SWITCH(80, 1, 0)
    case 160: ICD_ATTN_POLY_SWITCH(160, 1, 0)
    default: return set_error("icd_attention: unsupported head dim (40, 64, 80, 160)");
  }
#undef ICD_ATTN_POLY_SWITCH
}
