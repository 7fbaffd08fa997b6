// Persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[M,N] = epilogue( alpha * A[M,K] . B[N,K]^T )        fp16 operands, fp32 accumulation in TMEM
//
// One kernel serves every dense contraction of the iCD U-Net row-forward:
//   * linears (to_q/k/v/out, FF, proj_in/out, time MLPs)           A = tokens x channels (2-D)
//   * 3x3 convolutions as implicit GEMM over NHWC activations      A = 4-D TMA box shifted per filter tap,
//     (ResnetBlock2D conv1/conv2, Up/Downsample, conv_in/out)          halo zero-filled by TMA OOB handling
//   * channel-concat inputs (U-Net skip connections) without a copy     two A tensor maps split along K
//   * batched QK^T / PV for the explicit-probabilities attention path   z = (batch, head) grid dimension
//
// Warp roles (192 threads, 1 CTA/SM, persistent over output tiles):
//   warp 0      TMA producer      (one elected lane)     global -> 128B-swizzled smem ring, mbarrier tx-count
//   warp 1      MMA issuer        (one elected lane)     tcgen05.mma cta_group::1, M=128, N=BN, K=16 x4 per stage
//   warps 2..5  epilogue          TMEM -> registers -> fused epilogue -> global (double-buffered accumulators)
#pragma once
#include "icd_ptx.cuh"

namespace icd {

enum : int { GEMM_A_TILED = 0, GEMM_A_CONV3X3 = 1 };
enum : int { GEMM_OUT_ROWMAJOR = 0, GEMM_OUT_TRANSPOSED = 1 };

struct GemmParams {
  // problem
  int M, N;          // output rows / logical B rows (before GEGLU halving) per batch entry z
  int num_kb;        // number of 64-wide K blocks
  int kb_per_tap;    // conv: K blocks per filter tap (= ceil((C0+C1)/64)); tiled: = num_kb
  int kb_split;      // K blocks (within a tap) served by tensor map A0; the rest come from A1 (virtual concat)
  int a_mode;        // GEMM_A_TILED | GEMM_A_CONV3X3
  int Z, ZA1, ZB1;   // batch entries; A coords are (k, m, z % ZA1, z / ZA1), B coords (k, n, z % ZB1, z / ZB1)
  int b_mn_major;    // B tile is [K][N] (N contiguous) instead of [N][K]
  int b_batched;     // B has batch coordinates (else every z reads the same B)
  int splits;        // split-K: each tile is computed by `splits` CTAs over disjoint K ranges (fp32 partials)
  int kb_per_split;
  long long split_out_stride;  // elements between the partial outputs of consecutive splits
  int split_z;                 // staged fp32 epilogue: split s is stored at 4th TMA coordinate s * split_z
  // conv geometry (pixels of one 128-row M tile = tile_b images x tile_h rows x tile_w cols)
  int H, W, tile_w, tile_h, tile_b;
  // epilogue
  float alpha;
  const float* bias;      // [N] (permuted like B rows when geglu) or null
  const float* rowvec;    // [M / rows_per_img, ldv] added per image (ResnetBlock2D time_emb_proj term) or null
  int rows_per_img;       // rows per image for rowvec / NCHW stores (>= 1)
  int ldv;
  const __half* residual; // [M, ldr] fp16 added after everything else, or null
  long long ldr;
  long long res_zstride;
  void* out;
  long long ldc;          // row stride (row-major) / column stride (transposed) in elements
  long long out_z1_stride, out_z2_stride;  // batch-entry strides in elements: offset = (z % ZA1)*z1 + (z / ZA1)*z2
  long long out_imgstride;// transposed mode: stride between images (rows_per_img rows each)
  int out_fp32;
  int out_mode;           // GEMM_OUT_ROWMAJOR | GEMM_OUT_TRANSPOSED
  int geglu;              // out[:, j] = h_j * gelu(g_j); B rows are packed per N tile as [BN/2 h | BN/2 g]
  int epi_tma;            // 1: fp16 row-major output staged in smem (64B swizzle) and written with TMA stores;
                          //    the residual tile is prefetched into the same staging buffer by TMA
  int res_tma;            // residual present (epi_tma mode)
  int res_mma;            // EPI_WARP: the residual is added on the TENSOR CORE — behind the K loop of every tile the
                          // producer streams the [BM x 64] residual atoms of the tile through the A ring (tmRes: 128B
                          // swizzle, box 64 x 128) and the issuer runs M128 x N16 x K16 MMAs of each 16-column chunk
                          // against a 16x16 identity B tile: D[:, c..c+15] += R[:, c..c+15] . I  (exact in fp32)
  // fused consistency update (predicted_origin, utils/generation.py:136-155) on the conv_out tile:
  //   x_s = alpha_s * (x_t - sigma_t*eps) / alpha_t + sigma_s * eps      (same layout as the transposed output)
  const float* upd_x;     // current latent x_t (fp32, NCHW) or null
  float* upd_out;         // next latent x_s (fp32, NCHW)
  float alpha_t, sigma_t, alpha_s, sigma_s;
  // softmax-from-statistics (staged fp16 epilogue): out = exp2(alpha*acc - s.x) * s.y, s = exp_stats[z*M + row]
  const float2* exp_stats;
#ifdef ICD_GEMM_PROFILE
  long long* prof;        // [gridDim.x][32] clock64 stamps (debug builds: make GPROF=1, tools/gemm_prof.py)
  int dbg;                // ICD_EPI_DEBUG bit mask: 1 skip the TMA store, 2 skip tcgen05.ld, 4 skip the bias/rowvec loads,
                          // 8 skip fence.proxy.async, 16 skip st.shared, 32 skip bulk_wait + __syncwarp
#endif
};

#ifdef ICD_GEMM_PROFILE
#define ICD_DBG(bit) ((p.dbg & (bit)) != 0)
#define ICD_GSTAMP(slot)                                                                       \
  do {                                                                                         \
    if (p.prof != nullptr && (slot) < 32) p.prof[blockIdx.x * 32 + (slot)] = clock64();        \
  } while (0)
#else
#define ICD_DBG(bit) false
#define ICD_GSTAMP(slot) \
  do {                   \
  } while (0)
#endif

enum : int { EPI_DIRECT = 0, EPI_STAGED = 1, EPI_STAGED_GEGLU = 2, EPI_STAGED_RES = 3, EPI_STAGED_F32 = 4,
             EPI_WARP = 5, EPI_WARP_GEGLU = 6 };

template <int BM_, int BN, int EPI>
struct GemmCfg {
  static constexpr int BM = BM_, BK = 64;
  static constexpr int HALVES = BM / 128;            // 128-row MMA halves per tile (1 or 2)
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int C_BYTES = 4 * 8192;  // epilogue staging: 4 atoms of [128 rows x 32 cols] fp16, 64B swizzle
  // EPI_STAGED_RES: TMA-loaded residual units (2 x 16 KB). EPI_WARP: the 16x16 identity B tile (128B-swizzled,
  // 1024-byte aligned) of the tensor-core residual add
  static constexpr int R_BYTES = (EPI == EPI_STAGED_RES) ? 4 * 8192 : (EPI == EPI_WARP ? 2048 : 0);
  static constexpr int BAR_BYTES = 512;
  static constexpr int BUDGET = 232448 - 1024 - BAR_BYTES - C_BYTES - R_BYTES;
  static constexpr int STAGES = BUDGET / STAGE_BYTES > 8 ? 8 : BUDGET / STAGE_BYTES;
  static constexpr int HALF_STRIDE = BN <= 64 ? 64 : (BN <= 128 ? 128 : 256);  // TMEM columns per 128-row half
  static constexpr int ACC_COLS = HALVES * HALF_STRIDE;                        // one accumulator set
  static constexpr int ACC_STAGES = (2 * ACC_COLS <= 512) ? 2 : 1;             // double-buffer when TMEM allows
  static constexpr int TMEM_COLS = ACC_STAGES * ACC_COLS;                      // power of two in [64, 512]
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + C_BYTES + R_BYTES + 1024 /*align*/ + BAR_BYTES;
  static constexpr int EPI_THREADS = 256;            // 8 epilogue warps: two per TMEM lane quadrant
  static constexpr int THREADS = 64 + EPI_THREADS + ((EPI == EPI_STAGED_RES) ? 32 : 0);  // + residual-producer warp
};

template <int BM, int BN, int EPI>
__global__ void __launch_bounds__(GemmCfg<BM, BN, EPI>::THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
               const __grid_constant__ CUtensorMap tmRes, const GemmParams p) {
  using Cfg = GemmCfg<BM, BN, EPI>;
  constexpr int HALVES = Cfg::HALVES, ACC_STAGES = Cfg::ACC_STAGES;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
  uint8_t* smem_c = smem + STAGES * Cfg::STAGE_BYTES;
  uint8_t* smem_r = smem_c + Cfg::C_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_r + Cfg::R_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint64_t* res_full = bars + 2 * STAGES + 4;    // [2] residual unit landed in smem_r (TMA tx-count)
  uint64_t* res_empty = bars + 2 * STAGES + 6;   // [2] residual unit consumed by the 128 epilogue threads
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 8);
  static_assert((2 * 8 + 9) * 8 <= Cfg::BAR_BYTES, "barrier block");

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  pdl_launch_dependents();   // the next kernel may become resident as CTAs of this one retire (its prologue overlaps our tail)
  if (threadIdx.x == 0) ICD_GSTAMP(0);
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], Cfg::EPI_THREADS);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&res_full[i], 1);
      mbar_init(&res_empty[i], Cfg::EPI_THREADS);
    }

    tma_prefetch_desc(&tmOut);
    tma_prefetch_desc(&tmRes);
    fence_mbar_init();
  } else if (warp == 1) {
    tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr_smem);
  } else if (warp == 2) {
    if constexpr (EPI == EPI_WARP) {
      // 16x16 fp16 identity as a K-major, 128B-swizzled B tile (16 rows x 128 B; only the first K16 chunk is read)
      uint32_t* idt = reinterpret_cast<uint32_t*>(smem_r);
      for (int i = lane; i < 512; i += 32) idt[i] = 0u;
      __syncwarp();
      if (lane < 16) {
        const int n = lane;
        const uint32_t off = n * 128 + ((((n * 2) >> 4) ^ (n & 7)) << 4) + ((n * 2) & 15);
        *reinterpret_cast<__half*>(smem_r + off) = __float2half_rn(1.0f);
      }
      fence_proxy_async_smem();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (threadIdx.x == 0) ICD_GSTAMP(1);
  pdl_wait();                // everything above ran under the previous kernel's tail; global memory is touched below
  if (threadIdx.x == 0) ICD_GSTAMP(2);

  const int m_tiles = (p.M + BM - 1) / BM;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int tiles_per_z = m_tiles * n_tiles;
  const int tiles_all_z = tiles_per_z * p.Z;
  const int total_tiles = tiles_all_z * p.splits;   // split index is the slowest tile coordinate

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int split = tile / tiles_all_z;
        const int tz = tile - split * tiles_all_z;
        const int z = tz / tiles_per_z;
        const int t = tz - z * tiles_per_z;
        const int mt = t / n_tiles, nt = t - mt * n_tiles;
        const int n0 = nt * BN;
        const int kb_begin = p.splits > 1 ? split * p.kb_per_split : 0;
        const int kb_end = p.splits > 1 ? min(p.num_kb, kb_begin + p.kb_per_split) : p.num_kb;
        // origin of each 128-row half (conv: image / row / column of its first pixel)
        int m0h[HALVES], cb[HALVES], cy[HALVES], cx[HALVES];
#pragma unroll
        for (int h = 0; h < HALVES; ++h) {
          const int st = mt * HALVES + h;   // 128-row sub-tile index
          m0h[h] = st * 128;
          cb[h] = cy[h] = cx[h] = 0;
          if (p.a_mode == GEMM_A_CONV3X3) {
            const int hw = p.H * p.W;
            if (p.tile_b > 1) {
              cb[h] = st * p.tile_b;
            } else {
              cb[h] = m0h[h] / hw;
              const int rem = m0h[h] - cb[h] * hw;
              cy[h] = rem / p.W;
              cx[h] = rem - cy[h] * p.W;
            }
          }
        }
        const int bz1 = p.b_batched ? z % p.ZB1 : 0, bz2 = p.b_batched ? z / p.ZB1 : 0;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          const int tap = kb / p.kb_per_tap;
          const int r = kb - tap * p.kb_per_tap;
          const CUtensorMap* ma = (r < p.kb_split) ? &tmA0 : &tmA1;
          const int ka = (r < p.kb_split ? r : r - p.kb_split) * 64;
          uint8_t* sa = smem_a + stage * Cfg::A_BYTES;
          void* sb = smem_b + stage * Cfg::B_BYTES;
#pragma unroll
          for (int h = 0; h < HALVES; ++h) {
            if (p.a_mode == GEMM_A_CONV3X3) {
              const int ky = tap / 3, kx = tap - ky * 3;
              tma_load_4d(sa + h * 16384, ma, &full_bar[stage], ka, cx[h] + kx - 1, cy[h] + ky - 1, cb[h]);
            } else {
              tma_load_4d(sa + h * 16384, ma, &full_bar[stage], ka, m0h[h], z % p.ZA1, z / p.ZA1);
            }
          }
          // stream the weight tile PF K-blocks ahead into L2 (cold weights, few CTAs: DRAM latency bound otherwise)
          constexpr int PF = 24;
          if (!p.b_mn_major && !p.b_batched && kb + PF < kb_end) tma_prefetch_l2_4d(&tmB, (kb + PF) * 64, n0, 0, 0);
          if (p.b_mn_major) {
            // B tile = BN/64 boxes of [64 K rows][64 N], one per 64-wide N atom
#pragma unroll
            for (int j = 0; j < BN / 64; ++j)
              tma_load_4d(reinterpret_cast<uint8_t*>(sb) + j * 8192, &tmB, &full_bar[stage], n0 + j * 64, kb * 64,
                          bz1, bz2);
          } else {
            tma_load_4d(sb, &tmB, &full_bar[stage], kb * 64, n0, bz1, bz2);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if constexpr (EPI == EPI_WARP) {
          if (p.res_mma) {   // residual atoms [BM x 64] of this tile, streamed through the A half of the ring
            const int ncols = min(BN, p.N - n0);
            for (int atom = 0; atom * 64 < ncols; ++atom) {
              mbar_wait(&empty_bar[stage], phase ^ 1);
              mbar_expect_tx(&full_bar[stage], HALVES * 16384);
#pragma unroll
              for (int h = 0; h < HALVES; ++h)
                tma_load_4d(smem_a + stage * Cfg::A_BYTES + h * 16384, &tmRes, &full_bar[stage], n0 + atom * 64, m0h[h], 0,
                            0);
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
      ICD_GSTAMP(3);
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      // NOTE on issue overhead: building a 64-bit smem descriptor from an address costs ~15 dependent uniform-
      // datapath instructions (~120 cycles) — more than the MMA itself for N <= 128 (measured: ~630 cycles per
      // K block regardless of tile width). The descriptors are therefore formed by adding small constants to
      // precomputed 32-bit low words; the high word (SBO, version, swizzle mode) never changes.
      const uint32_t idesc = p.b_mn_major ? umma_idesc_f16(128, BN, false, true) : umma_idesc_f16(128, BN, false, false);
      const uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);            // SBO=1024 B | version 1 | SW128
      const uint32_t a_lo0 = ((smem_u32(smem_a) >> 4) & 0x3FFFu) | (1u << 16);    // LBO field = 1 (unused, K-major)
      const uint32_t b_lo0 = ((smem_u32(smem_b) >> 4) & 0x3FFFu) | ((p.b_mn_major ? (8192u >> 4) : 1u) << 16);
      const uint32_t b_kstep = p.b_mn_major ? (2048u >> 4) : 2u;                  // 16 K rows (MN-major) or 32 B
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
        const int acc = iter % ACC_STAGES;
        const uint32_t acc_phase = (iter / ACC_STAGES) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_COLS;
        const int split = tile / tiles_all_z;
        const int kb_begin = p.splits > 1 ? split * p.kb_per_split : 0;
        const int kb_end = p.splits > 1 ? min(p.num_kb, kb_begin + p.kb_per_split) : p.num_kb;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (iter == 0 && kb == kb_begin) ICD_GSTAMP(4);
          const uint32_t a_lo = a_lo0 + stage * (Cfg::A_BYTES >> 4);
          const uint32_t b_lo = b_lo0 + stage * (Cfg::B_BYTES >> 4);
          const uint32_t first = (kb > kb_begin) ? 1u : 0u;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t db = (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + k * b_kstep);
#pragma unroll
            for (int h = 0; h < HALVES; ++h) {
              const uint64_t da = (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + h * (16384u >> 4) + k * 2u);
              umma_f16_ss(d_tmem + h * Cfg::HALF_STRIDE, da, db, idesc, k == 0 ? first : 1u);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if constexpr (EPI == EPI_WARP) {
          if (p.res_mma) {   // D[:, c .. c+15] += R[:, c .. c+15] . I16 for every 16-column chunk of the tile
            const int tz = tile - split * tiles_all_z;
            const int tt = tz - (tz / tiles_per_z) * tiles_per_z;
            const int ncols = min(BN, p.N - (tt % n_tiles) * BN);
            const uint32_t idesc16 = umma_idesc_f16(128, 16, false, false);
            const uint64_t db_id = (static_cast<uint64_t>(desc_hi) << 32) | (((smem_u32(smem_r) >> 4) & 0x3FFFu) | (1u << 16));
            for (int atom = 0; atom * 64 < ncols; ++atom) {
              mbar_wait(&full_bar[stage], phase);
              tc_fence_after();
              const uint32_t a_lo = a_lo0 + stage * (Cfg::A_BYTES >> 4);
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                if (atom * 64 + c * 16 < ncols) {
#pragma unroll
                  for (int h = 0; h < HALVES; ++h)
                    umma_f16_ss(d_tmem + h * Cfg::HALF_STRIDE + atom * 64 + c * 16,
                                (static_cast<uint64_t>(desc_hi) << 32) | (a_lo + h * (16384u >> 4) + c * 2u), db_id,
                                idesc16, 1u);
                }
              }
              umma_commit(&empty_bar[stage]);
              if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
        umma_commit(&tmem_full[acc]);
        ICD_GSTAMP(8 + 3 * iter);          // all MMAs of tile `iter` issued
      }
      ICD_GSTAMP(5);
    }
  } else if (warp == 10) {
    // ------------------------------------------------------------------ residual producer (EPI_STAGED_RES only)
    if constexpr (EPI == EPI_STAGED_RES) {
      if (elect_one()) {
        constexpr int units = (BN + 63) / 64;
        uint32_t runit = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
          const int tz = tile % tiles_all_z;
          const int z = tz / tiles_per_z;
          const int t = tz - z * tiles_per_z;
          const int mt = t / n_tiles, nt = t - mt * n_tiles;
          for (int half = 0; half < HALVES; ++half) {
            for (int u = 0; u < units; ++u, ++runit) {
              const int b = runit & 1;
              mbar_wait(&res_empty[b], ((runit >> 1) & 1) ^ 1);
              const int cols = min(64, BN - u * 64);
              const int natoms = (cols + 31) / 32;
              mbar_expect_tx(&res_full[b], natoms * 8192);
              for (int a = 0; a < natoms; ++a)
                tma_load_4d(smem_r + b * 16384 + a * 8192, &tmRes, &res_full[b], nt * BN + u * 64 + a * 32,
                            mt * BM + half * 128, 0, 0);
            }
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    // NOTE: loops are deliberately rolled (#pragma unroll 1) — a fully unrolled epilogue is ~200 KB of SASS and
    // runs out of the instruction cache (measured: 21k cycles per tile, ncu profiles/r1_c_*).
    const int quad = warp & 3;           // TMEM lane quadrant this warp may access
    const int part = (warp - 2) >> 2;    // the two warps of a quadrant split every 64-column unit (32 columns each)
    const int row_in_tile = quad * 32 + lane;
    int iter = 0;
    if constexpr (EPI == EPI_STAGED_F32) {
      // ---- staged fp32 epilogue (split-K partials, fp32 row-major outputs): alpha * acc (+ bias) -> 128B-swizzled
      // smem units of [128 rows x 32 fp32 columns], double-buffered, written out with TMA stores.
      constexpr int units = (BN + 31) / 32;
      const bool leader = (warp == 2) && elect_one();
      const uint32_t stg = smem_u32(smem_c);
      const uint32_t sw = row_in_tile & 7;
      uint32_t unit = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
        const int split = tile / tiles_all_z;
        const int tz = tile - split * tiles_all_z;
        const int z = tz / tiles_per_z;
        const int t = tz - z * tiles_per_z;
        const int mt = t / n_tiles, nt = t - mt * n_tiles;
        const int acc = iter % ACC_STAGES;
        const uint32_t acc_phase = (iter / ACC_STAGES) & 1;
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int half = 0; half < HALVES; ++half) {
          const uint32_t t_addr = tmem_base + acc * Cfg::ACC_COLS + half * Cfg::HALF_STRIDE +
                                  (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
          for (int u = 0; u < units; ++u, ++unit) {
            const uint32_t buf = stg + (unit & 1) * 16384;
            if (leader) bulk_wait_read1();
            named_bar_sync(1, Cfg::EPI_THREADS);
            const int col_t = u * 32 + part * 16;
            const int ncol = nt * BN + col_t;
            if (col_t < BN && ncol < p.N) {   // warp-uniform
              float v[16];
              tmem_ld16(t_addr + col_t, v);
              tmem_ld_wait();
#pragma unroll
              for (int cc = 0; cc < 4; ++cc) {
                float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.bias != nullptr && ncol + cc * 4 < p.N) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + ncol) + cc);
                const float o0 = v[cc * 4 + 0] * p.alpha + b4.x, o1 = v[cc * 4 + 1] * p.alpha + b4.y;
                const float o2 = v[cc * 4 + 2] * p.alpha + b4.z, o3 = v[cc * 4 + 3] * p.alpha + b4.w;
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(buf + row_in_tile * 128 +
                                                                             (((part * 4 + cc) ^ sw) << 4)),
                             "f"(o0), "f"(o1), "f"(o2), "f"(o3)
                             : "memory");
              }
            }
            if (u == units - 1 && half == HALVES - 1) {
              tc_fence_before();
              mbar_arrive(&tmem_empty[acc]);
            }
            fence_proxy_async_smem();
            named_bar_sync(2, Cfg::EPI_THREADS);
            if (leader) {
              const int ncol0 = nt * BN + u * 32;
              if (u * 32 < BN && ncol0 < p.N)
                tma_store_4d(&tmOut, smem_c + (unit & 1) * 16384, ncol0, mt * BM + half * 128, z % p.ZA1,
                             z / p.ZA1 + split * p.split_z);
              bulk_commit();
            }
          }
        }
      }
      if (leader) bulk_wait0();
    } else if constexpr (EPI == EPI_WARP) {
      // ---- per-warp staged epilogue: every epilogue warp owns a [32 rows x 32 columns] block of each 64-column unit
      // (rows = its TMEM lane quadrant, columns = its half of the unit) and runs its own pipeline
      //   tcgen05.ld -> (alpha, +bias, +rowvec | softmax-from-stats) -> 64B-swizzled smem block -> TMA store,
      // double-buffered per warp, with NO CTA-wide barrier: the only shared event is the release of the accumulator.
      // The RESIDUAL never reaches this code: it is added on the tensor core (identity MMA behind the K loop, see the
      // producer / issuer), so the accumulator already holds acc + residual in fp32.
      // The loop body is kept lean on purpose: the K = N = C transformer linears are bound by the INSTRUCTION ISSUE of
      // this epilogue (measured, profiles/r2_gemm_k320_*: ~500 instructions per 32x32 block before this rewrite).
      constexpr int outw = BN;
      constexpr int units = (outw + 63) / 64;
      const int ew = warp - 2;                       // 0..7
      uint8_t* my_out = smem_c + ew * 4096;          // 2 x [32 rows x 64 B]
      const bool lane0 = lane == 0;
      const uint32_t sw = (lane >> 1) & 3;           // 64B swizzle: 16B-chunk index ^= (row / 2) % 4
      const int col_w = part * 32;                   // this warp's column offset inside a 64-column unit
      const bool has_bias = p.bias != nullptr, has_rv = p.rowvec != nullptr, has_est = p.exp_stats != nullptr;
      const bool unit_alpha = p.alpha == 1.0f;
      uint32_t ocount = 0;                           // staged blocks written by this warp
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
        const int tz = tile % tiles_all_z;
        const int z = tz / tiles_per_z;
        const int t = tz - z * tiles_per_z;
        const int mt = t / n_tiles, nt = t - mt * n_tiles;
        const int acc = iter % ACC_STAGES;
        const uint32_t acc_phase = (iter / ACC_STAGES) & 1;
        const int n_out0 = nt * outw;
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        if (warp == 2 && lane0) ICD_GSTAMP(9 + 3 * iter);
#pragma unroll 1
        for (int half = 0; half < HALVES; ++half) {
          const int row0 = mt * BM + half * 128 + quad * 32;
          const int row = row0 + lane;
          const float* rv = nullptr;
          if (has_rv) rv = p.rowvec + static_cast<long long>(min(row, p.M - 1) / p.rows_per_img) * p.ldv;
          float2 est = make_float2(0.f, 0.f);
          if (has_est && row < p.M) est = __ldg(p.exp_stats + static_cast<long long>(z) * p.M + row);
          const uint32_t t_half =
              tmem_base + acc * Cfg::ACC_COLS + half * Cfg::HALF_STRIDE + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
          for (int u = 0; u < units; ++u) {
            const int col_t = u * 64 + col_w;
            const int ncol = n_out0 + col_t;
            if (col_t >= outw || ncol >= p.N) continue;   // warp-uniform (BN = 160: the upper half skips unit 2)
            // this warp's staging buffer (ocount & 1) was last used two blocks ago: its TMA store must have read it
            if (lane0) bulk_wait_read1();
            __syncwarp();
            float v[32];
            tmem_ld32(t_half + col_t, v);
            tmem_ld_wait();
            if (!unit_alpha) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
            }
            if (ncol + 32 <= p.N) {   // whole block inside the matrix (always, except ragged N tails)
              if (has_bias) {
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                  const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + ncol) + j4);
                  v[j4 * 4 + 0] += b.x; v[j4 * 4 + 1] += b.y; v[j4 * 4 + 2] += b.z; v[j4 * 4 + 3] += b.w;
                }
              }
              if (has_rv) {
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                  const float4 b = __ldg(reinterpret_cast<const float4*>(rv + ncol) + j4);
                  v[j4 * 4 + 0] += b.x; v[j4 * 4 + 1] += b.y; v[j4 * 4 + 2] += b.z; v[j4 * 4 + 3] += b.w;
                }
              }
            } else {                  // N % 8 == 0 in staged mode: whole groups of 4
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                if (ncol + j4 * 4 < p.N) {
                  if (has_bias) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + ncol) + j4);
                    v[j4 * 4 + 0] += b.x; v[j4 * 4 + 1] += b.y; v[j4 * 4 + 2] += b.z; v[j4 * 4 + 3] += b.w;
                  }
                  if (has_rv) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(rv + ncol) + j4);
                    v[j4 * 4 + 0] += b.x; v[j4 * 4 + 1] += b.y; v[j4 * 4 + 2] += b.z; v[j4 * 4 + 3] += b.w;
                  }
                }
              }
            }
            if (has_est) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = exp2f(v[j] - est.x) * est.y;
            }
            const uint32_t obuf = smem_u32(my_out) + (ocount & 1) * 2048 + lane * 64;
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
              uint32_t o[4];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const __half2 h2 = __floats2half2_rn(v[cc * 8 + 2 * q], v[cc * 8 + 2 * q + 1]);
                o[q] = *reinterpret_cast<const uint32_t*>(&h2);
              }
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(obuf + ((cc ^ sw) << 4)), "r"(o[0]),
                           "r"(o[1]), "r"(o[2]), "r"(o[3])
                           : "memory");
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane0) {
              tma_store_4d(&tmOut, my_out + (ocount & 1) * 2048, ncol, row0, z % p.ZA1, z / p.ZA1);
              bulk_commit();
            }
            ++ocount;
          }
        }
        // every TMEM read of this accumulator by this thread is complete (tmem_ld_wait above)
        tc_fence_before();
        mbar_arrive(&tmem_empty[acc]);
        if (warp == 2 && lane0) ICD_GSTAMP(10 + 3 * iter);
      }
      if (lane0) bulk_wait0();
      if (warp == 2 && lane0) ICD_GSTAMP(6);
    } else if constexpr (EPI == EPI_WARP_GEGLU) {
      // ---- per-warp GEGLU epilogue: as EPI_WARP (no CTA-wide barrier, every warp stages and TMA-stores its own
      // [32 rows x 32 output columns] blocks), with  out = (h * alpha + b_h) * gelu(g * alpha + b_g)  from the hidden
      // and gate halves of the interleaved accumulator tile (columns [0, BN/2) | [BN/2, BN), packing.pack_geglu).
      constexpr int outw = BN / 2;
      constexpr int units = (outw + 63) / 64;
      const int n_out_total = p.N / 2;
      const int ew = warp - 2;                       // 0..7
      uint8_t* my_out = smem_c + ew * 4096;          // 2 x [32 rows x 64 B]
      const bool lane0 = lane == 0;
      const uint32_t sw = (lane >> 1) & 3;           // 64B swizzle: 16B-chunk index ^= (row / 2) % 4
      const int col_w = part * 32;
      const bool has_bias = p.bias != nullptr;
      uint32_t ocount = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
        const int tz = tile % tiles_all_z;
        const int z = tz / tiles_per_z;
        const int t = tz - z * tiles_per_z;
        const int mt = t / n_tiles, nt = t - mt * n_tiles;
        const int acc = iter % ACC_STAGES;
        const uint32_t acc_phase = (iter / ACC_STAGES) & 1;
        const int n_out0 = nt * outw;
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int half = 0; half < HALVES; ++half) {
          const int row0 = mt * BM + half * 128 + quad * 32;
          const uint32_t t_half =
              tmem_base + acc * Cfg::ACC_COLS + half * Cfg::HALF_STRIDE + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
          for (int u = 0; u < units; ++u) {
            const int col_t = u * 64 + col_w;
            const int ncol = n_out0 + col_t;
            if (col_t >= outw || ncol >= n_out_total) continue;   // warp-uniform
            if (lane0) bulk_wait_read1();
            __syncwarp();
            float hv[32], gv[32];
            tmem_ld32(t_half + col_t, hv);
            tmem_ld32(t_half + BN / 2 + col_t, gv);
            tmem_ld_wait();
            const float4* bh = reinterpret_cast<const float4*>(p.bias + nt * BN + col_t);
            const float4* bg = reinterpret_cast<const float4*>(p.bias + nt * BN + BN / 2 + col_t);
            const uint32_t obuf = smem_u32(my_out) + (ocount & 1) * 2048 + lane * 64;
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
              uint32_t o[4];
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                const int j4 = cc * 2 + q;
                float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
                if (has_bias) { b0 = __ldg(bh + j4); b1 = __ldg(bg + j4); }
                const float r0 = (hv[j4 * 4 + 0] * p.alpha + b0.x) * gelu_erf(gv[j4 * 4 + 0] * p.alpha + b1.x);
                const float r1 = (hv[j4 * 4 + 1] * p.alpha + b0.y) * gelu_erf(gv[j4 * 4 + 1] * p.alpha + b1.y);
                const float r2 = (hv[j4 * 4 + 2] * p.alpha + b0.z) * gelu_erf(gv[j4 * 4 + 2] * p.alpha + b1.z);
                const float r3 = (hv[j4 * 4 + 3] * p.alpha + b0.w) * gelu_erf(gv[j4 * 4 + 3] * p.alpha + b1.w);
                const __half2 h01 = __floats2half2_rn(r0, r1), h23 = __floats2half2_rn(r2, r3);
                o[q * 2] = *reinterpret_cast<const uint32_t*>(&h01);
                o[q * 2 + 1] = *reinterpret_cast<const uint32_t*>(&h23);
              }
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(obuf + ((cc ^ sw) << 4)), "r"(o[0]),
                           "r"(o[1]), "r"(o[2]), "r"(o[3])
                           : "memory");
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane0) {
              tma_store_4d(&tmOut, my_out + (ocount & 1) * 2048, ncol, row0, z % p.ZA1, z / p.ZA1);
              bulk_commit();
            }
            ++ocount;
          }
        }
        tc_fence_before();
        mbar_arrive(&tmem_empty[acc]);
      }
      if (lane0) bulk_wait0();
    } else if constexpr (EPI != EPI_DIRECT) {
      // ---- staged epilogue: TMEM -> regs -> (+bias/rowvec/residual | GEGLU) -> swizzled smem -> TMA store.
      // Output is produced in 64-column units through a double-buffered staging area (2 x [128 rows x 64 cols],
      // as 32-column atoms with 64B swizzle), so the TMA store of unit u overlaps the math of unit u+1.
      // The residual is prefetched from global memory into registers one unit ahead (the first unit of a tile
      // BEFORE waiting for the accumulator, i.e. behind the tile's main loop).
      constexpr bool GEGLU = (EPI == EPI_STAGED_GEGLU);
      constexpr int outw = GEGLU ? BN / 2 : BN;   // output columns per tile
      constexpr int units = (outw + 63) / 64;
      const int n_out_total = GEGLU ? p.N / 2 : p.N;
      const bool leader = (warp == 2) && elect_one();
      const uint32_t stg = smem_u32(smem_c);
      const uint32_t sw = (row_in_tile >> 1) & 3; // 64B swizzle: 16B-chunk index ^= (row / 2) % 4
      const bool has_res = p.res_tma != 0;
      uint32_t unit = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
        const int tz = tile % tiles_all_z;
        const int z = tz / tiles_per_z;
        const int t = tz - z * tiles_per_z;
        const int mt = t / n_tiles, nt = t - mt * n_tiles;
        const int acc = iter % ACC_STAGES;
        const uint32_t acc_phase = (iter / ACC_STAGES) & 1;
        const int n_out0 = nt * outw;
#pragma unroll 1
        for (int half = 0; half < HALVES; ++half) {
        const int row = mt * BM + half * 128 + row_in_tile;
        const bool row_ok = row < p.M;
        const __half* rp = p.residual + static_cast<long long>(row) * p.ldr + n_out0;
        uint4 rcur[4], rnxt[4];
        auto load_res = [&](uint4* dst, int u) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            dst[i] = make_uint4(0, 0, 0, 0);
            const int c = u * 64 + part * 32 + i * 8;
            if (has_res && row_ok && c < outw && n_out0 + c < n_out_total)
              dst[i] = __ldg(reinterpret_cast<const uint4*>(rp + c));
          }
        };
        if (!GEGLU && EPI != EPI_STAGED_RES) load_res(rcur, 0);
        if (half == 0) {
          mbar_wait(&tmem_full[acc], acc_phase);
          tc_fence_after();
        }
        const uint32_t t_addr = tmem_base + acc * Cfg::ACC_COLS + half * Cfg::HALF_STRIDE +
                                (static_cast<uint32_t>(quad * 32) << 16);
        const int img = min(row, p.M - 1) / p.rows_per_img;
        const float* rv = (p.rowvec != nullptr) ? p.rowvec + static_cast<long long>(img) * p.ldv : nullptr;
        float2 est = make_float2(0.f, 0.f);
        if (p.exp_stats != nullptr && row_ok) est = __ldg(p.exp_stats + static_cast<long long>(z) * p.M + row);
#pragma unroll 1
        for (int u = 0; u < units; ++u, ++unit) {
          const int unit_cols = min(64, outw - u * 64);
          const uint32_t buf = stg + (unit & 1) * 16384;
          if (!GEGLU && EPI != EPI_STAGED_RES && u + 1 < units) load_res(rnxt, u + 1);
          if constexpr (EPI == EPI_STAGED_RES) mbar_wait(&res_full[unit & 1], (unit >> 1) & 1);
          // staging buffer (unit & 1) was last used by unit-2: its TMA store must have finished reading smem
          if (leader) bulk_wait_read1();
          named_bar_sync(1, Cfg::EPI_THREADS);
          if constexpr (GEGLU) {
#pragma unroll 1
            for (int c16 = part * 32; c16 < min(unit_cols, part * 32 + 32); c16 += 16) {
              const int col_t = u * 64 + c16;
              float hv[16], gv[16];
              tmem_ld16(t_addr + col_t, hv);
              tmem_ld16(t_addr + BN / 2 + col_t, gv);
              tmem_ld_wait();
              const float4* bh = reinterpret_cast<const float4*>(p.bias + nt * BN + col_t);
              const float4* bg = reinterpret_cast<const float4*>(p.bias + nt * BN + BN / 2 + col_t);
              uint32_t o[8];
#pragma unroll
              for (int j4 = 0; j4 < 4; ++j4) {
                float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
                if (p.bias != nullptr) { b0 = __ldg(bh + j4); b1 = __ldg(bg + j4); }
                const float r0 = (hv[j4 * 4 + 0] * p.alpha + b0.x) * gelu_erf(gv[j4 * 4 + 0] * p.alpha + b1.x);
                const float r1 = (hv[j4 * 4 + 1] * p.alpha + b0.y) * gelu_erf(gv[j4 * 4 + 1] * p.alpha + b1.y);
                const float r2 = (hv[j4 * 4 + 2] * p.alpha + b0.z) * gelu_erf(gv[j4 * 4 + 2] * p.alpha + b1.z);
                const float r3 = (hv[j4 * 4 + 3] * p.alpha + b0.w) * gelu_erf(gv[j4 * 4 + 3] * p.alpha + b1.w);
                const __half2 h01 = __floats2half2_rn(r0, r1), h23 = __floats2half2_rn(r2, r3);
                o[j4 * 2] = *reinterpret_cast<const uint32_t*>(&h01);
                o[j4 * 2 + 1] = *reinterpret_cast<const uint32_t*>(&h23);
              }
              const uint32_t atom = buf + (c16 >> 5) * 8192 + row_in_tile * 64;
              const uint32_t ch = (c16 & 16) >> 3;   // first 16B chunk (0 or 2) of this 16-column group
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(atom + (((ch) ^ sw) << 4)), "r"(o[0]),
                           "r"(o[1]), "r"(o[2]), "r"(o[3]) : "memory");
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(atom + (((ch + 1) ^ sw) << 4)), "r"(o[4]),
                           "r"(o[5]), "r"(o[6]), "r"(o[7]) : "memory");
            }
          } else {
            {
              const int h32 = part;
              const int c0 = h32 * 32;
              const int col_t = u * 64 + c0;         // output column within the tile
              const int ncol = n_out0 + col_t;       // global output column
              if (c0 < unit_cols && ncol < n_out_total) {   // warp-uniform
                float v[32];
                tmem_ld32(t_addr + col_t, v);
                tmem_ld_wait();
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                  float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
                  if (ncol + j4 * 4 < n_out_total) {   // N_out % 8 == 0 in staged mode: whole groups
                    if (p.bias != nullptr) b0 = __ldg(reinterpret_cast<const float4*>(p.bias + ncol) + j4);
                    if (rv != nullptr) b1 = __ldg(reinterpret_cast<const float4*>(rv + ncol) + j4);
                  }
                  v[j4 * 4 + 0] = v[j4 * 4 + 0] * p.alpha + (b0.x + b1.x);
                  v[j4 * 4 + 1] = v[j4 * 4 + 1] * p.alpha + (b0.y + b1.y);
                  v[j4 * 4 + 2] = v[j4 * 4 + 2] * p.alpha + (b0.z + b1.z);
                  v[j4 * 4 + 3] = v[j4 * 4 + 3] * p.alpha + (b0.w + b1.w);
                }
                if (p.exp_stats != nullptr) {   // warp-uniform
#pragma unroll
                  for (int j = 0; j < 32; ++j) v[j] = exp2f(v[j] - est.x) * est.y;
                }
                const uint32_t atom = buf + h32 * 8192 + row_in_tile * 64;
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                  uint4 rr;
                  if constexpr (EPI == EPI_STAGED_RES) {
                    const uint32_t ra = smem_u32(smem_r) + (unit & 1) * 16384 + h32 * 8192 + row_in_tile * 64 +
                                        ((cc ^ sw) << 4);
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(rr.x), "=r"(rr.y), "=r"(rr.z), "=r"(rr.w)
                                 : "r"(ra));
                  } else {
                    rr = rcur[cc];
                  }
                  const uint32_t rw[4] = {rr.x, rr.y, rr.z, rr.w};
                  uint32_t o[4];
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&rw[q]));
                    const __half2 h2 = __floats2half2_rn(v[cc * 8 + 2 * q] + f.x, v[cc * 8 + 2 * q + 1] + f.y);
                    o[q] = *reinterpret_cast<const uint32_t*>(&h2);
                  }
                  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(atom + ((cc ^ sw) << 4)), "r"(o[0]),
                               "r"(o[1]), "r"(o[2]), "r"(o[3])
                               : "memory");
                }
              }
            }
            if constexpr (EPI == EPI_STAGED_RES) {
              mbar_arrive(&res_empty[unit & 1]);
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i) rcur[i] = rnxt[i];
            }
          }
          if (u == units - 1 && half == HALVES - 1) {   // all TMEM reads of this accumulator are done
            tc_fence_before();
            mbar_arrive(&tmem_empty[acc]);
          }
          fence_proxy_async_smem();
          named_bar_sync(2, Cfg::EPI_THREADS);
          if (leader) {
#pragma unroll
            for (int h32 = 0; h32 < 2; ++h32) {
              const int ncol = n_out0 + u * 64 + h32 * 32;
              if (h32 * 32 < unit_cols && ncol < n_out_total)
                tma_store_4d(&tmOut, smem_c + (unit & 1) * 16384 + h32 * 8192, ncol, mt * BM + half * 128,
                             z % p.ZA1, z / p.ZA1);
            }
            bulk_commit();
          }
        }
        }  // half
      }
      if (leader) bulk_wait0();
    } else {
      // ---- direct epilogue (fp32 / transposed / unaligned outputs: conv_out, time-embedding GEMMs, N tails)
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
        const int split = tile / tiles_all_z;
        const int tz = tile - split * tiles_all_z;
        const int z = tz / tiles_per_z;
        const int t = tz - z * tiles_per_z;
        const int mt = t / n_tiles, nt = t - mt * n_tiles;
        const int acc = iter % ACC_STAGES;
        const uint32_t acc_phase = (iter / ACC_STAGES) & 1;
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const long long out_zoff = (z % p.ZA1) * p.out_z1_stride + (z / p.ZA1) * p.out_z2_stride +
                                   split * p.split_out_stride;
#pragma unroll 1
        for (int half = 0; half < HALVES; ++half) {
        const uint32_t t_addr = tmem_base + acc * Cfg::ACC_COLS + half * Cfg::HALF_STRIDE +
                                (static_cast<uint32_t>(quad * 32) << 16);
        const int row = mt * BM + half * 128 + row_in_tile;
        const bool row_ok = row < p.M;
        const int img = row / p.rows_per_img;
        const int row_in_img = row - img * p.rows_per_img;
        const float* rv = (p.rowvec != nullptr && row_ok) ? p.rowvec + static_cast<long long>(img) * p.ldv : nullptr;
        const __half* res = (p.residual != nullptr && row_ok)
                                ? p.residual + z * p.res_zstride + static_cast<long long>(row) * p.ldr
                                : nullptr;
        const int n0 = nt * BN;
#pragma unroll 1
        for (int c0 = part * 16; c0 < BN; c0 += 32) {
          if (n0 + c0 >= p.N) break;  // warp-uniform
          float v[16];
          tmem_ld16(t_addr + c0, v);
          tmem_ld_wait();
          if (row_ok) {
            const int nbase = n0 + c0;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float x = v[j] * p.alpha;
              if (nbase + j < p.N) {
                if (p.bias != nullptr) x += p.bias[nbase + j];
                if (rv != nullptr) x += rv[nbase + j];
                if (res != nullptr) x += __half2float(res[nbase + j]);
              }
              v[j] = x;
            }
            if (p.out_mode == GEMM_OUT_ROWMAJOR) {
              const long long o0 = out_zoff + static_cast<long long>(row) * p.ldc + nbase;
              if (p.out_fp32 && nbase + 16 <= p.N &&
                  (reinterpret_cast<uintptr_t>(reinterpret_cast<float*>(p.out) + o0) & 15) == 0) {
                float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + o0);   // 64 B per thread
#pragma unroll
                for (int j = 0; j < 4; ++j) op[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  if (nbase + j < p.N) {
                    if (p.out_fp32) reinterpret_cast<float*>(p.out)[o0 + j] = v[j];
                    else reinterpret_cast<__half*>(p.out)[o0 + j] = __float2half_rn(v[j]);
                  }
                }
              }
            } else {
              // transposed: element (row, n) -> out[z][img][n][row_in_img]; consecutive lanes = consecutive rows
              const long long base = out_zoff + static_cast<long long>(img) * p.out_imgstride + row_in_img;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                if (nbase + j < p.N) {
                  const long long idx = base + static_cast<long long>(nbase + j) * p.ldc;
                  if (p.out_fp32) {
                    reinterpret_cast<float*>(p.out)[idx] = v[j];
                    if (p.upd_x != nullptr) {
                      const float eps = v[j];
                      const float x0 = (p.upd_x[idx] - p.sigma_t * eps) / p.alpha_t;
                      p.upd_out[idx] = p.alpha_s * x0 + p.sigma_s * eps;
                    }
                  } else {
                    reinterpret_cast<__half*>(p.out)[idx] = __float2half_rn(v[j]);
                  }
                }
              }
            }
          }
        }
        }  // half
        tc_fence_before();
        mbar_arrive(&tmem_empty[acc]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) ICD_GSTAMP(7);
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

}  // namespace icd
