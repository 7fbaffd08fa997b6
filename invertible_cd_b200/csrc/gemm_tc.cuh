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
  // fused consistency update (predicted_origin, utils/generation.py:136-155) on the conv_out tile:
  //   x_s = alpha_s * (x_t - sigma_t*eps) / alpha_t + sigma_s * eps      (same layout as the transposed output)
  const float* upd_x;     // current latent x_t (fp32, NCHW) or null
  float* upd_out;         // next latent x_s (fp32, NCHW)
  float alpha_t, sigma_t, alpha_s, sigma_s;
};

template <int BN>
struct GemmCfg {
  static constexpr int BM = 128, BK = 64;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int C_BYTES = 4 * 8192;  // epilogue staging: 4 atoms of [128 rows x 32 cols] fp16, 64B swizzle
  static constexpr int BUDGET = 232448 - 1024 - 256 - C_BYTES;
  static constexpr int STAGES = BUDGET / STAGE_BYTES > 8 ? 8 : BUDGET / STAGE_BYTES;
  static constexpr int ACC_STRIDE = BN <= 64 ? 64 : (BN <= 128 ? 128 : 256);  // TMEM columns per accumulator
  static constexpr int TMEM_COLS = 2 * ACC_STRIDE;                            // power of two >= 32
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + C_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int THREADS = 192;
};

template <int BN>
__global__ void __launch_bounds__(192, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
               const __grid_constant__ CUtensorMap tmRes, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
  uint8_t* smem_c = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_c + Cfg::C_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint64_t* cbuf_free = bars + 2 * STAGES + 4;   // staging buffer reusable (previous TMA stores have read it)
  uint64_t* res_full = bars + 2 * STAGES + 5;    // residual tile landed in the staging buffer
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 6);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

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
      mbar_init(&tmem_empty[i], 128);
    }
    mbar_init(cbuf_free, 1);
    mbar_init(res_full, 1);
    tma_prefetch_desc(&tmOut);
    tma_prefetch_desc(&tmRes);
    fence_mbar_init();
  } else if (warp == 1) {
    tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr_smem);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int m_tiles = (p.M + 127) / 128;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int tiles_per_z = m_tiles * n_tiles;
  const int total_tiles = tiles_per_z * p.Z;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int z = tile / tiles_per_z;
        const int t = tile - z * tiles_per_z;
        const int mt = t / n_tiles, nt = t - mt * n_tiles;
        const int m0 = mt * 128, n0 = nt * BN;
        // conv tile origin
        int cb = 0, cy = 0, cx = 0;
        if (p.a_mode == GEMM_A_CONV3X3) {
          const int hw = p.H * p.W;
          if (p.tile_b > 1) {
            cb = mt * p.tile_b;
          } else {
            cb = m0 / hw;
            const int rem = m0 - cb * hw;
            cy = rem / p.W;
            cx = rem - cy * p.W;
          }
        }
        const int bz1 = p.b_batched ? z % p.ZB1 : 0, bz2 = p.b_batched ? z / p.ZB1 : 0;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          const int tap = kb / p.kb_per_tap;
          const int r = kb - tap * p.kb_per_tap;
          const CUtensorMap* ma = (r < p.kb_split) ? &tmA0 : &tmA1;
          const int ka = (r < p.kb_split ? r : r - p.kb_split) * 64;
          void* sa = smem_a + stage * Cfg::A_BYTES;
          void* sb = smem_b + stage * Cfg::B_BYTES;
          if (p.a_mode == GEMM_A_CONV3X3) {
            const int ky = tap / 3, kx = tap - ky * 3;
            tma_load_4d(sa, ma, &full_bar[stage], ka, cx + kx - 1, cy + ky - 1, cb);
          } else {
            tma_load_4d(sa, ma, &full_bar[stage], ka, m0, z % p.ZA1, z / p.ZA1);
          }
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
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc_k = umma_idesc_f16(128, BN, false, false);
      const uint32_t idesc_mn = umma_idesc_f16(128, BN, false, true);
      const uint32_t idesc = p.b_mn_major ? idesc_mn : idesc_k;
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
        const int acc = iter & 1;
        const uint32_t acc_phase = (iter >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::ACC_STRIDE;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + stage * Cfg::A_BYTES);
          const uint32_t b_addr = smem_u32(smem_b + stage * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = umma_smem_desc(a_addr + k * 32, 16, 1024);
            const uint64_t db = p.b_mn_major ? umma_smem_desc(b_addr + k * 2048, 8192, 1024)
                                             : umma_smem_desc(b_addr + k * 32, 16, 1024);
            umma_f16_ss(d_tmem, da, db, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
          if (kb == p.num_kb - 1) umma_commit(&tmem_full[acc]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int row_in_tile = quad * 32 + lane;
    int iter = 0;
    if (p.epi_tma) {
      // ---- staged epilogue: TMEM -> regs -> (+bias/rowvec/residual, GEGLU) -> swizzled smem -> TMA store.
      // Output is produced in 64-column units through a double-buffered staging area (2 x [128 rows x 64 cols],
      // as 32-column atoms with 64B swizzle), so the TMA store of unit u overlaps the math of unit u+1.
      // The residual rows are prefetched straight from global memory into registers BEFORE waiting for the
      // accumulator, i.e. their latency hides behind the tile's main loop.
      constexpr int OUTW = BN;                    // accumulator columns; GEGLU emits OUTW/2 output columns
      constexpr int RES_CHUNKS = (BN + 7) / 8;    // 16-byte residual chunks per row
      const int outw = p.geglu ? OUTW / 2 : OUTW; // output columns per tile
      const int n_out_total = p.geglu ? p.N / 2 : p.N;
      const int units = (outw + 63) / 64;
      const bool leader = (warp == 2 && lane == 0);
      const uint32_t stg = smem_u32(smem_c);
      const uint32_t sw = (row_in_tile >> 1) & 3; // 64B swizzle: 16B-chunk index ^= (row / 2) % 4
      uint32_t unit = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
        const int z = tile / tiles_per_z;
        const int t = tile - z * tiles_per_z;
        const int mt = t / n_tiles, nt = t - mt * n_tiles;
        const int acc = iter & 1;
        const uint32_t acc_phase = (iter >> 1) & 1;
        const int row = mt * 128 + row_in_tile;
        const int n_out0 = nt * outw;
        // residual prefetch (global -> registers), issued while the MMAs of this tile are still running
        uint4 resv[RES_CHUNKS];
        if (p.res_tma) {
          const bool row_ok = row < p.M;
          const __half* rp = p.residual + static_cast<long long>(row) * p.ldr + n_out0;
#pragma unroll
          for (int i = 0; i < RES_CHUNKS; ++i) {
            resv[i] = make_uint4(0, 0, 0, 0);
            if (row_ok && n_out0 + i * 8 < n_out_total) resv[i] = __ldg(reinterpret_cast<const uint4*>(rp + i * 8));
          }
        }
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
        const uint32_t t_addr = tmem_base + acc * Cfg::ACC_STRIDE + (static_cast<uint32_t>(quad * 32) << 16);
        const int img = min(row, p.M - 1) / p.rows_per_img;
        const float* rv = (p.rowvec != nullptr) ? p.rowvec + static_cast<long long>(img) * p.ldv : nullptr;
#pragma unroll
        for (int u = 0; u < (OUTW + 63) / 64; ++u) {
          if (u < units) {
            const int unit_cols = min(64, outw - u * 64);
            const uint32_t buf = stg + (unit & 1) * 16384;
            // staging buffer (unit & 1) was last used by unit-2: its TMA store must have finished reading smem
            if (leader) bulk_wait_read1();
            named_bar_sync(1, 128);
#pragma unroll
            for (int h32 = 0; h32 < 2; ++h32) {
              const int c0 = h32 * 32;
              const int col_t = u * 64 + c0;         // output column within the tile
              const int ncol = n_out0 + col_t;       // global output column
              if (c0 < unit_cols && ncol < n_out_total) {   // warp-uniform
                float v[32];
                if (p.geglu) {
                  float gt[32];
                  tmem_ld32(t_addr + col_t, v);
                  tmem_ld32(t_addr + OUTW / 2 + col_t, gt);
                  tmem_ld_wait();
                  const float4* bh = reinterpret_cast<const float4*>(p.bias + nt * BN + col_t);
                  const float4* bg = reinterpret_cast<const float4*>(p.bias + nt * BN + OUTW / 2 + col_t);
#pragma unroll
                  for (int j4 = 0; j4 < 8; ++j4) {
                    float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
                    if (p.bias != nullptr) { b0 = __ldg(bh + j4); b1 = __ldg(bg + j4); }
                    const float hb[4] = {b0.x, b0.y, b0.z, b0.w}, gb[4] = {b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                      const int j = j4 * 4 + q;
                      v[j] = (v[j] * p.alpha + hb[q]) * gelu_erf(gt[j] * p.alpha + gb[q]);
                    }
                  }
                } else {
                  tmem_ld32(t_addr + col_t, v);
                  tmem_ld_wait();
                  if (ncol + 32 <= n_out_total && (p.bias == nullptr || (reinterpret_cast<uintptr_t>(p.bias + ncol) & 15) == 0) &&
                      (rv == nullptr || (reinterpret_cast<uintptr_t>(rv + ncol) & 15) == 0)) {
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                      float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
                      if (p.bias != nullptr) b0 = __ldg(reinterpret_cast<const float4*>(p.bias + ncol) + j4);
                      if (rv != nullptr) b1 = __ldg(reinterpret_cast<const float4*>(rv + ncol) + j4);
                      v[j4 * 4 + 0] = v[j4 * 4 + 0] * p.alpha + (b0.x + b1.x);
                      v[j4 * 4 + 1] = v[j4 * 4 + 1] * p.alpha + (b0.y + b1.y);
                      v[j4 * 4 + 2] = v[j4 * 4 + 2] * p.alpha + (b0.z + b1.z);
                      v[j4 * 4 + 3] = v[j4 * 4 + 3] * p.alpha + (b0.w + b1.w);
                    }
                  } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                      float x = v[j] * p.alpha;
                      if (ncol + j < n_out_total) {
                        if (p.bias != nullptr) x += __ldg(p.bias + ncol + j);
                        if (rv != nullptr) x += __ldg(rv + ncol + j);
                      }
                      v[j] = x;
                    }
                  }
                }
                const uint32_t atom = buf + h32 * 8192 + row_in_tile * 64;
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                  if (p.res_tma) {
                    const int ci = (col_t >> 3) + cc;
                    if (ci < RES_CHUNKS) {
                      const uint4 rr = resv[ci];
                      const uint32_t rw[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
                      for (int q = 0; q < 4; ++q) {
                        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&rw[q]));
                        v[cc * 8 + 2 * q] += f.x;
                        v[cc * 8 + 2 * q + 1] += f.y;
                      }
                    }
                  }
                  uint32_t o[4];
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
                    const __half2 h2 = __floats2half2_rn(v[cc * 8 + 2 * q], v[cc * 8 + 2 * q + 1]);
                    o[q] = *reinterpret_cast<const uint32_t*>(&h2);
                  }
                  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(atom + ((cc ^ sw) << 4)), "r"(o[0]),
                               "r"(o[1]), "r"(o[2]), "r"(o[3])
                               : "memory");
                }
              }
            }
            if (u == units - 1) {   // all TMEM reads of this accumulator are done
              tc_fence_before();
              mbar_arrive(&tmem_empty[acc]);
            }
            fence_proxy_async_smem();
            named_bar_sync(2, 128);
            if (leader) {
#pragma unroll
              for (int h32 = 0; h32 < 2; ++h32) {
                const int ncol = n_out0 + u * 64 + h32 * 32;
                if (h32 * 32 < unit_cols && ncol < n_out_total)
                  tma_store_4d(&tmOut, smem_c + (unit & 1) * 16384 + h32 * 8192, ncol, mt * 128, z % p.ZA1,
                               z / p.ZA1);
              }
              bulk_commit();
            }
            ++unit;
          }
        }
      }
      if (leader) bulk_wait0();
    } else
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
      const int z = tile / tiles_per_z;
      const int t = tile - z * tiles_per_z;
      const int mt = t / n_tiles, nt = t - mt * n_tiles;
      const int acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const long long out_zoff = (z % p.ZA1) * p.out_z1_stride + (z / p.ZA1) * p.out_z2_stride;
      const uint32_t t_addr = tmem_base + acc * Cfg::ACC_STRIDE + (static_cast<uint32_t>(quad * 32) << 16);
      const int row = mt * 128 + row_in_tile;
      const bool row_ok = row < p.M;
      const int img = row / p.rows_per_img;
      const int row_in_img = row - img * p.rows_per_img;
      const float* rv = (p.rowvec != nullptr && row_ok) ? p.rowvec + static_cast<long long>(img) * p.ldv : nullptr;
      const __half* res = (p.residual != nullptr && row_ok)
                              ? p.residual + z * p.res_zstride + static_cast<long long>(row) * p.ldr
                              : nullptr;

      constexpr int OUT_COLS = BN;  // accumulator columns of this tile
      if (p.geglu) {
        // columns [0, BN/2) = h, [BN/2, BN) = gate; output column = nt*BN/2 + j
        const int n_out0 = nt * (BN / 2);
        const int n_out_total = p.N / 2;
#pragma unroll 1
        for (int c0 = 0; c0 < BN / 2; c0 += 16) {
          float h[16], g[16];
          tmem_ld16(t_addr + c0, h);
          tmem_ld16(t_addr + BN / 2 + c0, g);
          tmem_ld_wait();
          if (row_ok) {
            __half o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int nb = nt * BN + c0 + j;  // index into (permuted) bias
              float hv = h[j] * p.alpha, gv = g[j] * p.alpha;
              if (p.bias != nullptr && n_out0 + c0 + j < n_out_total) {
                hv += p.bias[nb];
                gv += p.bias[nb + BN / 2];
              }
              o[j] = __float2half_rn(hv * gelu_erf(gv));
            }
            __half* op = reinterpret_cast<__half*>(p.out) + out_zoff + static_cast<long long>(row) * p.ldc +
                         n_out0 + c0;
            if (n_out0 + c0 + 16 <= n_out_total && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
              reinterpret_cast<uint4*>(op)[0] = reinterpret_cast<uint4*>(o)[0];
              reinterpret_cast<uint4*>(op)[1] = reinterpret_cast<uint4*>(o)[1];
            } else {
              for (int j = 0; j < 16; ++j)
                if (n_out0 + c0 + j < n_out_total) op[j] = o[j];
            }
          }
        }
      } else {
        const int n0 = nt * BN;
#pragma unroll 1
        for (int c0 = 0; c0 < OUT_COLS; c0 += 16) {
          if (n0 + c0 >= p.N) break;  // warp-uniform
          float v[16];
          tmem_ld16(t_addr + c0, v);
          tmem_ld_wait();
          if (row_ok) {
            const int nbase = n0 + c0;
            const bool full = nbase + 16 <= p.N;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float x = v[j] * p.alpha;
              if (full || nbase + j < p.N) {
                if (p.bias != nullptr) x += p.bias[nbase + j];
                if (rv != nullptr) x += rv[nbase + j];
              }
              v[j] = x;
            }
            if (res != nullptr) {
              if (full && ((reinterpret_cast<uintptr_t>(res + nbase) & 15) == 0)) {
                uint4 r0 = reinterpret_cast<const uint4*>(res + nbase)[0];
                uint4 r1 = reinterpret_cast<const uint4*>(res + nbase)[1];
                const __half* rh0 = reinterpret_cast<const __half*>(&r0);
                const __half* rh1 = reinterpret_cast<const __half*>(&r1);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  v[j] += __half2float(rh0[j]);
                  v[8 + j] += __half2float(rh1[j]);
                }
              } else {
                for (int j = 0; j < 16; ++j)
                  if (nbase + j < p.N) v[j] += __half2float(res[nbase + j]);
              }
            }
            if (p.out_mode == GEMM_OUT_ROWMAJOR) {
              if (p.out_fp32) {
                float* op = reinterpret_cast<float*>(p.out) + out_zoff + static_cast<long long>(row) * p.ldc +
                            nbase;
                if (full && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    reinterpret_cast<float4*>(op)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                } else {
                  for (int j = 0; j < 16; ++j)
                    if (nbase + j < p.N) op[j] = v[j];
                }
              } else {
                __half o[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) o[j] = __float2half_rn(v[j]);
                __half* op = reinterpret_cast<__half*>(p.out) + out_zoff +
                             static_cast<long long>(row) * p.ldc + nbase;
                if (full && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
                  reinterpret_cast<uint4*>(op)[0] = reinterpret_cast<uint4*>(o)[0];
                  reinterpret_cast<uint4*>(op)[1] = reinterpret_cast<uint4*>(o)[1];
                } else {
                  for (int j = 0; j < 16; ++j)
                    if (nbase + j < p.N) op[j] = o[j];
                }
              }
            } else {
              // transposed: element (row, n) -> out[z][img][n][row_in_img]; consecutive lanes = consecutive rows
              const long long base = out_zoff + static_cast<long long>(img) * p.out_imgstride + row_in_img;
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
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

}  // namespace icd
