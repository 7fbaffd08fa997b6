// fp32 validation path (load_models(dtype='fp32'): the reference runs SD1.5 *editing* in fp32,
// running/sd1.5/launch_editing_iCD_sd1.5.sh:38, utils/loading.py:38-41).
//
// Same executor, same packed layouts, same epilogue semantics as the fp16 tensor-core path, but every operand and
// every activation is fp32 and the contractions run on the FMA pipe with fp32 accumulation — no tensor-core operand
// rounding, no fp16 activation storage. It exists so that the hot path can be checked element-wise against the fp32
// CPU oracle (rtol 1e-3 / atol 1e-4 holds with two orders of magnitude to spare) and so that `dtype='fp32'` means
// what it means in the reference. It is a correctness mode (SGEMM 15-25 TFLOP/s; one SD1.5 row-forward 111 ms vs 8.8 ms in fp16, tools/f32_speed.py).
#include <cuda_runtime.h>

#include <cmath>
#include <string>

#include "../../include/icd_b200.h"
#include "host_util.h"
#include "icd_ptx.cuh"

namespace icd {

// ------------------------------------------------------------------------------------------------ SGEMM
//   out[z][m][n] = alpha * sum_k A[z][m][k] * Bm[z][n][k]  (+ bias[n]) (+ rowvec[m / rows_per_img][n]) (+ residual[m][n])
// A is [a0 | a1] along K; with CONV the K axis is (filter tap, channel) of a 3x3 / pad 1 / stride 1 convolution over
// NHWC images gathered on the fly. BKN: the B operand is stored [K][N] (P.V with V rows as they lie in memory).
constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 8, SG_PAD = 4;

template <bool CONV, bool BKN>
__global__ void __launch_bounds__(256) sgemm_f32_kernel(const IcdSgemm p) {
  __shared__ __align__(16) float As[SG_BK][SG_BM + SG_PAD];
  __shared__ __align__(16) float Bs[SG_BK][SG_BN + SG_PAD];
  pdl_launch_dependents();
  pdl_wait();
  const int t = threadIdx.x;
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  const int z = blockIdx.z, zb = z / p.ZH, zh = z - zb * p.ZH;
  const float* a0 = p.a0 + zb * p.a_zb + zh * p.a_zh;
  const float* a1 = p.a1;
  const float* bm = p.b + zb * p.b_zb + zh * p.b_zh;
  float* out = p.out + zb * p.c_zb + zh * p.c_zh;
  const int Ct = p.C0 + p.C1;
  const int K = CONV ? 9 * Ct : p.K;
  const bool vec = p.vec != 0;

  // A loader: one tile row and 4 consecutive k per thread
  const int a_row = t >> 1, a_kq = (t & 1) * 4;
  const int am = m0 + a_row;
  int ab = 0, ay = 0, ax = 0;
  if (CONV && am < p.M) {
    const int hw = p.H * p.W;
    ab = am / hw;
    const int rem = am - ab * hw;
    ay = rem / p.W;
    ax = rem - ay * p.W;
  }
  // B loader: [N][K] -> same shape as the A loader; [K][N] -> one k row, 4 consecutive n per thread
  const int b_row = BKN ? (t >> 5) : (t >> 1), b_q = BKN ? (t & 31) * 4 : (t & 1) * 4;

  const int tx = t & 15, ty = t >> 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  auto load_a_elem = [&](int k) -> float {
    if (am >= p.M || k >= K) return 0.f;
    if (CONV) {
      const int tap = k / Ct, c = k - tap * Ct;
      const int yy = ay + tap / 3 - 1, xx = ax + tap % 3 - 1;
      if (yy < 0 || yy >= p.H || xx < 0 || xx >= p.W) return 0.f;
      const long long pix = (static_cast<long long>(ab) * p.H + yy) * p.W + xx;
      return c < p.C0 ? a0[pix * p.a0_ld + c] : a1[pix * p.a1_ld + (c - p.C0)];
    }
    return k < p.C0 ? a0[static_cast<long long>(am) * p.a0_ld + k] : a1[static_cast<long long>(am) * p.a1_ld + (k - p.C0)];
  };
  auto load_b_elem = [&](int n, int k) -> float {
    if (n >= p.N || k >= K) return 0.f;
    if (BKN) return bm[static_cast<long long>(k) * p.b_ld + n];
    if (CONV) {
      const int tap = k / Ct, c = k - tap * Ct;
      return bm[static_cast<long long>(n) * p.b_ld + tap * p.w_tap_ld + c];
    }
    return bm[static_cast<long long>(n) * p.b_ld + k];
  };

  // operands of one K block -> registers (the block after the one being multiplied: its latency hides under the FMAs)
  auto fetch = [&](int k0, float4& av, float4& bv) {
    av = make_float4(0.f, 0.f, 0.f, 0.f);
    bv = make_float4(0.f, 0.f, 0.f, 0.f);
    {
      const int k = k0 + a_kq;
      if (vec && am < p.M && k + 3 < K) {
        if (CONV) {
          const int tap = k / Ct, c = k - tap * Ct;
          const int yy = ay + tap / 3 - 1, xx = ax + tap % 3 - 1;
          if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W) {
            const long long pix = (static_cast<long long>(ab) * p.H + yy) * p.W + xx;
            av = c < p.C0 ? *reinterpret_cast<const float4*>(a0 + pix * p.a0_ld + c)
                          : *reinterpret_cast<const float4*>(a1 + pix * p.a1_ld + (c - p.C0));
          }
        } else {
          av = k < p.C0 ? *reinterpret_cast<const float4*>(a0 + static_cast<long long>(am) * p.a0_ld + k)
                        : *reinterpret_cast<const float4*>(a1 + static_cast<long long>(am) * p.a1_ld + (k - p.C0));
        }
      } else {
        av = make_float4(load_a_elem(k), load_a_elem(k + 1), load_a_elem(k + 2), load_a_elem(k + 3));
      }
    }
    if (BKN) {
      const int k = k0 + b_row, n = n0 + b_q;
      if (vec && k < K && n + 3 < p.N) {
        bv = *reinterpret_cast<const float4*>(bm + static_cast<long long>(k) * p.b_ld + n);
      } else {
        bv = make_float4(load_b_elem(n, k), load_b_elem(n + 1, k), load_b_elem(n + 2, k), load_b_elem(n + 3, k));
      }
    } else {
      const int n = n0 + b_row, k = k0 + b_q;
      if (vec && n < p.N && k + 3 < K) {
        if (CONV) {
          const int tap = k / Ct, c = k - tap * Ct;
          bv = *reinterpret_cast<const float4*>(bm + static_cast<long long>(n) * p.b_ld + tap * p.w_tap_ld + c);
        } else {
          bv = *reinterpret_cast<const float4*>(bm + static_cast<long long>(n) * p.b_ld + k);
        }
      } else {
        bv = make_float4(load_b_elem(n, k), load_b_elem(n, k + 1), load_b_elem(n, k + 2), load_b_elem(n, k + 3));
      }
    }
  };
  float4 av, bv;
  fetch(0, av, bv);
  for (int k0 = 0; k0 < K; k0 += SG_BK) {
    __syncthreads();   // previous tile consumed
    As[a_kq + 0][a_row] = av.x;
    As[a_kq + 1][a_row] = av.y;
    As[a_kq + 2][a_row] = av.z;
    As[a_kq + 3][a_row] = av.w;
    if (BKN) {
      *reinterpret_cast<float4*>(&Bs[b_row][b_q]) = bv;
    } else {
      Bs[b_q + 0][b_row] = bv.x;
      Bs[b_q + 1][b_row] = bv.y;
      Bs[b_q + 2][b_row] = bv.z;
      Bs[b_q + 3][b_row] = bv.w;
    }
    __syncthreads();
    if (k0 + SG_BK < K) fetch(k0 + SG_BK, av, bv);
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      const float4 a_lo = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 a_hi = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      const float4 b_lo = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float4 b_hi = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
      const float a[8] = {a_lo.x, a_lo.y, a_lo.z, a_lo.w, a_hi.x, a_hi.y, a_hi.z, a_hi.w};
      const float b[8] = {b_lo.x, b_lo.y, b_lo.z, b_lo.w, b_hi.x, b_hi.y, b_hi.z, b_hi.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= p.M) continue;
    const float* rv = p.rowvec != nullptr ? p.rowvec + static_cast<long long>(m / p.rows_per_img) * p.ldv : nullptr;
    const float* res = p.residual != nullptr ? p.residual + zb * p.r_zb + zh * p.r_zh + static_cast<long long>(m) * p.ldr
                                             : nullptr;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= p.N) continue;
      float v = acc[i][j] * p.alpha;
      if (p.bias != nullptr) v += p.bias[n];
      if (rv != nullptr) v += rv[n];
      if (res != nullptr) v += res[n];
      out[static_cast<long long>(m) * p.ldc + n] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------ normalisations
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float s = 0.f;
  for (int i = 0; i < nw; ++i) s += red[i];   // same order in every thread: deterministic
  return s;
}

// GroupNorm (+ SiLU) over NHWC fp32, one CTA per (image, group); two-pass statistics (mean, then centred variance).
__global__ void __launch_bounds__(512) groupnorm_f32_kernel(const float* __restrict__ x0, int C0, const float* __restrict__ x1,
                                                            int C1, float* __restrict__ y, int HW, int groups, float eps,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            int silu_on) {
  __shared__ float red[16];
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x / groups, g = blockIdx.x - b * groups;
  const int C = C0 + C1, cpg = C / groups;
  const long long n = static_cast<long long>(HW) * cpg;
  auto at = [&](long long e) -> float {
    const long long pix = e / cpg;
    const int c = g * cpg + static_cast<int>(e - pix * cpg);
    const long long row = static_cast<long long>(b) * HW + pix;
    return c < C0 ? x0[row * C0 + c] : x1[row * C1 + (c - C0)];
  };
  float s = 0.f;
  for (long long e = threadIdx.x; e < n; e += blockDim.x) s += at(e);
  const float mean = block_sum(s, red) / static_cast<float>(n);
  float q = 0.f;
  for (long long e = threadIdx.x; e < n; e += blockDim.x) {
    const float d = at(e) - mean;
    q = fmaf(d, d, q);
  }
  const float rstd = rsqrtf(block_sum(q, red) / static_cast<float>(n) + eps);
  for (long long e = threadIdx.x; e < n; e += blockDim.x) {
    const long long pix = e / cpg;
    const int c = g * cpg + static_cast<int>(e - pix * cpg);
    float v = (at(e) - mean) * rstd * gamma[c] + beta[c];
    if (silu_on) v = v / (1.f + expf(-v));
    y[(static_cast<long long>(b) * HW + pix) * C + c] = v;
  }
}

// LayerNorm over the last dim, one warp per row, two-pass statistics.
__global__ void __launch_bounds__(256) layernorm_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int rows, int C,
                                                            float eps, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + static_cast<long long>(row) * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += xr[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / C;
  float q = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float d = xr[c] - mean;
    q = fmaf(d, d, q);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / C + eps);
  float* yr = y + static_cast<long long>(row) * C;
  for (int c = lane; c < C; c += 32) yr[c] = (xr[c] - mean) * rstd * gamma[c] + beta[c];
}

// In-place softmax over the first `cols` entries of rows with stride ld; one warp per row; pad entries become 0.
__global__ void __launch_bounds__(256) softmax_f32_kernel(float* __restrict__ x, long long rows, int cols, long long ld) {
  pdl_launch_dependents();
  pdl_wait();
  const long long row = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* xr = x + row * ld;
  float m = -INFINITY;
  for (int c = lane; c < cols; c += 32) m = fmaxf(m, xr[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) {
    const float e = expf(xr[c] - m);
    xr[c] = e;
    s += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float inv = 1.f / s;
  for (int c = lane; c < cols; c += 32) xr[c] *= inv;
  for (long long c = cols + lane; c < ld; c += 32) xr[c] = 0.f;
}

// ------------------------------------------------------------------------------------------------ element-wise / layout
__global__ void __launch_bounds__(256) silu_f32_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = x[i];
    y[i] = v / (1.f + expf(-v));
  }
}

// GEGLU on a projection whose rows were interleaved per bn-wide tile as [bn/2 hidden | bn/2 gate] (packing.pack_geglu):
// y[m][j] = h * gelu_erf(g)
__global__ void __launch_bounds__(256) geglu_f32_kernel(const float* __restrict__ x, float* __restrict__ y, long long M, int F,
                                                        int bn) {
  pdl_launch_dependents();
  pdl_wait();
  const int half = bn / 2;
  const long long total = M * F;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long m = i / F;
    const int j = static_cast<int>(i - m * F);
    const int tile = j / half, w = j - tile * half;
    const float* row = x + m * 2 * F + static_cast<long long>(tile) * bn;
    const float h = row[w], g = row[half + w];
    y[i] = h * (0.5f * g * (1.0f + erff(g * 0.70710678118654752440f)));
  }
}

// nearest 2x upsampling, NHWC
__global__ void __launch_bounds__(256) upsample2x_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H,
                                                             int W, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = static_cast<long long>(B) * 4 * H * W * C;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    long long r = i / C;
    const int xo = static_cast<int>(r % (2 * W));
    r /= 2 * W;
    const int yo = static_cast<int>(r % (2 * H));
    const int b = static_cast<int>(r / (2 * H));
    y[i] = x[((static_cast<long long>(b) * H + yo / 2) * W + xo / 2) * C + c];
  }
}

// operand of a stride-2 3x3 / pad 1 convolution: y[(b, yo, xo)][(tap, c)]
__global__ void __launch_bounds__(256) im2col_s2_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int H,
                                                            int W, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int Ho = H / 2, Wo = W / 2;
  const long long total = static_cast<long long>(B) * Ho * Wo * 9 * C;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    long long r = i / C;
    const int tap = static_cast<int>(r % 9);
    r /= 9;
    const int xo = static_cast<int>(r % Wo);
    r /= Wo;
    const int yo = static_cast<int>(r % Ho);
    const int b = static_cast<int>(r / Ho);
    const int yy = 2 * yo + tap / 3 - 1, xx = 2 * xo + tap % 3 - 1;
    y[i] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? x[((static_cast<long long>(b) * H + yy) * W + xx) * C + c] : 0.f;
  }
}

// NCHW -> NHWC with the channel dim zero-padded to Cpad, and back (C leading channels of an NHWC matrix with row
// stride ld)
__global__ void __launch_bounds__(256) nchw_to_nhwc_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int C,
                                                               int HW, int Cpad) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = static_cast<long long>(B) * HW * Cpad;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % Cpad);
    const long long r = i / Cpad;
    const long long b = r / HW, pix = r - b * HW;
    y[i] = c < C ? x[(b * C + c) * HW + pix] : 0.f;
  }
}
__global__ void __launch_bounds__(256) nhwc_to_nchw_f32_kernel(const float* __restrict__ x, long long ld, float* __restrict__ y,
                                                               int B, int C, int HW) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = static_cast<long long>(B) * C * HW;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long pix = i % HW;
    const long long r = i / HW;
    const int c = static_cast<int>(r % C);
    const long long b = r / C;
    y[i] = x[(b * HW + pix) * ld + c];
  }
}

// y[r] = sin_first ? [sin(a) | cos(a)] : [cos(a) | sin(a)],  a = (v[r] * scale) * freqs[k]
// (scale 1, cos first: diffusers Timesteps(flip_sin_to_cos=True); scale 1000, sin first: utils/generation.py:96-122)
__global__ void sincos_embedding_f32_kernel(const float* __restrict__ v, const float* __restrict__ freqs, float* __restrict__ y,
                                            int n, int half_dim, float scale, int sin_first) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * half_dim) return;
  const int r = i / half_dim, k = i - r * half_dim;
  const float arg = (v[r] * scale) * freqs[k];
  const float s = sinf(arg), c = cosf(arg);
  y[static_cast<long long>(r) * 2 * half_dim + k] = sin_first ? s : c;
  y[static_cast<long long>(r) * 2 * half_dim + half_dim + k] = sin_first ? c : s;
}

static inline int grid_for_n(long long n) {
  long long g = (n + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace icd

using namespace icd;

extern "C" int icd_sgemm_f32(const IcdSgemm* g, void* stream) {
  if (g == nullptr || g->a0 == nullptr || g->b == nullptr || g->out == nullptr) return set_error("icd_sgemm_f32: null operand");
  if (g->M <= 0 || g->N <= 0 || g->Z <= 0 || g->ZH <= 0) return set_error("icd_sgemm_f32: empty problem");
  if (g->conv && g->b_kn) return set_error("icd_sgemm_f32: conv with a [K][N] weight operand is not supported");
  if (g->C1 > 0 && g->a1 == nullptr) return set_error("icd_sgemm_f32: C1 > 0 without a1");
  if (g->rowvec != nullptr && g->rows_per_img <= 0) return set_error("icd_sgemm_f32: rowvec needs rows_per_img");
  IcdSgemm p = *g;
  if (!p.conv) p.K = p.C0 + p.C1;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  // vector loads: every 4-element group the loaders touch must be 16-byte aligned and inside one source / one tap
  bool vec = (p.C0 % 4 == 0) && (p.C1 % 4 == 0) && (p.a0_ld % 4 == 0) && (p.C1 == 0 || p.a1_ld % 4 == 0) &&
             (p.b_ld % 4 == 0) && al16(p.a0) && (p.a1 == nullptr || al16(p.a1)) && al16(p.b) &&
             (p.a_zb % 4 == 0) && (p.a_zh % 4 == 0) && (p.b_zb % 4 == 0) && (p.b_zh % 4 == 0);
  if (p.conv) vec = vec && (p.w_tap_ld % 4 == 0);
  p.vec = vec ? 1 : 0;
  dim3 grid((p.N + SG_BN - 1) / SG_BN, (p.M + SG_BM - 1) / SG_BM, p.Z);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (p.conv) launch_k(sgemm_f32_kernel<true, false>, grid, dim3(256), 0, st, p);
  else if (p.b_kn) launch_k(sgemm_f32_kernel<false, true>, grid, dim3(256), 0, st, p);
  else launch_k(sgemm_f32_kernel<false, false>, grid, dim3(256), 0, st, p);
  return check_launch("sgemm_f32");
}

extern "C" int icd_groupnorm_f32(const float* x0, int C0, const float* x1, int C1, float* y, int B, int HW, int groups,
                                 float eps, const float* gamma, const float* beta, int silu, void* stream) {
  if (x0 == nullptr || y == nullptr || gamma == nullptr || beta == nullptr) return set_error("icd_groupnorm_f32: null operand");
  if (groups <= 0 || (C0 + C1) % groups != 0) return set_error("icd_groupnorm_f32: channels not divisible by groups");
  launch_k(groupnorm_f32_kernel, dim3(B * groups), dim3(512), 0, reinterpret_cast<cudaStream_t>(stream), x0, C0, x1, C1, y, HW,
           groups, eps, gamma, beta, silu);
  return check_launch("groupnorm_f32");
}

extern "C" int icd_layernorm_f32(const float* x, float* y, int rows, int C, float eps, const float* gamma, const float* beta,
                                 void* stream) {
  launch_k(layernorm_f32_kernel, dim3((rows + 7) / 8), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), x, y, rows, C, eps,
           gamma, beta);
  return check_launch("layernorm_f32");
}

extern "C" int icd_softmax_f32(float* x, long long rows, int cols, long long ld, void* stream) {
  if (cols <= 0 || ld < cols) return set_error("icd_softmax_f32: bad row length");
  launch_k(softmax_f32_kernel, dim3(static_cast<unsigned>((rows + 7) / 8)), dim3(256), 0,
           reinterpret_cast<cudaStream_t>(stream), x, rows, cols, ld);
  return check_launch("softmax_f32");
}

extern "C" int icd_silu_f32(const float* x, float* y, long long n, void* stream) {
  launch_k(silu_f32_kernel, dim3(grid_for_n(n)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), x, y, n);
  return check_launch("silu_f32");
}

extern "C" int icd_geglu_f32(const float* x, float* y, long long M, int F, int bn, void* stream) {
  if (bn <= 0 || (bn & 1) || F % (bn / 2) != 0) return set_error("icd_geglu_f32: width not a multiple of bn/2");
  launch_k(geglu_f32_kernel, dim3(grid_for_n(M * F)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), x, y, M, F, bn);
  return check_launch("geglu_f32");
}

extern "C" int icd_upsample2x_f32(const float* x, float* y, int B, int H, int W, int C, void* stream) {
  launch_k(upsample2x_f32_kernel, dim3(grid_for_n(static_cast<long long>(B) * 4 * H * W * C)), dim3(256), 0,
           reinterpret_cast<cudaStream_t>(stream), x, y, B, H, W, C);
  return check_launch("upsample2x_f32");
}

extern "C" int icd_im2col_s2_f32(const float* x, float* y, int B, int H, int W, int C, void* stream) {
  launch_k(im2col_s2_f32_kernel, dim3(grid_for_n(static_cast<long long>(B) * (H / 2) * (W / 2) * 9 * C)), dim3(256), 0,
           reinterpret_cast<cudaStream_t>(stream), x, y, B, H, W, C);
  return check_launch("im2col_s2_f32");
}

extern "C" int icd_nchw_to_nhwc_f32(const float* x, float* y, int B, int C, int HW, int Cpad, void* stream) {
  if (Cpad < C) return set_error("icd_nchw_to_nhwc_f32: Cpad < C");
  launch_k(nchw_to_nhwc_f32_kernel, dim3(grid_for_n(static_cast<long long>(B) * HW * Cpad)), dim3(256), 0,
           reinterpret_cast<cudaStream_t>(stream), x, y, B, C, HW, Cpad);
  return check_launch("nchw_to_nhwc_f32");
}

extern "C" int icd_nhwc_to_nchw_f32(const float* x, long long ld, float* y, int B, int C, int HW, void* stream) {
  launch_k(nhwc_to_nchw_f32_kernel, dim3(grid_for_n(static_cast<long long>(B) * C * HW)), dim3(256), 0,
           reinterpret_cast<cudaStream_t>(stream), x, ld, y, B, C, HW);
  return check_launch("nhwc_to_nchw_f32");
}

extern "C" int icd_sincos_embedding_f32(const float* v, const float* freqs, float* y, int n, int dim, float scale,
                                        int sin_first, void* stream) {
  if (dim & 1) return set_error("icd_sincos_embedding_f32: odd dim");
  const int total = n * (dim / 2);
  launch_k(sincos_embedding_f32_kernel, dim3((total + 127) / 128), dim3(128), 0, reinterpret_cast<cudaStream_t>(stream), v,
           freqs, y, n, dim / 2, scale, sin_first);
  return check_launch("sincos_embedding_f32");
}
