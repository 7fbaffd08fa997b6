// HBM-bound normalisation kernels: GroupNorm(+SiLU) over channels-last activations, LayerNorm, row softmax.
// All loads/stores are 128-bit and coalesced along the channel axis; statistics are fp32 (GroupNorm partials
// are combined in fp64) and the reductions are deterministic (no float atomics to global memory).
#include <cooperative_groups.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>

#include "../../include/icd_b200.h"
#include "host_util.h"
#include "icd_ptx.cuh"

namespace icd {

#ifndef ICD_GN_THREADS
#define ICD_GN_THREADS 320
#endif
constexpr int GN_THREADS = ICD_GN_THREADS;   // multiple of every (C/8) <= 320 that occurs: 40, 80, 160, 320 (120/240 leave idle lanes)
// dynamic shared memory the single-pass kernel may use: 227 KB per SM minus its static arrays (s_part, s_mean, s_rstd)
constexpr int GN_FUSED_SMEM_MAX = 220 * 1024 - (GN_THREADS > 320 ? (GN_THREADS - 320) * 16 : 0);
constexpr int GN_MAX_CHUNKS = 64; // pixel chunks per image -> workspace = B * 64 * 2 * 32 floats

struct GnSrc {
  const __half* x0;
  const __half* x1;
  int C0, C1;
};

__device__ __forceinline__ const __half* gn_vec_ptr(const GnSrc& s, long long pix, int c) {
  // channel c (multiple of 8) of pixel `pix` in the virtual concat [x0 | x1]
  return c < s.C0 ? s.x0 + pix * s.C0 + c : s.x1 + pix * s.C1 + (c - s.C0);
}

// partial sums: ws[((b * chunks + chunk) * 2 + {0:sum,1:sumsq}) * 32 + g]
__global__ void __launch_bounds__(GN_THREADS) gn_stats_kernel(GnSrc src, int HW, int cpg, int chunks, float* ws) {
  pdl_launch_dependents();
  pdl_wait();
  const int C = src.C0 + src.C1;
  const int vpr = C >> 3;                      // 16-byte vectors per pixel
  const int rows_per_iter = GN_THREADS / vpr;  // >= 1 (C <= 2560)
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int v = threadIdx.x % vpr, r = threadIdx.x / vpr;
  const bool active = r < rows_per_iter;
  const int c = v * 8;
  const int g_lo = c / cpg;
  __shared__ float4 s_part[GN_THREADS];        // per-thread (sum_lo, sq_lo, sum_hi, sq_hi): fixed-order reduction
  const int rows_per_chunk = (HW + chunks - 1) / chunks;
  const int p0 = chunk * rows_per_chunk;
  const int p1 = min(HW, p0 + rows_per_chunk);
  float sl = 0.f, ql = 0.f, sh = 0.f, qh = 0.f;
  const int split = (g_lo + 1) * cpg - c;  // elements j < split belong to g_lo, the rest to g_lo + 1
  if (active) {
    int pix = p0 + r;
    // four pixels per iteration: four independent 16-byte loads in flight per thread
    for (; pix + 3 * rows_per_iter < p1; pix += 4 * rows_per_iter) {
      uint4 raw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        raw[u] = *reinterpret_cast<const uint4*>(
            gn_vec_ptr(src, static_cast<long long>(b) * HW + pix + u * rows_per_iter, c));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const __half* h = reinterpret_cast<const __half*>(&raw[u]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float x = __half2float(h[j]);
          if (j < split) {
            sl += x;
            ql += x * x;
          } else {
            sh += x;
            qh += x * x;
          }
        }
      }
    }
    for (; pix < p1; pix += rows_per_iter) {
      const uint4 raw = *reinterpret_cast<const uint4*>(gn_vec_ptr(src, static_cast<long long>(b) * HW + pix, c));
      const __half* h = reinterpret_cast<const __half*>(&raw);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float x = __half2float(h[j]);
        if (j < split) {
          sl += x;
          ql += x * x;
        } else {
          sh += x;
          qh += x * x;
        }
      }
    }
  }
  s_part[threadIdx.x] = make_float4(sl, ql, sh, qh);
  __syncthreads();
  if (threadIdx.x < 32) {
    // group g gathers, in a fixed order, the "lo" parts of vectors starting inside it and the "hi" parts of the
    // vectors that straddle into it (deterministic: no floating-point atomics anywhere)
    const int g = threadIdx.x;
    const int v_first = max(0, (g * cpg - 7 + 7) / 8 - 1);
    const int v_last = min(vpr - 1, ((g + 1) * cpg - 1) / 8);
    float s = 0.f, q = 0.f;
    for (int vv = v_first; vv <= v_last; ++vv) {
      const int glo = (vv * 8) / cpg;
      const int sp = (glo + 1) * cpg - vv * 8;
      for (int rr = 0; rr < rows_per_iter; ++rr) {
        const float4 pt = s_part[rr * vpr + vv];
        if (glo == g) {
          s += pt.x;
          q += pt.y;
        } else if (glo + 1 == g && sp < 8) {
          s += pt.z;
          q += pt.w;
        }
      }
    }
    float* o = ws + (static_cast<long long>(b) * chunks + chunk) * 64;
    o[g] = s;
    o[32 + g] = q;
  }
}

__global__ void __launch_bounds__(GN_THREADS)
gn_apply_kernel(GnSrc src, __half* __restrict__ y, int HW, int cpg, int chunks, int apply_chunks, float eps,
                const float* __restrict__ gamma, const float* __restrict__ beta, int do_silu,
                const float* __restrict__ ws, double inv_n) {
  const int C = src.C0 + src.C1;
  const int vpr = C >> 3;
  const int rows_per_iter = GN_THREADS / vpr;
  const int b = blockIdx.y, chunk = blockIdx.x;
  __shared__ float s_mean[32], s_rstd[32];
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x < 256) {
    // finalize the statistics: 8 lanes per group sum the per-chunk partials (strided, fixed order) in fp64, then a
    // 3-step shuffle tree combines them — deterministic, and ~10x shorter than one thread walking all chunks
    const int g = threadIdx.x >> 3, l = threadIdx.x & 7;
    double s = 0.0, q = 0.0;
    const float* w = ws + static_cast<long long>(b) * chunks * 64;
    float sv[8], qv[8];                       // all partials requested up front: one L2 round trip (see gn_fused_kernel)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = l + 8 * k;
      const bool ok = i < chunks;
      sv[k] = ok ? w[i * 64 + g] : 0.f;
      qv[k] = ok ? w[i * 64 + 32 + g] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      s += static_cast<double>(sv[k]);
      q += static_cast<double>(qv[k]);
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (l == 0) {
      // E[x^2] - mean^2 stays in fp64 (cancellation); the reciprocal of the count comes from the host and the
      // inverse square root runs on the fp32 MUFU: fp64 division and square root are long software sequences that
      // sat on the critical path of every CTA right behind the grid barrier (tools/gn_prof.py)
      const double mean = s * inv_n;
      double var = q * inv_n - mean * mean;
      if (var < 0.0) var = 0.0;
      s_mean[g] = static_cast<float>(mean);
      s_rstd[g] = rsqrtf(static_cast<float>(var) + eps);
    }
  }
  __syncthreads();
  const int v = threadIdx.x % vpr, r = threadIdx.x / vpr;
  if (r >= rows_per_iter) return;
  const int c = v * 8;
  float a[8], sft[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (c + j) / cpg;
    a[j] = s_rstd[g] * gamma[c + j];
    sft[j] = beta[c + j] - s_mean[g] * a[j];
  }
  const int rows_per_chunk = (HW + apply_chunks - 1) / apply_chunks;
  const int p0 = chunk * rows_per_chunk;
  const int p1 = min(HW, p0 + rows_per_chunk);
  auto emit = [&](const uint4& raw, long long gp) {
    const __half* h = reinterpret_cast<const __half*>(&raw);
    uint4 outv;
    __half* o = reinterpret_cast<__half*>(&outv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float x = __half2float(h[j]) * a[j] + sft[j];
      if (do_silu) x = silu(x);
      o[j] = __float2half_rn(x);
    }
    *reinterpret_cast<uint4*>(y + gp * C + c) = outv;
  };
  int pix = p0 + r;
  for (; pix + 3 * rows_per_iter < p1; pix += 4 * rows_per_iter) {   // four independent 16-byte loads in flight
    const long long g0 = static_cast<long long>(b) * HW + pix;
    const long long g1 = g0 + rows_per_iter, g2 = g1 + rows_per_iter, g3 = g2 + rows_per_iter;
    const uint4 r0 = *reinterpret_cast<const uint4*>(gn_vec_ptr(src, g0, c));
    const uint4 r1 = *reinterpret_cast<const uint4*>(gn_vec_ptr(src, g1, c));
    const uint4 r2 = *reinterpret_cast<const uint4*>(gn_vec_ptr(src, g2, c));
    const uint4 r3 = *reinterpret_cast<const uint4*>(gn_vec_ptr(src, g3, c));
    emit(r0, g0);
    emit(r1, g1);
    emit(r2, g2);
    emit(r3, g3);
  }
  for (; pix < p1; pix += rows_per_iter) {
    const long long gp = static_cast<long long>(b) * HW + pix;
    emit(*reinterpret_cast<const uint4*>(gn_vec_ptr(src, gp, c)), gp);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Single-pass GroupNorm(+SiLU): one read of x, one write of y, one launch.
// Each CTA keeps its pixel chunk of one image in shared memory between the statistics pass and the normalisation
// pass; the per-chunk partial sums are exchanged through the caller's workspace across ONE grid-wide barrier. The
// kernel is launched COOPERATIVELY (launch_k_coop): the driver guarantees that all CTAs are resident together or
// refuses the launch, and the barrier is cooperative_groups' grid sync, whose state belongs to the launch — no
// library-global counters, so launches on different streams (or from different CUDA graphs) cannot interfere, and
// there is nothing to reset between CUDA-graph replays. The host picks this path only when chunks * B fits
// (occupancy calculator, gn_fused_plan); otherwise the two-kernel path above runs. pdl_wait() precedes every
// global access.
#ifdef ICD_GN_PROFILE
// debug builds only (make VARIANT=gnprof EXTRA_DEFS=-DICD_GN_PROFILE, tools/gn_prof.py): clock64 stamps of thread 0 of
// the first 256 CTAs of the LAST gn_fused launch: [cta][8] = start | after pdl_wait | loads issued+stored | partials
// published | grid barrier passed | statistics final | written out
__device__ long long* g_gn_prof = nullptr;
#define ICD_GN_STAMP(k)                                                                                   \
  if (threadIdx.x == 0 && g_gn_prof != nullptr && blockIdx.y * gridDim.x + blockIdx.x < 256)               \
    g_gn_prof[(blockIdx.y * gridDim.x + blockIdx.x) * 8 + (k)] = clock64();
#else
#define ICD_GN_STAMP(k)
#endif

__global__ void __launch_bounds__(GN_THREADS)
gn_fused_kernel(GnSrc src, __half* __restrict__ y, int HW, int cpg, int chunks, float eps,
                const float* __restrict__ gamma, const float* __restrict__ beta, int do_silu, float* ws, double inv_n) {
  extern __shared__ uint4 s_x[];               // [rows of this chunk][vpr] 16-byte vectors
  __shared__ float4 s_part[GN_THREADS];
  __shared__ float s_mean[32], s_rstd[32];
  const int C = src.C0 + src.C1;
  const int vpr = C >> 3;
  const int rows_per_iter = GN_THREADS / vpr;
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int v = threadIdx.x % vpr, r = threadIdx.x / vpr;
  const bool active = r < rows_per_iter;
  const int c = v * 8;
  const int g_lo = c / cpg;
  const int split = (g_lo + 1) * cpg - c;      // elements j < split belong to g_lo, the rest to g_lo + 1
  const int rows_per_chunk = (HW + chunks - 1) / chunks;
  const int p0 = chunk * rows_per_chunk;
  const int p1 = min(HW, p0 + rows_per_chunk);
  ICD_GN_STAMP(0)
  pdl_wait();
  ICD_GN_STAMP(1)
  // ---- pass 1: global -> shared, partial sums (each thread only ever touches its own smem slots)
  float sl = 0.f, ql = 0.f, sh = 0.f, qh = 0.f;
  auto accum = [&](const uint4& raw) {
    const __half* h = reinterpret_cast<const __half*>(&raw);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float x = __half2float(h[j]);
      if (j < split) {
        sl += x;
        ql += x * x;
      } else {
        sh += x;
        qh += x * x;
      }
    }
  };
  if (active) {
    int pix = p0 + r;
    // eight independent 16-byte loads in flight per thread: with ONE CTA (320 threads) per SM the read phase is bound
    // by bytes in flight (Little's law), not by bandwidth; the accumulation order is unchanged (row order)
    for (; pix + 7 * rows_per_iter < p1; pix += 8 * rows_per_iter) {
      uint4 raw[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        raw[u] = *reinterpret_cast<const uint4*>(
            gn_vec_ptr(src, static_cast<long long>(b) * HW + pix + u * rows_per_iter, c));
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        s_x[(pix - p0 + u * rows_per_iter) * vpr + v] = raw[u];
        accum(raw[u]);
      }
    }
    for (; pix + 3 * rows_per_iter < p1; pix += 4 * rows_per_iter) {
      uint4 raw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        raw[u] = *reinterpret_cast<const uint4*>(
            gn_vec_ptr(src, static_cast<long long>(b) * HW + pix + u * rows_per_iter, c));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        s_x[(pix - p0 + u * rows_per_iter) * vpr + v] = raw[u];
        accum(raw[u]);
      }
    }
    for (; pix < p1; pix += rows_per_iter) {
      const uint4 raw = *reinterpret_cast<const uint4*>(gn_vec_ptr(src, static_cast<long long>(b) * HW + pix, c));
      s_x[(pix - p0) * vpr + v] = raw;
      accum(raw);
    }
  }
  s_part[threadIdx.x] = make_float4(sl, ql, sh, qh);
  __syncthreads();
  ICD_GN_STAMP(2)
  if (threadIdx.x < 32) {
    // same fixed-order gather as gn_stats_kernel (deterministic: no floating-point atomics anywhere)
    const int g = threadIdx.x;
    const int v_first = max(0, (g * cpg - 7 + 7) / 8 - 1);
    const int v_last = min(vpr - 1, ((g + 1) * cpg - 1) / 8);
    float s = 0.f, q = 0.f;
    for (int vv = v_first; vv <= v_last; ++vv) {
      const int glo = (vv * 8) / cpg;
      const int sp = (glo + 1) * cpg - vv * 8;
      for (int rr = 0; rr < rows_per_iter; ++rr) {
        const float4 pt = s_part[rr * vpr + vv];
        if (glo == g) {
          s += pt.x;
          q += pt.y;
        } else if (glo + 1 == g && sp < 8) {
          s += pt.z;
          q += pt.w;
        }
      }
    }
    float* o = ws + (static_cast<long long>(b) * chunks + chunk) * 64;
    __stcg(o + g, s);
    __stcg(o + 32 + g, q);
  }
  // ---- publish, then wait until every chunk has published: grid-wide barrier of the cooperative launch
  __threadfence();
  ICD_GN_STAMP(3)
  cooperative_groups::this_grid().sync();
  ICD_GN_STAMP(4)
  pdl_launch_dependents();
  // ---- finalize the statistics (every CTA of the image does this redundantly: 8 lanes per group, fp64, fixed order)
  if (threadIdx.x < 256) {
    const int g = threadIdx.x >> 3, l = threadIdx.x & 7;
    double s = 0.0, q = 0.0;
    const float* w = ws + static_cast<long long>(b) * chunks * 64;
    // all (<= GN_MAX_CHUNKS / 8 = 8 per lane) partials are requested before the first one is consumed: ONE L2 round
    // trip instead of a chain of them (tools/gn_prof.py: this phase was 7.0k of the kernel's 28.7k cycles); same summation order
    static_assert(GN_MAX_CHUNKS <= 64, "finalize holds 8 partials per lane");
    float sv[8], qv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = l + 8 * k;
      const bool ok = i < chunks;
      sv[k] = ok ? __ldcg(w + i * 64 + g) : 0.f;
      qv[k] = ok ? __ldcg(w + i * 64 + 32 + g) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      s += static_cast<double>(sv[k]);
      q += static_cast<double>(qv[k]);
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (l == 0) {
      // E[x^2] - mean^2 stays in fp64 (cancellation); the reciprocal of the count comes from the host and the
      // inverse square root runs on the fp32 MUFU: fp64 division and square root are long software sequences that
      // sat on the critical path of every CTA right behind the grid barrier (tools/gn_prof.py)
      const double mean = s * inv_n;
      double var = q * inv_n - mean * mean;
      if (var < 0.0) var = 0.0;
      s_mean[g] = static_cast<float>(mean);
      s_rstd[g] = rsqrtf(static_cast<float>(var) + eps);
    }
  }
  __syncthreads();
  ICD_GN_STAMP(5)
  // (the partials in `ws` are overwritten by the NEXT GroupNorm launch only: stream order / pdl_wait() there)
  if (!active) return;
  // ---- pass 2: shared -> normalise (+SiLU) -> global
  float a[8], sft[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int g = (c + j) / cpg;
    a[j] = s_rstd[g] * gamma[c + j];
    sft[j] = beta[c + j] - s_mean[g] * a[j];
  }
  auto emit = [&](int pix) {
    const uint4 raw = s_x[(pix - p0) * vpr + v];
    const __half* h = reinterpret_cast<const __half*>(&raw);
    uint4 outv;
    __half* o = reinterpret_cast<__half*>(&outv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float x = __half2float(h[j]) * a[j] + sft[j];
      if (do_silu) x = silu(x);
      o[j] = __float2half_rn(x);
    }
    *reinterpret_cast<uint4*>(y + (static_cast<long long>(b) * HW + pix) * C + c) = outv;
  };
  {
    // four rows per iteration: 32 independent cvt -> fma -> tanh -> fma -> cvt chains cover the MUFU latency that a
    // single CTA per SM (10 warps) cannot hide by switching warps
    int pix = p0 + r;
    for (; pix + 3 * rows_per_iter < p1; pix += 4 * rows_per_iter) {
#pragma unroll
      for (int u = 0; u < 4; ++u) emit(pix + u * rows_per_iter);
    }
    for (; pix < p1; pix += rows_per_iter) emit(pix);
  }
  ICD_GN_STAMP(6)
}

// one warp per row; values stay in registers between the mean and variance passes
template <int MAXV>  // max 16-byte vectors per lane
__global__ void __launch_bounds__(256) layernorm_kernel(const __half* __restrict__ x, __half* __restrict__ y, int rows,
                                                        int C, float eps, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta) {
  pdl_launch_dependents();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int nvec = C >> 3;
  const __half* xr = x + static_cast<long long>(warp) * C;
  float vals[MAXV][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      const uint4 raw = *reinterpret_cast<const uint4*>(xr + v * 8);
      const __half* h = reinterpret_cast<const __half*>(&raw);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        vals[i][j] = __half2float(h[j]);
        sum += vals[i][j];
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = vals[i][j] - mean;
        sq += d * d;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / C + eps);
  __half* yr = y + static_cast<long long>(warp) * C;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = lane + i * 32;
    if (v < nvec) {
      const float4 g0 = *reinterpret_cast<const float4*>(gamma + v * 8);
      const float4 g1 = *reinterpret_cast<const float4*>(gamma + v * 8 + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(beta + v * 8);
      const float4 b1 = *reinterpret_cast<const float4*>(beta + v * 8 + 4);
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      uint4 outv;
      __half* o = reinterpret_cast<__half*>(&outv);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = __float2half_rn((vals[i][j] - mean) * rstd * g[j] + bb[j]);
      *reinterpret_cast<uint4*>(yr + v * 8) = outv;
    }
  }
}

// one warp per row, in place; three passes (row is L1/L2 resident). causal_period > 0: row r is query r % period of
// its sequence and attends to keys 0..(r % period) only (CLIP text transformer); the masked tail is written as 0.
__global__ void __launch_bounds__(256) softmax_kernel(__half* __restrict__ x, long long rows, int cols, long long ld,
                                                      int causal_period) {
  pdl_launch_dependents();
  pdl_wait();
  const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  __half* xr = x + warp * ld;
  const int all_cols = cols;
  if (causal_period > 0) cols = min(cols, static_cast<int>(warp % causal_period) + 1);
  float m = -INFINITY;
  for (int c = lane; c < cols; c += 32) m = fmaxf(m, __half2float(xr[c]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) s += __expf(__half2float(xr[c]) - m);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float inv = 1.f / s;
  for (int c = lane; c < cols; c += 32) xr[c] = __float2half_rn(__expf(__half2float(xr[c]) - m) * inv);
  for (int c = cols + lane; c < all_cols; c += 32) xr[c] = __float2half_rn(0.f);
}

}  // namespace icd

using namespace icd;

// Chunking of the single-pass GroupNorm: the largest chunk count (<= 64 per image) whose CTAs all fit on the device
// at once with their pixel chunk in shared memory. Returns false when the tensor is too large to stay on chip.
static bool gn_fused_plan(int B, int HW, int C, int* chunks_out, size_t* smem_out) {
  static const bool enabled = [] { const char* e = getenv("ICD_GN_FUSED"); return e == nullptr || atoi(e) != 0; }();
  if (!enabled || B > 1024 || B < 1 || HW < 1) return false;
  static PerDeviceFlag configured;
  if (!configured.cur()) {
    if (cudaFuncSetAttribute(gn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GN_FUSED_SMEM_MAX) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    configured.cur() = true;
  }
  const int max_by_rows = (HW * (C / 8) + 4 * GN_THREADS - 1) / (4 * GN_THREADS);
  // CTAs per SM to aim for: more, smaller CTAs keep more loads in flight; fewer, fatter ones make the grid barrier and
  // the (per-CTA, redundant) statistics finalize cheaper. ICD_GN_CPS overrides for the sweep in tools/gn_bench.py.
  // Measured (profiles/r2_c_gn_cps_sweep.txt): ONE CTA per SM wins on every shape of both networks (8x64^2x320:
  // 20.0 -> 13.7 us, 8x32^2x640: 16.9 -> 10.2 us) — at 4 per SM the barrier took 7.3k and the finalize 7.1k cycles of a
  // 29.8k-cycle kernel (512 CTAs each re-reading all partials of their image from L2), at 1 per SM 3.0k and 1.7k.
  // Denser plans are tried only when the sparser one does not exist (more images than SMs, or a chunk too large for
  // one CTA's shared memory).
  static const int cps_first = [] { const char* e = getenv("ICD_GN_CPS"); const int v = e ? atoi(e) : 1; return v < 1 ? 1 : (v > 4 ? 4 : v); }();
  for (int cps = cps_first; cps <= 4; ++cps) {
    int fc = (sm_count() * cps) / B;
    if (fc > GN_MAX_CHUNKS) fc = GN_MAX_CHUNKS;
    if (fc > max_by_rows) fc = max_by_rows;
    if (fc > HW) fc = HW;
    if (fc < 1) continue;
    const int rows = (HW + fc - 1) / fc;
    const size_t smem = static_cast<size_t>(rows) * C * 2;
    if (smem > static_cast<size_t>(GN_FUSED_SMEM_MAX)) continue;
    int granted = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&granted, gn_fused_kernel, GN_THREADS, smem) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    if (granted < 1 || static_cast<long long>(fc) * B > static_cast<long long>(granted) * sm_count()) continue;
    *chunks_out = fc;
    *smem_out = smem;
    return true;
  }
  return false;
}

#ifdef ICD_GN_PROFILE
extern "C" long long* icd_gn_prof_buffer(void) {
  static long long* buf = nullptr;
  if (buf == nullptr) {
    cudaMallocManaged(&buf, 256 * 8 * sizeof(long long));
    cudaMemcpyToSymbol(g_gn_prof, &buf, sizeof(buf));
  }
  return buf;
}
#endif

extern "C" int icd_groupnorm_launches(int B, int HW, int C) {
  int fc = 0;
  size_t smem = 0;
  return gn_fused_plan(B, HW, C, &fc, &smem) ? 1 : 2;
}

extern "C" int icd_groupnorm(const void* x0, int C0, const void* x1, int C1, void* y, int B, int HW, int groups,
                             float eps, const float* gamma, const float* beta, int apply_silu, float* stats_ws,
                             void* stream) {
  const int C = C0 + (x1 != nullptr ? C1 : 0);
  if (groups != 32) return set_error("icd_groupnorm: only 32 groups supported");
  if (C % 32 != 0 || C0 % 8 != 0 || (x1 != nullptr && C1 % 8 != 0)) return set_error("icd_groupnorm: bad channels");
  if (C / 8 > GN_THREADS) return set_error("icd_groupnorm: C > 2560 unsupported");
  const int cpg = C / 32;
  // a 16-byte vector (8 channels) may straddle at most two groups: cpg >= 8, or cpg == 4 (VAE: 128 channels)
  if (cpg < 8 && cpg != 4) return set_error("icd_groupnorm: channels per group must be 4 or >= 8");
  GnSrc src{reinterpret_cast<const __half*>(x0), reinterpret_cast<const __half*>(x1), C0, x1 != nullptr ? C1 : 0};
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int max_by_rows = (HW * (C / 8) + 4 * GN_THREADS - 1) / (4 * GN_THREADS);
  const double inv_n = 1.0 / (static_cast<double>(HW) * cpg);
  // ---- single-pass path: every CTA of an image must be resident at once (see gn_fused_kernel)
  int fc = 0;
  size_t fsmem = 0;
  if (gn_fused_plan(B, HW, C, &fc, &fsmem)) {
    const cudaError_t e = launch_k_coop(gn_fused_kernel, dim3(fc, B), dim3(GN_THREADS), fsmem, st, src,
                                        reinterpret_cast<__half*>(y), HW, cpg, fc, eps, gamma, beta, apply_silu, stats_ws, inv_n);
    if (e == cudaSuccess) return check_launch("gn_fused");
    if (e != cudaErrorCooperativeLaunchTooLarge) return set_error(std::string("gn_fused launch: ") + cudaGetErrorString(e));
    cudaGetLastError();   // the driver could not co-schedule the grid (e.g. a partitioned device): two-kernel path
  }
  // stats chunks index the caller's workspace ([B][chunks][2][32] floats, contract: B * 4096 floats): never more
  // than GN_MAX_CHUNKS per image
  // CTAs per SM of the two-kernel path (ICD_GN2_CPS = "stats,apply"; sweep in profiles/r2_c_gn_cps_sweep.txt)
  static const int cps_stats = [] { const char* e = getenv("ICD_GN2_CPS"); const int v = e ? atoi(e) : 4; return v < 1 ? 1 : (v > 8 ? 8 : v); }();
  static const int cps_apply = [] {
    const char* e = getenv("ICD_GN2_CPS");
    const char* c = e ? strchr(e, ',') : nullptr;
    const int v = c ? atoi(c + 1) : 4;
    return v < 1 ? 1 : (v > 8 ? 8 : v);
  }();
  int chunks = (cps_stats * sm_count() + B - 1) / B;
  if (chunks > max_by_rows) chunks = max_by_rows;
  if (chunks > GN_MAX_CHUNKS) chunks = GN_MAX_CHUNKS;
  if (chunks < 1) chunks = 1;
  launch_k(gn_stats_kernel, dim3(chunks, B), dim3(GN_THREADS), 0, st, src, HW, cpg, chunks, stats_ws);
  if (check_launch("gn_stats")) return 1;
  int apply_chunks = (cps_apply * sm_count() + B - 1) / B;
  if (apply_chunks > max_by_rows) apply_chunks = max_by_rows;
  if (apply_chunks > HW) apply_chunks = HW;
  if (apply_chunks < 1) apply_chunks = 1;
  launch_k(gn_apply_kernel, dim3(apply_chunks, B), dim3(GN_THREADS), 0, st, src, reinterpret_cast<__half*>(y), HW, cpg, chunks,
                                                                apply_chunks, eps, gamma, beta, apply_silu, stats_ws, inv_n);
  return check_launch("gn_apply");
}

extern "C" int icd_layernorm(const void* x, void* y, int rows, int C, float eps, const float* gamma,
                             const float* beta, void* stream) {
  if (C % 8 != 0) return set_error("icd_layernorm: C must be a multiple of 8");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int warps_per_block = 8;
  const int grid = (rows + warps_per_block - 1) / warps_per_block;
  const int nvec = C / 8;
  const __half* xp = reinterpret_cast<const __half*>(x);
  __half* yp = reinterpret_cast<__half*>(y);
  if (nvec <= 64)
    launch_k(layernorm_kernel<2>, dim3(grid), dim3(256), 0, st, xp, yp, rows, C, eps, gamma, beta);
  else if (nvec <= 160)
    launch_k(layernorm_kernel<5>, dim3(grid), dim3(256), 0, st, xp, yp, rows, C, eps, gamma, beta);
  else if (nvec <= 320)
    launch_k(layernorm_kernel<10>, dim3(grid), dim3(256), 0, st, xp, yp, rows, C, eps, gamma, beta);
  else
    return set_error("icd_layernorm: C > 2560 unsupported");
  return check_launch("layernorm");
}

extern "C" int icd_softmax_causal(void* x, long long rows, int cols, long long ld, int causal_period, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long grid = (rows + 7) / 8;
  launch_k(softmax_kernel, dim3(static_cast<unsigned>(grid)), dim3(256), 0, st, reinterpret_cast<__half*>(x), rows, cols, ld,
           causal_period);
  return check_launch("softmax");
}

extern "C" int icd_softmax(void* x, long long rows, int cols, long long ld, void* stream) {
  return icd_softmax_causal(x, rows, cols, ld, 0, stream);
}
