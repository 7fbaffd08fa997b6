// HBM-bound layout / embedding / update kernels of the iCD step (all 128-bit vectorised where the layout allows).
#include <cuda_fp16.h>

#include "../../include/icd_b200.h"
#include "host_util.h"
#include "icd_ptx.cuh"

namespace icd {

// y[b][2h+dy][2w+dx][:] = x[b][h][w][:]   (Upsample2D: F.interpolate(scale_factor=2, mode="nearest"))
__global__ void __launch_bounds__(256) upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int B,
                                                         int H, int W, int vpc /*vectors per pixel*/) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = static_cast<long long>(B) * H * W * vpc;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % vpc);
    long long pix = i / vpc;
    const int w = static_cast<int>(pix % W);
    pix /= W;
    const int h = static_cast<int>(pix % H);
    const int b = static_cast<int>(pix / H);
    const uint4 val = x[i];
    const long long W2 = 2LL * W;
    const long long o = ((static_cast<long long>(b) * 2 * H + 2 * h) * W2 + 2 * w) * vpc + v;
    y[o] = val;
    y[o + vpc] = val;
    y[o + W2 * vpc] = val;
    y[o + W2 * vpc + vpc] = val;
  }
}

// y[(b,ho,wo)][tap][c] = x[b][2ho+ky-pad][2wo+kx-pad][c] (zero outside): operand of the stride-2 Downsample2D conv.
// pad = 1: the U-Net's Downsample2D (conv padding 1); pad = 0: the VAE encoder's (F.pad(x, (0,1,0,1)) then an
// unpadded conv, i.e. zeros only on the bottom / right edge).
__global__ void __launch_bounds__(256) im2col_s2_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int B, int H,
                                                        int W, int vpc, int pad) {
  pdl_launch_dependents();
  pdl_wait();
  const int Ho = H / 2, Wo = W / 2;
  const long long total = static_cast<long long>(B) * Ho * Wo * 9 * vpc;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int v = static_cast<int>(i % vpc);
    long long r = i / vpc;
    const int tap = static_cast<int>(r % 9);
    r /= 9;
    const int wo = static_cast<int>(r % Wo);
    r /= Wo;
    const int ho = static_cast<int>(r % Ho);
    const int b = static_cast<int>(r / Ho);
    const int hy = 2 * ho + tap / 3 - pad, wx = 2 * wo + tap % 3 - pad;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (hy >= 0 && hy < H && wx >= 0 && wx < W) val = x[((static_cast<long long>(b) * H + hy) * W + wx) * vpc + v];
    y[i] = val;
  }
}

// NCHW fp32 -> NHWC fp16, channels zero-padded to Cpad
__global__ void __launch_bounds__(256) latent_to_nhwc_kernel(const float* __restrict__ x, __half* __restrict__ y, int B,
                                                             int C, int HW, int Cpad) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = static_cast<long long>(B) * HW;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int p = static_cast<int>(i % HW);
    const int b = static_cast<int>(i / HW);
    for (int c = 0; c < Cpad; ++c) {
      const float v = c < C ? x[(static_cast<long long>(b) * C + c) * HW + p] : 0.f;
      y[i * Cpad + c] = __float2half_rn(v);
    }
  }
}

// y[r][:] = [cos(t_r f_i) | sin(t_r f_i)]: diffusers Timesteps(flip_sin_to_cos=True); freqs computed by the host
// with the reference's fp32 op order so the arguments are bit-identical.
__global__ void timestep_embedding_kernel(const float* __restrict__ t, const float* __restrict__ freqs,
                                          __half* __restrict__ y, int n, int half_dim) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * half_dim) return;
  const int r = i / half_dim, k = i - r * half_dim;
  const float arg = t[r] * freqs[k];
  y[static_cast<long long>(r) * 2 * half_dim + k] = __float2half_rn(cosf(arg));
  y[static_cast<long long>(r) * 2 * half_dim + half_dim + k] = __float2half_rn(sinf(arg));
}

// guidance_scale_embedding (utils/generation.py:96-122): emb = (1000 w) * f_i ; y = [sin | cos]
__global__ void guidance_embedding_kernel(const float* __restrict__ w, const float* __restrict__ freqs,
                                          __half* __restrict__ y, int n, int half_dim) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * half_dim) return;
  const int r = i / half_dim, k = i - r * half_dim;
  const float arg = (w[r] * 1000.0f) * freqs[k];
  y[static_cast<long long>(r) * 2 * half_dim + k] = __float2half_rn(sinf(arg));
  y[static_cast<long long>(r) * 2 * half_dim + half_dim + k] = __float2half_rn(cosf(arg));
}

__global__ void __launch_bounds__(256) silu_kernel(const __half* __restrict__ x, __half* __restrict__ y, long long n) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = __half2float(x[i]);
    y[i] = __float2half_rn(v / (1.0f + expf(-v)));
  }
}

// y = act(x): 0 SiLU, 1 quick_gelu (x * sigmoid(1.702 x), CLIP-L), 2 exact-erf GELU (OpenCLIP bigG)
__global__ void __launch_bounds__(256) act_kernel(const __half2* __restrict__ x, __half2* __restrict__ y, long long n2,
                                                  int kind) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n2;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float2 v = __half22float2(x[i]);
    float2 r;
    if (kind == 0) {
      r = make_float2(v.x / (1.0f + expf(-v.x)), v.y / (1.0f + expf(-v.y)));
    } else if (kind == 1) {
      r = make_float2(v.x / (1.0f + expf(-1.702f * v.x)), v.y / (1.0f + expf(-1.702f * v.y)));
    } else {
      r = make_float2(gelu_erf(v.x), gelu_erf(v.y));
    }
    y[i] = __floats2half2_rn(r.x, r.y);
  }
}

// out[b*T + t][:] = tok[ids[b*T + t]][:] + pos[t][:]   (CLIPTextEmbeddings), 16-byte vectors
__global__ void __launch_bounds__(256) embed_tokens_kernel(const long long* __restrict__ ids, const uint4* __restrict__ tok,
                                                           const uint4* __restrict__ pos, uint4* __restrict__ out,
                                                           long long rows, int T, int vpr, int vocab) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = rows * vpr;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / vpr;
    const int v = static_cast<int>(i - r * vpr);
    long long id = ids[r];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    const uint4 a = tok[id * vpr + v], b = pos[static_cast<long long>(r % T) * vpr + v];
    uint4 o;
    const __half2* ah = reinterpret_cast<const __half2*>(&a);
    const __half2* bh = reinterpret_cast<const __half2*>(&b);
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 fa = __half22float2(ah[j]), fb = __half22float2(bh[j]);
      oh[j] = __floats2half2_rn(fa.x + fb.x, fa.y + fb.y);
    }
    out[i] = o;
  }
}

__global__ void __launch_bounds__(256) add_kernel(const __half* __restrict__ a, const __half* __restrict__ b,
                                                  __half* __restrict__ y, long long n) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    y[i] = __float2half_rn(__half2float(a[i]) + __half2float(b[i]));
}

// predicted_origin (utils/generation.py:136-155), epsilon prediction, same fp32 operation order as the reference
__global__ void __launch_bounds__(256)
consistency_update_kernel(const float* __restrict__ eps, const float* __restrict__ x, float* __restrict__ out,
                          long long per_sample, int B, const float* __restrict__ alpha_t,
                          const float* __restrict__ sigma_t, const float* __restrict__ alpha_s,
                          const float* __restrict__ sigma_s) {
  pdl_launch_dependents();
  pdl_wait();
  const long long total = per_sample * B;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / per_sample);
    const float e = eps[i];
    const float x0 = (x[i] - sigma_t[b] * e) / alpha_t[b];
    out[i] = alpha_s[b] * x0 + sigma_s[b] * e;
  }
}

static inline unsigned grid_for(long long n, int threads = 256) {
  long long g = (n + threads - 1) / threads;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<unsigned>(g);
}

}  // namespace icd

using namespace icd;

extern "C" int icd_upsample2x(const void* x, void* y, int B, int H, int W, int C, void* stream) {
  if (C % 8 != 0) return set_error("icd_upsample2x: C must be a multiple of 8");
  const int vpc = C / 8;
  launch_k(upsample2x_kernel, dim3(grid_for(static_cast<long long>(B) * H * W * vpc)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<const uint4*>(x),
                                                                reinterpret_cast<uint4*>(y), B, H, W, vpc);
  return check_launch("upsample2x");
}

extern "C" int icd_im2col_s2_pad(const void* x, void* y, int B, int H, int W, int C, int pad, void* stream) {
  if (C % 8 != 0 || (H & 1) || (W & 1)) return set_error("icd_im2col_s2: C % 8 != 0 or odd H/W");
  if (pad != 0 && pad != 1) return set_error("icd_im2col_s2: pad must be 0 or 1");
  const int vpc = C / 8;
  launch_k(im2col_s2_kernel, dim3(grid_for(static_cast<long long>(B) * (H / 2) * (W / 2) * 9 * vpc)), dim3(256), 0,
           reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), B, H, W,
           vpc, pad);
  return check_launch("im2col_s2");
}

extern "C" int icd_im2col_s2(const void* x, void* y, int B, int H, int W, int C, void* stream) {
  return icd_im2col_s2_pad(x, y, B, H, W, C, 1, stream);
}

extern "C" int icd_latent_to_nhwc(const float* x, void* y, int B, int C, int HW, int Cpad, void* stream) {
  launch_k(latent_to_nhwc_kernel, dim3(grid_for(static_cast<long long>(B) * HW)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      x, reinterpret_cast<__half*>(y), B, C, HW, Cpad);
  return check_launch("latent_to_nhwc");
}

extern "C" int icd_timestep_embedding(const float* t, const float* freqs, void* y, int n, int dim, void* stream) {
  if (dim & 1) return set_error("icd_timestep_embedding: odd dim");
  const int total = n * (dim / 2);
  launch_k(timestep_embedding_kernel, dim3((total + 127) / 128), dim3(128), 0, reinterpret_cast<cudaStream_t>(stream), 
      t, freqs, reinterpret_cast<__half*>(y), n, dim / 2);
  return check_launch("timestep_embedding");
}

extern "C" int icd_guidance_embedding(const float* w, const float* freqs, void* y, int n, int dim, void* stream) {
  if (dim & 1) return set_error("icd_guidance_embedding: odd dim");
  const int total = n * (dim / 2);
  launch_k(guidance_embedding_kernel, dim3((total + 127) / 128), dim3(128), 0, reinterpret_cast<cudaStream_t>(stream), 
      w, freqs, reinterpret_cast<__half*>(y), n, dim / 2);
  return check_launch("guidance_embedding");
}

extern "C" int icd_silu(const void* x, void* y, long long n, void* stream) {
  launch_k(silu_kernel, dim3(grid_for(n)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), reinterpret_cast<const __half*>(x),
                                                                              reinterpret_cast<__half*>(y), n);
  return check_launch("silu");
}

extern "C" int icd_act(const void* x, void* y, long long n, int kind, void* stream) {
  if ((n & 1) || kind < 0 || kind > 2) return set_error("icd_act: odd element count or unknown activation");
  launch_k(act_kernel, dim3(grid_for(n / 2)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream),
           reinterpret_cast<const __half2*>(x), reinterpret_cast<__half2*>(y), n / 2, kind);
  return check_launch("act");
}

extern "C" int icd_embed_tokens(const long long* ids, const void* tok, const void* pos, void* out, long long rows, int T,
                                int C, int vocab, void* stream) {
  if (C % 8 != 0 || T < 1) return set_error("icd_embed_tokens: C must be a multiple of 8");
  launch_k(embed_tokens_kernel, dim3(grid_for(rows * (C / 8))), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), ids,
           reinterpret_cast<const uint4*>(tok), reinterpret_cast<const uint4*>(pos), reinterpret_cast<uint4*>(out), rows, T,
           C / 8, vocab);
  return check_launch("embed_tokens");
}

extern "C" int icd_add(const void* a, const void* b, void* y, long long n, void* stream) {
  launch_k(add_kernel, dim3(grid_for(n)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      reinterpret_cast<const __half*>(a), reinterpret_cast<const __half*>(b), reinterpret_cast<__half*>(y), n);
  return check_launch("add");
}

extern "C" int icd_consistency_update(const float* eps, const float* x, float* out, long long per_sample, int B,
                                      const float* alpha_t, const float* sigma_t, const float* alpha_s,
                                      const float* sigma_s, void* stream) {
  launch_k(consistency_update_kernel, dim3(grid_for(per_sample * B)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), 
      eps, x, out, per_sample, B, alpha_t, sigma_t, alpha_s, sigma_s);
  return check_launch("consistency_update");
}
