// Host side of the tcgen05 GEMM: TMA descriptor construction (cached), tile-width heuristic, launch.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>

#include "../../include/icd_b200.h"
#include "gemm_tc.cuh"
#include "host_util.h"

namespace icd {

// ------------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency).
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess || p == nullptr) {
      cudaGetLastError();
      return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

struct TmapKey {
  uint64_t v[13];
  bool operator==(const TmapKey& o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 1469598103934665603ull;
    for (uint64_t x : k.v) {
      h ^= x;
      h *= 1099511628211ull;
    }
    return static_cast<size_t>(h);
  }
};
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
static std::mutex g_tmap_mu;

// fp16, rank-4, 128B (or 64B) swizzle. dims/box in elements, strides (dims 1..3) in bytes.
int make_tmap_4d(CUtensorMap* out, const void* ptr, const uint64_t dims[4], const uint64_t strides_b[3],
                 const uint32_t box[4], int swizzle_bytes, int elem_bytes) {
  TmapKey key;
  key.v[0] = reinterpret_cast<uint64_t>(ptr);
  for (int i = 0; i < 4; ++i) key.v[1 + i] = dims[i];
  for (int i = 0; i < 3; ++i) key.v[5 + i] = strides_b[i];
  for (int i = 0; i < 4; ++i) key.v[8 + i] = box[i];
  key.v[12] = static_cast<uint64_t>(swizzle_bytes) | (static_cast<uint64_t>(elem_bytes) << 16);
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  if ((reinterpret_cast<uint64_t>(ptr) & 15) != 0) return set_error("TMA operand pointer must be 16-byte aligned");
  for (int i = 0; i < 3; ++i)
    if ((strides_b[i] & 15) != 0) return set_error("TMA operand strides must be multiples of 16 bytes");
  cuuint64_t gd[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t gs[3] = {strides_b[0], strides_b[1], strides_b[2]};
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t es[4] = {1, 1, 1, 1};
  alignas(64) CUtensorMap m;
  CUresult r = fn(&m, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(ptr), gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[512];
    snprintf(buf, sizeof(buf),
             "cuTensorMapEncodeTiled failed (%d): dims=(%llu,%llu,%llu,%llu) strides=(%llu,%llu,%llu) "
             "box=(%u,%u,%u,%u)",
             static_cast<int>(r), (unsigned long long)dims[0], (unsigned long long)dims[1],
             (unsigned long long)dims[2], (unsigned long long)dims[3], (unsigned long long)strides_b[0],
             (unsigned long long)strides_b[1], (unsigned long long)strides_b[2], box[0], box[1], box[2], box[3]);
    return set_error(buf);
  }
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    if (g_tmap_cache.size() > 65536) g_tmap_cache.clear();
    g_tmap_cache.emplace(key, m);
  }
  *out = m;
  return 0;
}

// Tile-shape heuristic. Every operand byte is fetched from L2 by each CTA that needs it, and a 128x160 tile at
// full tensor rate would need ~31 TB/s of L2->SM traffic (measured cap ~12 TB/s, profiles/r1_c_*), so the model
// charges each K block max(MMA time, L2 time) and prefers 256-row tiles (two MMA halves share one B tile).
// gemm2sm_tc.cu
bool gemm2sm_wanted(int M, int N, int K, bool geglu);
int launch_gemm2sm(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& o, int M, int N, int num_kb,
                   const float* bias, bool geglu, cudaStream_t st);

struct TileChoice { int bm, bn, splits; };
static TileChoice pick_tile(int M, int N, int Z, int num_kb, int geglu, int b_mn_major, int force_bn, int force_bm,
                            int max_splits = 1) {
  const int sms = sm_count();
  // Measured override (tools/gemm_bench.py #15, #18-#20; profiles/r1i_gemm_qkv_tile_sweep.txt): the Q|K|V projections
  // (fat N, K = C <= 1280). The model below prefers 256-row tiles for them, but 256x256 and 256x160 tiles fill the
  // TMEM (single-buffered accumulator), so after every short main loop the epilogue and the burst of output writes
  // that all CTAs issue together are exposed; 128x256 with double-buffered accumulators is 8-32 % faster.
  if (force_bn == 0 && force_bm == 0 && !geglu && !b_mn_major && Z == 1 && N >= 1536 && M >= 2048 && num_kb < 48)
    return TileChoice{128, 256, 1};
  const int bns[4] = {256, 160, 128, 64};
  const int bms[2] = {128, 256};
  TileChoice best{128, 128, 1};
  double best_cost = 1e30;
  for (int bi = 0; bi < 2; ++bi) {
    const int bm = bms[bi];
    if (force_bm != 0 && bm != force_bm) continue;
    if (force_bm == 0 && bm == 256 && (geglu || b_mn_major || M <= 128)) continue;
    for (int i = 0; i < 4; ++i) {
      const int bn = bns[i];
      if (force_bn != 0 && bn != force_bn) continue;
      if (force_bn == 0) {
        if ((b_mn_major || geglu) && bn == 160) continue;
        if (geglu && ((N % bn) != 0 || bn < 128)) continue;
        if (bn > 64 && N <= bn / 2) continue;  // mostly-empty tile
      }
      const long long tiles = static_cast<long long>((M + bm - 1) / bm) * ((N + bn - 1) / bn) * Z;
      // 4 K-steps per block; back-to-back MMAs into the SAME accumulator are latency-chained (~157 cycles each,
      // measured), so a 128-row tile never beats ~630 cycles per K block; 256-row tiles interleave two chains
      const double chain = 157.0 / (bm / 128);
      const double mma = 4.0 * (bm / 128) * ((bn / 2.0) > chain ? (bn / 2.0) : chain);
      const double l2 = 3.0 * (bm + bn);               // (bm+bn) * 128 B / ~42.5 B/clk/SM
      const double epi = (bm / 128) * (600.0 + 6.0 * bn) * (geglu ? 3.0 : 1.0);
      const int half_stride = bn <= 64 ? 64 : (bn <= 128 ? 128 : 256);
      const bool dbl = 2 * (bm / 128) * half_stride <= 512;
      for (int sp = 1; sp <= max_splits; ++sp) {
        if (sp > 1 && (tiles * sp > sms || num_kb / sp < 8)) break;   // split-K only to fill idle SMs
        const int kbs = (num_kb + sp - 1) / sp;
        const long long waves = (tiles * sp + sms - 1) / sms;
        const double mainloop = kbs * (mma > l2 ? mma : l2) + 1500.0;
        const double per_tile = dbl ? (mainloop > epi ? mainloop : epi) : mainloop + epi;
        double cost = waves * per_tile + (dbl ? epi : 0.0);
        // fp32 partials written + re-read by the reduce kernel (~3400 B/clk of HBM), plus its launch
        if (sp > 1) cost += 5000.0 + static_cast<double>(sp + 1) * M * N * 4.0 / 3400.0;
        if (cost < best_cost) {
          best_cost = cost;
          best = TileChoice{bm, bn, sp};
        }
      }
    }
  }
  return best;
}

template <int BM, int BN, int EPI>
static int launch(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const CUtensorMap& o,
                  const CUtensorMap& r, const GemmParams& p, cudaStream_t st) {
  using Cfg = GemmCfg<BM, BN, EPI>;
  static PerDeviceFlag configured;
  if (!configured.cur()) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BM, BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e));
    configured.cur() = true;
  }
  const long long tiles = static_cast<long long>((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN) * p.Z * p.splits;
  int grid = static_cast<int>(tiles < sm_count() ? tiles : sm_count());
  if (grid < 1) grid = 1;
  launch_k(gemm_tc_kernel<BM, BN, EPI>, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, st, a0, a1, b, o, r, p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(std::string("gemm_tc launch: ") + cudaGetErrorString(e));
  return 0;
}

// split-K pass 2: out = fp16( alpha * sum_s ws[s] + bias + rowvec[img] + residual ), 8 columns per thread
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ ws, int splits, int M, int N, float alpha, const float* __restrict__ bias,
                     const float* __restrict__ rowvec, int rows_per_img, int ldv, const __half* __restrict__ res,
                     long long ldr, __half* __restrict__ out, long long ldc) {
  pdl_launch_dependents();
  pdl_wait();
  const int nv = N >> 3;
  const long long total = static_cast<long long>(M) * nv;
  const long long mn = static_cast<long long>(M) * N;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int m = static_cast<int>(i / nv);
    const int n = static_cast<int>(i - static_cast<long long>(m) * nv) << 3;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int s = 0; s < splits; ++s) {
      const float4* w = reinterpret_cast<const float4*>(ws + s * mn + static_cast<long long>(m) * N + n);
      const float4 a = w[0], b = w[1];
      acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
      acc[4] += b.x; acc[5] += b.y; acc[6] += b.z; acc[7] += b.w;
    }
    const float* rv = rowvec != nullptr ? rowvec + static_cast<long long>(m / rows_per_img) * ldv + n : nullptr;
    uint4 rr = make_uint4(0, 0, 0, 0);
    if (res != nullptr) rr = *reinterpret_cast<const uint4*>(res + static_cast<long long>(m) * ldr + n);
    const __half* rh = reinterpret_cast<const __half*>(&rr);
    uint4 o;
    __half* oh = reinterpret_cast<__half*>(&o);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float x = acc[j] * alpha;
      if (bias != nullptr) x += bias[n + j];
      if (rv != nullptr) x += rv[j];
      x += __half2float(rh[j]);
      oh[j] = __float2half_rn(x);
    }
    *reinterpret_cast<uint4*>(out + static_cast<long long>(m) * ldc + n) = o;
  }
}

static int launch_splitk_reduce(const float* ws, int splits, int M, int N, float alpha, const float* bias,
                                const float* rowvec, int rows_per_img, int ldv, const __half* res, long long ldr,
                                __half* out, long long ldc, cudaStream_t st) {
  const long long total = static_cast<long long>(M) * (N / 8);
  long long grid = (total + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 8;
  if (grid > cap) grid = cap;
  launch_k(splitk_reduce_kernel, dim3(static_cast<unsigned>(grid)), dim3(256), 0, st, ws, splits, M, N, alpha, bias, rowvec,
                                                                    rows_per_img, ldv, res, ldr, out, ldc);
  return check_launch("splitk_reduce");
}

}  // namespace icd

using namespace icd;

#ifdef ICD_GEMM_PROFILE
// debug builds only (make GPROF=1): the clock64 stamps of the LAST gemm launch, [256 CTAs][32 slots] (managed memory)
static long long* g_gemm_prof_buf = nullptr;
extern "C" long long* icd_gemm_prof_buffer(void) { return g_gemm_prof_buf; }
#endif

extern "C" int icd_gemm_pick_bn(int M, int N, int Z, int geglu, int b_mn_major, int force_bn) {
  return pick_tile(M, N, Z, 16, geglu, b_mn_major, force_bn, 0, 1).bn;
}

extern "C" int icd_gemm(const IcdGemm* g, void* stream) {
  if (g == nullptr || g->a0 == nullptr || g->b == nullptr || g->out == nullptr)
    return set_error("icd_gemm: null operand");
  if (g->M <= 0 || g->N <= 0 || g->Z <= 0) return set_error("icd_gemm: empty problem");
  const int kb_est = ((g->K0 + 63) / 64 + (g->a1 != nullptr ? (g->K1 + 63) / 64 : 0)) * (g->a_mode == 1 ? 9 : 1);
  // small-K GEMMs with a residual use the TMA-streamed residual variant, which exists for 128-row tiles only
  const int fbm = g->force_bm;
  // split-K (fp32 partials in the caller's workspace + a reduce kernel) for few-tile, deep-K problems
  const bool split_ok = g->ws != nullptr && !g->geglu && !g->out_fp32 && g->out_mode == GEMM_OUT_ROWMAJOR &&
                        g->Z == 1 && (g->N % 16) == 0 && g->upd_x == nullptr && g->exp_stats == nullptr &&
                        (g->ldc % 8) == 0 &&
                        (g->residual == nullptr || (g->ldr % 8) == 0);
  int max_splits = 1;
  if (split_ok) {
    const long long per = static_cast<long long>(g->M) * g->N * 4;
    max_splits = static_cast<int>(g->ws_bytes / (per > 0 ? per : 1));
    if (max_splits > 16) max_splits = 16;
    if (max_splits < 1) max_splits = 1;
  }
  TileChoice tc = pick_tile(g->M, g->N, g->Z, kb_est, g->geglu, g->b_mn_major, g->force_bn, fbm, max_splits);
  if (g->force_splits > 0) {
    tc = pick_tile(g->M, g->N, g->Z, kb_est, g->geglu, g->b_mn_major, g->force_bn, fbm, 1);
    if (split_ok && static_cast<long long>(g->force_splits) * g->M * g->N * 4 <= g->ws_bytes) tc.splits = g->force_splits;
  }
  const int bn = tc.bn, bm = tc.bm;
  if (bn != 64 && bn != 128 && bn != 160 && bn != 256) return set_error("icd_gemm: unsupported BN");
  if (g->b_mn_major && (bn % 64) != 0) return set_error("icd_gemm: MN-major B needs BN multiple of 64");
  if (g->geglu && (g->N % bn) != 0) return set_error("icd_gemm: GEGLU needs N % BN == 0");
  if (g->a1 != nullptr && (g->K0 % 64) != 0) return set_error("icd_gemm: concat needs K0 % 64 == 0");

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = g->M;
  p.N = g->N;
  p.Z = g->Z;
  p.ZA1 = g->ZA1 > 0 ? g->ZA1 : 1;
  p.ZB1 = g->ZB1 > 0 ? g->ZB1 : 1;
  p.a_mode = g->a_mode;
  p.b_mn_major = g->b_mn_major;
  const int kb0 = (g->K0 + 63) / 64;
  const int kb1 = g->a1 != nullptr ? (g->K1 + 63) / 64 : 0;
  p.kb_split = kb0;
  p.kb_per_tap = kb0 + kb1;
  p.splits = 1;

  CUtensorMap tmA0, tmA1, tmB;
  if (g->a_mode == GEMM_A_CONV3X3) {
    const int H = g->H, W = g->W, B = g->B;
    if (W > 128 ? (W % 128) != 0 : (128 % W) != 0)
      return set_error("icd_gemm(conv): W must divide 128 or be a multiple of it");
    const int hw = H * W;
    int tile_w = W, tile_h, tile_b;
    if (W > 128) {   // VAE resolutions (256 .. 1024 wide): one 128-row M tile = 128 consecutive pixels of one image row
      tile_w = 128;
      tile_h = 1;
      tile_b = 1;
    } else if (hw >= 128) {
      if ((hw % 128) != 0) return set_error("icd_gemm(conv): H*W must be a multiple of 128 (or divide it)");
      tile_h = 128 / W;
      tile_b = 1;
    } else {
      if ((128 % hw) != 0) return set_error("icd_gemm(conv): H*W must divide 128");
      tile_h = H;
      tile_b = 128 / hw;
    }
    if (g->M != B * hw) return set_error("icd_gemm(conv): M != B*H*W");
    p.H = H; p.W = W; p.tile_w = tile_w; p.tile_h = tile_h; p.tile_b = tile_b;
    p.num_kb = 9 * p.kb_per_tap;
    const uint32_t box[4] = {64, (uint32_t)tile_w, (uint32_t)tile_h, (uint32_t)tile_b};
    {
      const uint64_t dims[4] = {(uint64_t)g->K0, (uint64_t)W, (uint64_t)H, (uint64_t)B};
      const uint64_t str[3] = {(uint64_t)g->a0_ld * 2, (uint64_t)g->a0_ld * W * 2, (uint64_t)g->a0_ld * hw * 2};
      if (make_tmap_4d(&tmA0, g->a0, dims, str, box, 128, 2)) return 1;
    }
    if (g->a1 != nullptr) {
      const uint64_t dims[4] = {(uint64_t)g->K1, (uint64_t)W, (uint64_t)H, (uint64_t)B};
      const uint64_t str[3] = {(uint64_t)g->a1_ld * 2, (uint64_t)g->a1_ld * W * 2, (uint64_t)g->a1_ld * hw * 2};
      if (make_tmap_4d(&tmA1, g->a1, dims, str, box, 128, 2)) return 1;
    } else {
      tmA1 = tmA0;
    }
  } else {
    p.num_kb = p.kb_per_tap;
    const uint64_t z2 = (uint64_t)((g->Z + p.ZA1 - 1) / p.ZA1);
    const uint32_t box[4] = {64, 128, 1, 1};
    {
      const uint64_t dims[4] = {(uint64_t)g->K0, (uint64_t)g->M, (uint64_t)p.ZA1, z2};
      const uint64_t s1 = g->a_z1_stride > 0 ? (uint64_t)g->a_z1_stride * 2 : (uint64_t)g->a0_ld * 2;
      const uint64_t s2 = g->a_z2_stride > 0 ? (uint64_t)g->a_z2_stride * 2 : s1;
      const uint64_t str[3] = {(uint64_t)g->a0_ld * 2, s1, s2};
      if (make_tmap_4d(&tmA0, g->a0, dims, str, box, 128, 2)) return 1;
    }
    if (g->a1 != nullptr) {
      const uint64_t dims[4] = {(uint64_t)g->K1, (uint64_t)g->M, (uint64_t)p.ZA1, z2};
      const uint64_t s1 = g->a_z1_stride > 0 ? (uint64_t)g->a_z1_stride * 2 : (uint64_t)g->a1_ld * 2;
      const uint64_t s2 = g->a_z2_stride > 0 ? (uint64_t)g->a_z2_stride * 2 : s1;
      const uint64_t str[3] = {(uint64_t)g->a1_ld * 2, s1, s2};
      if (make_tmap_4d(&tmA1, g->a1, dims, str, box, 128, 2)) return 1;
    } else {
      tmA1 = tmA0;
    }
  }
  {
    const uint64_t ktot = (uint64_t)p.num_kb * 64;  // weights are zero-padded to whole K blocks only when needed:
    // the true extent is what the caller laid out; OOB columns are zero-filled by TMA.
    const uint64_t kreal = g->a_mode == GEMM_A_CONV3X3 ? ktot : (uint64_t)(g->a1 != nullptr ? g->K0 + g->K1 : g->K0);
    const uint64_t z2 = (uint64_t)((g->Z + p.ZB1 - 1) / p.ZB1);
    const uint64_t s1 = g->b_z1_stride > 0 ? (uint64_t)g->b_z1_stride * 2 : (uint64_t)g->b_ld * 2;
    const uint64_t s2 = g->b_z2_stride > 0 ? (uint64_t)g->b_z2_stride * 2 : s1;
    const uint64_t str[3] = {(uint64_t)g->b_ld * 2, s1, s2};
    const bool batched_b = g->b_z1_stride > 0 || g->b_z2_stride > 0;
    if (g->b_mn_major) {
      const uint64_t dims[4] = {(uint64_t)g->N, kreal, batched_b ? (uint64_t)p.ZB1 : 1, batched_b ? z2 : 1};
      const uint32_t box[4] = {64, 64, 1, 1};
      if (make_tmap_4d(&tmB, g->b, dims, str, box, 128, 2)) return 1;
    } else {
      const uint64_t dims[4] = {kreal, (uint64_t)g->N, batched_b ? (uint64_t)p.ZB1 : 1, batched_b ? z2 : 1};
      const uint32_t box[4] = {64, (uint32_t)bn, 1, 1};
      if (make_tmap_4d(&tmB, g->b, dims, str, box, 128, 2)) return 1;
    }
    p.b_batched = batched_b ? 1 : 0;
  }

  p.alpha = g->alpha;
  p.bias = g->bias;
  p.rowvec = g->rowvec;
  p.rows_per_img = g->rows_per_img > 0 ? g->rows_per_img : g->M;
  p.ldv = g->ldv;
  p.residual = reinterpret_cast<const __half*>(g->residual);
  p.ldr = g->ldr;
  p.res_zstride = g->res_zstride;
  p.out = g->out;
  p.ldc = g->ldc;
  p.out_z1_stride = g->out_z1_stride;
  p.out_z2_stride = g->out_z2_stride;
  p.out_imgstride = g->out_imgstride;
  p.out_fp32 = g->out_fp32;
  p.out_mode = g->out_mode;
  p.geglu = g->geglu;
  p.upd_x = g->upd_x;
  p.upd_out = g->upd_out;
  p.alpha_t = g->alpha_t; p.sigma_t = g->sigma_t; p.alpha_s = g->alpha_s; p.sigma_s = g->sigma_s;
  if (p.upd_x != nullptr && !(p.out_fp32 && p.out_mode == GEMM_OUT_TRANSPOSED))
    return set_error("icd_gemm: fused update needs fp32 transposed output");
  p.exp_stats = reinterpret_cast<const float2*>(g->exp_stats);
#ifdef ICD_GEMM_PROFILE
  {
    if (g_gemm_prof_buf == nullptr) cudaMallocManaged(&g_gemm_prof_buf, 256 * 32 * sizeof(long long));
    p.prof = g_gemm_prof_buf;
    const char* e = getenv("ICD_EPI_DEBUG");
    p.dbg = e ? atoi(e) : 0;
  }
#endif

  p.kb_per_split = p.num_kb;   // single split: the whole K range (num_kb is final here)
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define ICD_LAUNCH(BM_, BN_, EPI_) return launch<BM_, BN_, EPI_>(tmA0, tmA1, tmB, tmOut, tmRes, p, st)
#define ICD_LAUNCH_BN(BM_, EPI_)                 \
  switch (bn) {                                  \
    case 64: ICD_LAUNCH(BM_, 64, EPI_);          \
    case 128: ICD_LAUNCH(BM_, 128, EPI_);        \
    case 160: ICD_LAUNCH(BM_, 160, EPI_);        \
    default: ICD_LAUNCH(BM_, 256, EPI_);         \
  }
  if (tc.splits > 1) {
    // pass 1: raw fp32 partial products of each K range -> ws[split][M][N]; pass 2: reduce + fused epilogue
    GemmParams q = p;
    q.splits = tc.splits;
    q.kb_per_split = (p.num_kb + tc.splits - 1) / tc.splits;
    q.split_out_stride = static_cast<long long>(g->M) * g->N;
    q.alpha = 1.0f;
    q.bias = nullptr; q.rowvec = nullptr; q.residual = nullptr;
    q.out = g->ws; q.ldc = g->N; q.out_z1_stride = 0; q.out_z2_stride = 0;
    q.out_fp32 = 1; q.out_mode = GEMM_OUT_ROWMAJOR; q.epi_tma = 0; q.res_tma = 0;
    q.split_z = 1;
    CUtensorMap tmOut, tmRes = tmB;
    {   // ws viewed as [splits][M][N] fp32: 4th TMA coordinate = split
      const uint32_t box[4] = {32, 128, 1, 1};
      const uint64_t dims[4] = {(uint64_t)g->N, (uint64_t)g->M, 1, (uint64_t)tc.splits};
      const uint64_t str[3] = {(uint64_t)g->N * 4, (uint64_t)g->N * 4 * g->M, (uint64_t)g->N * 4 * g->M};
      if (make_tmap_4d(&tmOut, g->ws, dims, str, box, 128, 4)) return 1;
    }
    auto pass1 = [&]() -> int {
      const GemmParams& p = q;
      if (bm == 256) { ICD_LAUNCH_BN(256, EPI_STAGED_F32) }
      ICD_LAUNCH_BN(128, EPI_STAGED_F32)
    };
    if (pass1()) return 1;
    return launch_splitk_reduce(reinterpret_cast<const float*>(g->ws), tc.splits, g->M, g->N, g->alpha, g->bias,
                                g->rowvec, p.rows_per_img, g->ldv, reinterpret_cast<const __half*>(g->residual),
                                g->ldr, reinterpret_cast<__half*>(g->out), g->ldc, st);
  }

  // staged TMA-store epilogue whenever the output is an fp16 row-major matrix TMA can address
  CUtensorMap tmOut = tmB, tmRes = tmB;
  const bool aligned = (g->ldc % 8) == 0 && (reinterpret_cast<uint64_t>(g->out) & 15) == 0 &&
                       (g->out_z1_stride % 8) == 0 && (g->out_z2_stride % 8) == 0;
  const bool res_ok = g->residual == nullptr ||
                      ((g->ldr % 8) == 0 && (reinterpret_cast<uint64_t>(g->residual) & 15) == 0 && g->Z == 1);
  const int n_out_chk = g->geglu ? g->N / 2 : g->N;
  p.epi_tma = (!g->out_fp32 && g->out_mode == GEMM_OUT_ROWMAJOR && aligned && res_ok && (n_out_chk % 8) == 0) ? 1 : 0;
  if (g->geglu && !p.epi_tma) return set_error("icd_gemm: GEGLU needs an aligned fp16 row-major output");
  if (g->exp_stats != nullptr && (!p.epi_tma || g->geglu || g->residual != nullptr))
    return set_error("icd_gemm: exp_stats needs a plain aligned fp16 row-major output (N % 8 == 0, no residual)");
  // per-warp epilogue (EPI_WARP / EPI_WARP_RES): every epilogue warp stages and stores its own 32x32 block
  static const bool warp_epi_enabled = [] { const char* e = getenv("ICD_GEMM_WARP_EPI"); return e == nullptr || atoi(e) != 0; }();
  // Large plain / GEGLU linears: one 256x256 tile per CTA pair (tcgen05.mma.cta_group::2, gemm2sm_tc.cu) where the
  // measurements say it beats the single-CTA kernel
  const bool use_2sm = p.epi_tma && g->a_mode == GEMM_A_TILED && g->a1 == nullptr && g->Z == 1 && !g->b_mn_major &&
                       g->residual == nullptr && g->rowvec == nullptr && g->alpha == 1.0f && g->exp_stats == nullptr &&
                       g->force_bm == 0 && g->force_splits == 0 && (g->force_bn == 0 || g->force_bn == 256) &&
                       g->a_z1_stride == 0 && g->a_z2_stride == 0 && g->b_z1_stride == 0 && g->b_z2_stride == 0 &&
                       gemm2sm_wanted(g->M, g->N, g->K0, g->geglu != 0);
  const bool warp_epi = warp_epi_enabled && p.epi_tma && !g->geglu && !use_2sm &&
                        (g->residual == nullptr || g->alpha == 1.0f);   // the MMA-added residual is not scaled by alpha
  // GEGLU projections that stay on the single-CTA kernel (K < 1024): per-warp epilogue too (ICD_GEMM_WARP_GEGLU=0:
  // the CTA-wide staged one)
  static const bool warp_geglu_enabled = [] { const char* e = getenv("ICD_GEMM_WARP_GEGLU"); return e == nullptr || atoi(e) != 0; }();
  const bool warp_geglu = warp_epi_enabled && warp_geglu_enabled && p.epi_tma && g->geglu && !use_2sm;
  if (p.epi_tma) {
    const uint64_t n_out = (uint64_t)(g->geglu ? g->N / 2 : g->N);
    const uint64_t z2 = (uint64_t)((g->Z + p.ZA1 - 1) / p.ZA1);
    const uint32_t box[4] = {32, (warp_epi || warp_geglu) ? 32u : 128u, 1, 1};
    const uint64_t dims[4] = {n_out, (uint64_t)g->M, (uint64_t)p.ZA1, z2};
    const uint64_t s1 = g->out_z1_stride > 0 ? (uint64_t)g->out_z1_stride * 2 : (uint64_t)g->ldc * 2;
    const uint64_t s2 = g->out_z2_stride > 0 ? (uint64_t)g->out_z2_stride * 2 : s1;
    const uint64_t str[3] = {(uint64_t)g->ldc * 2, s1, s2};
    if (make_tmap_4d(&tmOut, g->out, dims, str, box, 64, 2)) return 1;
    p.res_tma = g->residual != nullptr ? 1 : 0;
    if (g->residual != nullptr) {
      const uint64_t rdims[4] = {n_out, (uint64_t)g->M, 1, 1};
      const uint64_t rstr[3] = {(uint64_t)g->ldr * 2, (uint64_t)g->ldr * 2, (uint64_t)g->ldr * 2};
      if (warp_epi) {
        // EPI_WARP adds the residual on the tensor core: A-operand style atoms (64 columns x 128 rows, 128B swizzle)
        const uint32_t rbox[4] = {64, 128, 1, 1};
        if (make_tmap_4d(&tmRes, g->residual, rdims, rstr, rbox, 128, 2)) return 1;
        p.res_mma = 1;
      } else if (make_tmap_4d(&tmRes, g->residual, rdims, rstr, box, 64, 2)) {
        return 1;
      }
    }
  }

  if (use_2sm) {
    CUtensorMap tmB2;
    const uint64_t dims[4] = {(uint64_t)g->K0, (uint64_t)g->N, 1, 1};
    const uint64_t str[3] = {(uint64_t)g->b_ld * 2, (uint64_t)g->b_ld * 2, (uint64_t)g->b_ld * 2};
    const uint32_t box[4] = {64, 128, 1, 1};   // each CTA of the pair loads half of the 256-row B tile
    if (make_tmap_4d(&tmB2, g->b, dims, str, box, 128, 2)) return 1;
    return launch_gemm2sm(tmA0, tmB2, tmOut, g->M, g->N, p.num_kb, g->bias, g->geglu != 0, st);
  }

  if (g->out_fp32 && g->out_mode == GEMM_OUT_ROWMAJOR && g->residual == nullptr && g->rowvec == nullptr &&
      (g->N % 16) == 0 && (g->ldc % 4) == 0 && (reinterpret_cast<uint64_t>(g->out) & 15) == 0 &&
      (g->out_z1_stride % 4) == 0 && (g->out_z2_stride % 4) == 0) {
    const uint64_t z2 = (uint64_t)((g->Z + p.ZA1 - 1) / p.ZA1);
    const uint32_t box[4] = {32, 128, 1, 1};
    const uint64_t dims[4] = {(uint64_t)g->N, (uint64_t)g->M, (uint64_t)p.ZA1, z2};
    const uint64_t s1 = g->out_z1_stride > 0 ? (uint64_t)g->out_z1_stride * 4 : (uint64_t)g->ldc * 4;
    const uint64_t s2 = g->out_z2_stride > 0 ? (uint64_t)g->out_z2_stride * 4 : s1;
    const uint64_t str[3] = {(uint64_t)g->ldc * 4, s1, s2};
    if (make_tmap_4d(&tmOut, g->out, dims, str, box, 128, 4)) return 1;
    p.split_z = 0;
    if (bm == 256) { ICD_LAUNCH_BN(256, EPI_STAGED_F32) }
    ICD_LAUNCH_BN(128, EPI_STAGED_F32)
  }
  if (g->geglu) {
    if (bm != 128) return set_error("icd_gemm: GEGLU uses 128-row tiles");
    if (warp_geglu) {
      if (bn == 128) ICD_LAUNCH(128, 128, EPI_WARP_GEGLU);
      if (bn == 256) ICD_LAUNCH(128, 256, EPI_WARP_GEGLU);
    }
    if (bn == 128) ICD_LAUNCH(128, 128, EPI_STAGED_GEGLU);
    if (bn == 256) ICD_LAUNCH(128, 256, EPI_STAGED_GEGLU);
    return set_error("icd_gemm: GEGLU supports BN 128 / 256");
  }
  if (warp_epi) {
    if (bm == 256) { ICD_LAUNCH_BN(256, EPI_WARP) }
    ICD_LAUNCH_BN(128, EPI_WARP)
  }
  // short main loops cannot hide the row-per-thread residual reads: stream the residual through TMA + smem
  if (p.epi_tma && p.res_tma && p.num_kb <= 24) {
    if (bm == 128) { ICD_LAUNCH_BN(128, EPI_STAGED_RES) }
    if (bn == 64) ICD_LAUNCH(256, 64, EPI_STAGED_RES);
    if (bn == 128) ICD_LAUNCH(256, 128, EPI_STAGED_RES);
    if (bn == 160) ICD_LAUNCH(256, 160, EPI_STAGED_RES);
  }
  if (p.epi_tma) {
    if (bm == 256) { ICD_LAUNCH_BN(256, EPI_STAGED) }
    ICD_LAUNCH_BN(128, EPI_STAGED)
  }
  if (bm == 256) { ICD_LAUNCH_BN(256, EPI_DIRECT) }
  ICD_LAUNCH_BN(128, EPI_DIRECT)
#undef ICD_LAUNCH_BN
#undef ICD_LAUNCH
}
