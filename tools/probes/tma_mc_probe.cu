// Micro-probe: aggregate L2 -> SM delivery rate of TMA tile loads when the CTAs of a thread-block cluster need the
// SAME tile (the A tile shared by GEMM CTAs of one M row, the K/V tile shared by attention CTAs of one head):
//   unicast   every CTA loads the whole [256 x 64] fp16 tile (32 KB) itself
//   multicast every CTA loads 1/C of the rows and multicasts them to all C CTAs of the cluster
// Delivered bytes per CTA are identical in both modes; only the number of L2 reads / crossbar transfers differs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I invertible_cd_b200/csrc tools/probes/tma_mc_probe.cu -o tools/probes/tma_mc_probe
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "icd_ptx.cuh"
using namespace icd;

constexpr int ROWS = 256, COLS = 64, TILE_BYTES = ROWS * COLS * 2, STAGES = 4;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                               uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, "
      "%4}], [%2], %5;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

// tmFull: box [64 x 256]; tmSlice: box [64 x 256/C]
template <int C, bool MC>
__global__ void __launch_bounds__(64) probe(const __grid_constant__ CUtensorMap tmFull,
                                            const __grid_constant__ CUtensorMap tmSlice, int tiles, int row_blocks,
                                            int col_blocks, int shared_tile) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[STAGES], empty[STAGES];
  const uint32_t rank = C > 1 ? cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / C;
  const int n_clusters = gridDim.x / C;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], MC ? C : 1);
    }
    fence_mbar_init();
  }
  __syncthreads();
  if (C > 1) cluster_sync();
  if (threadIdx.x == 0) {
    for (int t = 0; t < tiles; ++t) {
      const int s = t % STAGES;
      const uint32_t ph = (t / STAGES) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      // the tile this cluster needs (shared_tile: all CTAs of the cluster read the same one, as GEMM CTAs share A)
      const int owner = shared_tile ? cluster_id : blockIdx.x;
      const int n_own = shared_tile ? n_clusters : gridDim.x;
      const int tile = (owner + t * n_own) % (row_blocks * col_blocks);
      const int rb = tile % row_blocks, cb = tile / row_blocks;
      mbar_expect_tx(&full[s], TILE_BYTES);
      uint8_t* dst = smem + s * TILE_BYTES;
      if (MC) {
        constexpr int SL = ROWS / C;
        tma_load_2d_mc(dst + rank * SL * COLS * 2, &tmSlice, &full[s], cb * COLS, rb * ROWS + rank * SL,
                       static_cast<uint16_t>((1u << C) - 1));
      } else {
        tma_load_2d(dst, &tmFull, &full[s], cb * COLS, rb * ROWS);
      }
    }
  } else if (threadIdx.x == 32) {
    for (int t = 0; t < tiles; ++t) {
      const int s = t % STAGES;
      mbar_wait(&full[s], (t / STAGES) & 1);
      if (MC) {
        for (uint32_t c = 0; c < C; ++c) mbar_arrive_remote(&empty[s], c);
      } else {
        mbar_arrive(&empty[s]);
      }
    }
  }
  __syncthreads();
  if (C > 1) cluster_sync();   // no CTA may exit while peers can still multicast into its smem / barriers
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int C, bool MC>
static void run(EncodeFn enc, void* data, int rows, int cols, int grid, int tiles, int shared_tile) {
  CUtensorMap full, slice;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t es[2] = {1, 1};
  cuuint32_t box[2] = {COLS, ROWS};
  enc(&full, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, data, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  cuuint32_t box2[2] = {COLS, (cuuint32_t)(ROWS / C)};
  enc(&slice, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, data, dims, strides, box2, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  const int smem = STAGES * TILE_BYTES;
  cudaFuncSetAttribute(probe<C, MC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(64);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const int rb = rows / ROWS, cbk = cols / COLS;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int it = 0; it < 4; ++it) {
    cudaEventRecord(e0);
    cudaLaunchKernelEx(&cfg, probe<C, MC>, full, slice, tiles, rb, cbk, shared_tile);
    cudaEventRecord(e1);
    cudaError_t err = cudaEventSynchronize(e1);
    if (err != cudaSuccess) { printf("C=%d MC=%d: %s\n", C, (int)MC, cudaGetErrorString(err)); exit(1); }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (it > 0 && ms < best) best = ms;
  }
  const double bytes = (double)grid * tiles * TILE_BYTES;
  printf("cluster %d  %-9s  %-12s grid %3d : %8.1f us   delivered %7.2f TB/s  (%5.1f B/clk/SM @1.9GHz)\n", C,
         MC ? "multicast" : "unicast", shared_tile ? "shared-tile" : "private-tile", grid, best * 1e3,
         bytes / (best * 1e-3) / 1e12, bytes / grid / (best * 1e-3) / 1.9e9);
}

int main() {
  const int rows = 16384, cols = 1024;   // 32 MB fp16: L2-resident after the first pass
  void* data;
  cudaMalloc(&data, (size_t)rows * cols * 2);
  cudaMemset(data, 0, (size_t)rows * cols * 2);
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", reinterpret_cast<void**>(&enc), cudaEnableDefault, &qres);
  if (enc == nullptr) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  const int tiles = 512;
  run<1, false>(enc, data, rows, cols, 144, tiles, 0);
  run<2, false>(enc, data, rows, cols, 144, tiles, 1);
  run<2, true>(enc, data, rows, cols, 144, tiles, 1);
  run<4, false>(enc, data, rows, cols, 144, tiles, 1);
  run<4, true>(enc, data, rows, cols, 144, tiles, 1);
  run<8, false>(enc, data, rows, cols, 144, tiles, 1);
  run<8, true>(enc, data, rows, cols, 144, tiles, 1);
  return 0;
}
