// Probe: can a kernel that needs grid-wide co-residency (single-pass GroupNorm) be launched COOPERATIVELY together
// with the programmatic-dependent-launch attribute, eagerly and under stream capture, and what does it cost?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o coop_probe coop_probe.cu && ./coop_probe
#include <cuda_runtime.h>

#include <cstdio>
#include <vector>

__device__ unsigned int g_ctr[2];

__global__ void plain_kernel(float* x, int n) {
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = x[i] * 1.0001f + 1.0f;
}

// grid barrier through global counters (what gn_fused_kernel does): needs every CTA resident
__global__ void barrier_kernel(float* x, int n) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float v = i < n ? x[i] : 0.f;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(&g_ctr[0], 1u);
    volatile unsigned int* a = &g_ctr[0];
    while (*a < gridDim.x) __nanosleep(32);
    __threadfence();
  }
  __syncthreads();
  asm volatile("griddepcontrol.launch_dependents;");
  if (threadIdx.x == 0) {
    const unsigned d = atomicAdd(&g_ctr[1], 1u);
    if (d == gridDim.x - 1) {
      g_ctr[1] = 0u;
      __threadfence();
      g_ctr[0] = 0u;
    }
  }
  if (i < n) x[i] = v + 1.0f;
}

static cudaError_t launch(void (*k)(float*, int), int grid, int block, cudaStream_t st, bool coop, bool pdl, float* x,
                          int n) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (coop) {
    attr[na].id = cudaLaunchAttributeCooperative;
    attr[na].val.cooperative = 1;
    ++na;
  }
  if (pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, k, x, n);
}

static float time_graph(cudaStream_t st, bool coop, bool pdl_on_barrier, float* x, int n, int grid) {
  cudaGraph_t g;
  cudaGraphExec_t ge;
  if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) return -1.f;
  cudaError_t e = cudaSuccess;
  for (int i = 0; i < 100 && e == cudaSuccess; ++i) {
    e = launch(plain_kernel, grid, 256, st, false, true, x, n);
    if (e == cudaSuccess) e = launch(barrier_kernel, grid, 256, st, coop, pdl_on_barrier, x, n);
  }
  cudaError_t e2 = cudaStreamEndCapture(st, &g);
  if (e != cudaSuccess || e2 != cudaSuccess) {
    printf("    capture failed: launch=%s end=%s\n", cudaGetErrorString(e), cudaGetErrorString(e2));
    cudaGetLastError();
    return -1.f;
  }
  if (cudaGraphInstantiate(&ge, g, 0) != cudaSuccess) {
    printf("    instantiate failed: %s\n", cudaGetErrorString(cudaGetLastError()));
    return -1.f;
  }
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) cudaGraphLaunch(ge, st);
  cudaEventRecord(a, st);
  for (int i = 0; i < 10; ++i) cudaGraphLaunch(ge, st);
  cudaEventRecord(b, st);
  cudaError_t es = cudaStreamSynchronize(st);
  if (es != cudaSuccess) {
    printf("    replay failed: %s\n", cudaGetErrorString(es));
    return -1.f;
  }
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  cudaGraphExecDestroy(ge);
  cudaGraphDestroy(g);
  return ms / 10.f / 200.f * 1000.f;   // us per kernel
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int coop_ok = 0;
  cudaDeviceGetAttribute(&coop_ok, cudaDevAttrCooperativeLaunch, 0);
  printf("SMs %d cooperativeLaunch attr %d\n", sms, coop_ok);
  const int grid = sms * 2, n = grid * 256;
  float* x;
  cudaMalloc(&x, n * sizeof(float));
  cudaMemset(x, 0, n * sizeof(float));
  cudaStream_t st;
  cudaStreamCreate(&st);
  for (int coop = 0; coop < 2; ++coop)
    for (int pdl = 0; pdl < 2; ++pdl) {
      cudaError_t e = launch(barrier_kernel, grid, 256, st, coop, pdl, x, n);
      cudaError_t s = cudaStreamSynchronize(st);
      printf("eager  coop=%d pdl=%d : launch=%s sync=%s\n", coop, pdl, cudaGetErrorString(e), cudaGetErrorString(s));
      cudaGetLastError();
    }
  for (int coop = 0; coop < 2; ++coop)
    for (int pdl = 0; pdl < 2; ++pdl) {
      const float us = time_graph(st, coop, pdl, x, n, grid);
      printf("graph  coop=%d pdl=%d : %.2f us per kernel (100 x [plain+PDL, barrier])\n", coop, pdl, us);
    }
  // too-large cooperative grid must be refused, not deadlock
  cudaError_t e = launch(barrier_kernel, sms * 64, 256, st, true, false, x, n);
  printf("coop oversubscribed grid (%d CTAs): %s\n", sms * 64, cudaGetErrorString(e));
  cudaGetLastError();
  cudaDeviceSynchronize();
  return 0;
}
