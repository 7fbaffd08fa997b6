// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>

#include <string>

namespace icd {
int set_error(const std::string& msg);  // records the message, returns 1
int sm_count();                          // SMs of the CURRENT device (148 on B200); 148 if no device is visible
int device_ordinal();                    // cudaGetDevice (0 if no device is visible)
constexpr int kMaxDevices = 64;
// One-time-per-DEVICE flag (cudaFuncSetAttribute and friends are per device, not per process): a library loaded
// once may serve several GPUs of one process (load_models(model_id, 'cuda:1', ...)).
struct PerDeviceFlag {
  bool v[kMaxDevices] = {};
  bool& cur() { return v[device_ordinal() & (kMaxDevices - 1)]; }
};
int check_launch(const char* what);      // cudaGetLastError -> set_error
bool pdl_enabled();                      // programmatic dependent launch (ICD_PDL=0 or icd_set_pdl(0) disables)

// Launch with the programmatic-stream-serialization attribute: the kernel may start (up to its pdl_wait()) before
// its predecessor in the stream has finished. Every kernel launched through this calls pdl_wait() before touching
// global memory. Works under stream capture (becomes a programmatic edge of the CUDA graph).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            Args&&... args) {
  cudaLaunchConfig_t cfg;
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Cooperative launch (+ the same programmatic-stream-serialization attribute): the driver guarantees that every CTA
// of the grid is resident at once (or refuses the launch), which grid-wide barriers inside the kernel rely on.
// Measured on B200 (tools/probes/coop_probe.cu, profiles/r2_coop_probe.txt): accepted together with the PDL attribute,
// eagerly and under stream capture, at no extra cost per launch in a graph.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k_coop(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg;
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
}  // namespace icd
