#include "host_util.h"

#include <cstdlib>

#include "../../include/icd_b200.h"

namespace icd {
static thread_local std::string g_err;
int set_error(const std::string& msg) {
  g_err = msg;
  return 1;
}
int device_ordinal() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return dev;
}
int sm_count() {
  static int n[kMaxDevices] = {};
  const int dev = device_ordinal() & (kMaxDevices - 1);
  if (n[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0)
      n[dev] = v;
    else {
      cudaGetLastError();
      return 148;
    }
  }
  return n[dev];
}
static int g_pdl = -1;
bool pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("ICD_PDL");
    g_pdl = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return g_pdl != 0;
}
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(std::string(what) + ": " + cudaGetErrorString(e));
  return 0;
}
}  // namespace icd

extern "C" const char* icd_last_error(void) { return icd::g_err.c_str(); }
extern "C" int icd_abi_version(void) { return 3; }   // 2: IcdGemm.exp_stats, icd_attention_ex, VAE / CLIP helper kernels; 3: fp32 validation path (icd_*_f32)
extern "C" int icd_set_pdl(int enabled) {
  const int prev = icd::pdl_enabled() ? 1 : 0;
  icd::g_pdl = enabled ? 1 : 0;
  return prev;
}
extern "C" int icd_device_info(int* sms, int* major, int* minor) {
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    cudaGetLastError();
    return icd::set_error("no CUDA device visible");
  }
  if (sms) *sms = prop.multiProcessorCount;
  if (major) *major = prop.major;
  if (minor) *minor = prop.minor;
  return 0;
}
