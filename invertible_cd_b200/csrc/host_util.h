// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>

#include <string>

namespace icd {
int set_error(const std::string& msg);  // records the message, returns 1
int sm_count();                          // SMs of the current device (148 on B200); 148 if no device is visible
int check_launch(const char* what);      // cudaGetLastError -> set_error
}  // namespace icd
