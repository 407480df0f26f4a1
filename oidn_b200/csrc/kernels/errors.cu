// Per-thread error text + device enumeration for the C ABI.
#include "common.h"
#include "../../../include/oidn_b200_kernels.h"
#include <cuda_runtime.h>

namespace oidnb200 {
static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
} // namespace oidnb200

extern "C" {

const char* oidnb200_last_error(void) { return oidnb200::g_error.c_str(); }

int oidnb200_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess)
  {
    cudaGetLastError();
    return 0;
  }
  int ok = 0;
  for (int i = 0; i < n; ++i)
  {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10)
      ++ok;
  }
  return ok;
}

} // extern "C"
