// Small device helpers shared by the convolution kernels (conv_tc.cu, conv_pair_tc.cu).
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace oidnb200 {
namespace {

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b)
{
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// {lo, hi} -> packed fp16 pair with max(x, 0) folded into the conversion
__device__ __forceinline__ uint32_t pack_half2_relu(float lo, float hi)
{
  uint32_t r;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

__device__ __forceinline__ uint32_t max_half2(uint32_t a, uint32_t b)
{
  uint32_t r;
  asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}

__device__ __forceinline__ void st_global_32B(void* ptr, const uint32_t* h)
{
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :: "l"(ptr), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]), "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7])
               : "memory");
}

} // namespace
} // namespace oidnb200
