// Transfer functions and the per-pixel output math, shared by the elementwise passes
// (elementwise.cu) and the conv kernel's fused output-process epilogue (conv_tc.cu).
// Functional spec: core/color.h:29-165, devices/gpu/gpu_output_process.h:35-73.
#pragma once
#include "../../../include/oidn_b200_kernels.h"
#include <cuda_runtime.h>
#include <cfloat>
#include <cmath>

namespace oidnb200 {

// ------------------------------------------------------------------------------------------------
// Transfer function (core/color.h:29-165)
// ------------------------------------------------------------------------------------------------
struct Transfer
{
  int type;
  float norm, rcp_norm;
  float input_scale;
  const float* input_scale_ptr;
};

constexpr float kSrgbA = 12.92f, kSrgbB = 1.055f, kSrgbC = 1.f / 2.4f, kSrgbD = -0.055f;
constexpr float kSrgbY0 = 0.0031308f, kSrgbX0 = 0.04045f;
constexpr float kPuA = 1.41283765e+03f, kPuB = 1.64593172e+00f, kPuC = 4.31384981e-01f;
constexpr float kPuD = -2.94139609e-03f, kPuE = 1.92653254e-01f, kPuF = 6.26026094e-03f;
constexpr float kPuG = 9.98620152e-01f, kPuY0 = 1.57945760e-06f, kPuY1 = 3.22087631e-02f;
constexpr float kPuX0 = 2.23151711e-03f, kPuX1 = 3.70974749e-01f;

__host__ __device__ inline float tf_raw_forward(int type, float y)
{
  switch (type)
  {
  case OIDNB200_TF_SRGB:
    return y <= kSrgbY0 ? kSrgbA * y : kSrgbB * powf(y, kSrgbC) + kSrgbD;
  case OIDNB200_TF_PU:
    if (y <= kPuY0) return kPuA * y;
    if (y <= kPuY1) return kPuB * powf(y, kPuC) + kPuD;
    return kPuE * logf(y + kPuF) + kPuG;
  case OIDNB200_TF_LOG:
    return logf(y + 1.f);
  default:
    return y;
  }
}

// Forward transfer function of the input process. The result is stored as fp16 (relative spacing
// 4.9e-4), so the power / logarithm go through the SFU (lg2.approx / ex2.approx: relative error of
// the result < 4e-6 over the segments' ranges) instead of libdevice's ~100-instruction powf: with
// powf the pass was issue-bound (ncu: 466 instructions per pixel, 45 % of HBM peak), not HBM-bound.
// The output process keeps the exact functions (its result is the user's fp32 image).
__device__ __forceinline__ float tf_forward(const Transfer& t, float y)
{
  switch (t.type)
  {
  case OIDNB200_TF_SRGB:
    return y <= kSrgbY0 ? kSrgbA * y : kSrgbB * __powf(y, kSrgbC) + kSrgbD;
  case OIDNB200_TF_PU:
  {
    float x;
    if (y <= kPuY0)      x = kPuA * y;
    else if (y <= kPuY1) x = kPuB * __powf(y, kPuC) + kPuD;
    else                 x = kPuE * __logf(y + kPuF) + kPuG;
    return x * t.norm;
  }
  case OIDNB200_TF_LOG:
    return logf(y + 1.f) * t.norm; // exact: lg2.approx's absolute error would show for y << 1
  default:
    return y;
  }
}

// Inverse transfer function of the output process. The segment formulas are evaluated branch-free
// (a warp's pixels fall in different segments) with multiplications by the reciprocal constants and
// the power through the SFU (lg2.approx / ex2.approx, relative error of the result < 4e-6 on the
// segment's range); expf stays libdevice's (2 ulp). This function runs per pixel in the epilogue of
// the last convolution, where libdevice's powf and IEEE divisions made the epilogue warps the
// bottleneck. (The reference's CPU device evaluates these with ISPC's approximate math library.)
__device__ __forceinline__ float tf_inverse(const Transfer& t, float x)
{
  switch (t.type)
  {
  case OIDNB200_TF_SRGB:
  {
    const float lin = x * (1.f / kSrgbA);
    const float pw  = __powf((x - kSrgbD) * (1.f / kSrgbB), 1.f / kSrgbC);
    return x <= kSrgbX0 ? lin : pw;
  }
  case OIDNB200_TF_PU:
  {
    const float u   = x * t.rcp_norm;
    const float lin = u * (1.f / kPuA);
    const float pw  = __powf((u - kPuD) * (1.f / kPuB), 1.f / kPuC);
    const float ex  = expf((u - kPuG) * (1.f / kPuE)) - kPuF;
    return u <= kPuX0 ? lin : (u <= kPuX1 ? pw : ex);
  }
  case OIDNB200_TF_LOG:
    return expf(x * t.rcp_norm) - 1.f;
  default:
    return x;
  }
}

inline bool make_transfer(const oidnb200_transfer* tf, Transfer& t)
{
  if (!tf || tf->type < OIDNB200_TF_LINEAR || tf->type > OIDNB200_TF_LOG) return false;
  t.type = tf->type;
  // core/color.cpp:9-16: normScale = 1/forward(yMax), evaluated on the host in fp32
  const float xmax = tf_raw_forward(tf->type, 65504.f);
  t.norm = (float)(1. / xmax);
  t.rcp_norm = xmax;
  t.input_scale = tf->input_scale;
  t.input_scale_ptr = tf->input_scale_ptr;
  return true;
}

__device__ __forceinline__ float nan_to_zero(float x) { return isnan(x) ? 0.f : x; }
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

__device__ __forceinline__ float3 output_pixel(const Transfer& tf, bool hdr, bool snorm, bool mono, float oscale, float x, float y, float z)
{
  float v[3] = {x, y, z};
#pragma unroll
  for (int k = 0; k < 3; ++k)
    v[k] = tf_inverse(tf, clampf(nan_to_zero(v[k]), 0.f, FLT_MAX));
  if (mono)
  {
    const float m = (v[0] + v[1] + v[2]) * (1.f / 3.f);
    v[0] = v[1] = v[2] = m;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k)
  {
    if (snorm) v[k] = fmaxf(v[k] * 2.f - 1.f, -1.f);
    if (!hdr) v[k] = fminf(v[k], 1.f);
    v[k] *= oscale;
  }
  return make_float3(v[0], v[1], v[2]);
}

__device__ __forceinline__ float output_scale(const Transfer& tf)
{
  const float s = tf.input_scale_ptr ? *tf.input_scale_ptr : tf.input_scale;
  return s != 0.f ? 1.f / s : 0.f; // core/color.h:95-123
}

} // namespace oidnb200
