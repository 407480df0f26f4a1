// Bandwidth-bound passes of the denoising path: input process, output process, autoexposure,
// image copy, and the stand-alone pool / upsample ops.
//
// Functional spec: devices/gpu/gpu_input_process.h:37-176, gpu_output_process.h:35-73,
// gpu_autoexposure.h:13-164, gpu_image_copy.h:15-27, gpu_pool.h:33-52, gpu_upsample.h:33-52 and
// their CPU twins (devices/cpu/cpu_input_process.isph:31-136, cpu_output_process.isph:29-70,
// cpu_autoexposure.cpp:22-64). Transfer functions: core/color.h:29-165.
//
// These kernels move bytes: each thread owns whole pixels; on the fast paths a warp owns a span of
// 128 pixels, moves packed fp32 RGB rows as consecutive 16-byte chunks (staged through shared
// memory to change ownership from chunks to pixels) and writes a pixel of the 16-channel network
// input with one 32-byte store; the autoexposure is a bin pass with warp-shuffle reductions plus
// a fixed-order fold instead of the reference's three launches.
#include "common.h"
#include "transfer.cuh"
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cfloat>
#include <cmath>
#include <string>

namespace oidnb200 {
namespace {

// ------------------------------------------------------------------------------------------------
// Image accessor (core/image_accessor.h:16-94): C==1 -> (x,x,x), C==2 -> (x,y,y)
// ------------------------------------------------------------------------------------------------
struct Img
{
  uint8_t* ptr;
  int C;
  int is_half;
  int W, H;
  size_t ps, rs;
};

bool make_img(const oidnb200_image* im, Img& o)
{
  o = Img{nullptr, 0, 0, 0, 0, 0, 0};
  if (!im || !im->ptr) return true;
  switch (im->format)
  {
  case OIDNB200_FORMAT_FLOAT: case OIDNB200_FORMAT_FLOAT2: case OIDNB200_FORMAT_FLOAT3:
    o.C = im->format - OIDNB200_FORMAT_FLOAT + 1; o.is_half = 0; break;
  case OIDNB200_FORMAT_HALF: case OIDNB200_FORMAT_HALF2: case OIDNB200_FORMAT_HALF3:
    o.C = im->format - OIDNB200_FORMAT_HALF + 1; o.is_half = 1; break;
  default:
    return false;
  }
  o.ptr = static_cast<uint8_t*>(im->ptr);
  o.W = im->W; o.H = im->H; o.ps = im->pixel_stride; o.rs = im->row_stride;
  return true;
}

__device__ __forceinline__ float3 img_get3(const Img& im, int h, int w)
{
  const uint8_t* px = im.ptr + (size_t)h * im.rs + (size_t)w * im.ps;
  float x, y, z;
  if (im.is_half)
  {
    const __half* p = reinterpret_cast<const __half*>(px);
    x = __half2float(p[0]);
    y = im.C >= 2 ? __half2float(p[1]) : x;
    z = im.C == 3 ? __half2float(p[2]) : y;
  }
  else
  {
    const float* p = reinterpret_cast<const float*>(px);
    x = p[0];
    y = im.C >= 2 ? p[1] : x;
    z = im.C == 3 ? p[2] : y;
  }
  return make_float3(x, y, z);
}

__device__ __forceinline__ void img_set3(const Img& im, int h, int w, float3 v)
{
  uint8_t* px = im.ptr + (size_t)h * im.rs + (size_t)w * im.ps;
  if (im.is_half)
  {
    __half* p = reinterpret_cast<__half*>(px);
    p[0] = __float2half_rn(v.x);
    if (im.C >= 2) p[1] = __float2half_rn(v.y);
    if (im.C == 3) p[2] = __float2half_rn(v.z);
  }
  else
  {
    float* p = reinterpret_cast<float*>(px);
    p[0] = v.x;
    if (im.C >= 2) p[1] = v.y;
    if (im.C == 3) p[2] = v.z;
  }
}

__device__ __forceinline__ uint32_t pack2(float a, float b)
{
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ------------------------------------------------------------------------------------------------
// Input process
// ------------------------------------------------------------------------------------------------
struct InputParams
{
  Img input, albedo, normal;
  oidnb200_tile tile;
  Transfer tf;
  int hdr, snorm;
  __half* dst;
  int TH, TW, C; // C = 16
};

__device__ __forceinline__ void input_pixel(const InputParams& p, float scale, float3 c, float3 a, float3 n,
                                            bool has_a, bool has_n, uint4& lo, uint4& hi)
{
  const float cmin = p.snorm ? -1.f : 0.f, cmax = p.hdr ? FLT_MAX : 1.f;
  float v[3] = {c.x, c.y, c.z};
#pragma unroll
  for (int k = 0; k < 3; ++k)
  {
    float x = clampf(nan_to_zero(v[k] * scale), cmin, cmax);
    if (p.snorm) x = x * 0.5f + 0.5f;
    v[k] = tf_forward(p.tf, x);
  }
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, n0 = 0.f, n1 = 0.f, n2 = 0.f;
  if (has_a)
  {
    a0 = clampf(nan_to_zero(a.x), 0.f, 1.f); a1 = clampf(nan_to_zero(a.y), 0.f, 1.f); a2 = clampf(nan_to_zero(a.z), 0.f, 1.f);
  }
  if (has_n)
  {
    n0 = clampf(nan_to_zero(n.x), -1.f, 1.f) * 0.5f + 0.5f;
    n1 = clampf(nan_to_zero(n.y), -1.f, 1.f) * 0.5f + 0.5f;
    n2 = clampf(nan_to_zero(n.z), -1.f, 1.f) * 0.5f + 0.5f;
  }
  lo = make_uint4(pack2(v[0], v[1]), pack2(v[2], a0), pack2(a1, a2), pack2(n0, n1));
  hi = make_uint4(pack2(n2, 0.f), 0u, 0u, 0u);
}

// Generic path: one thread per tile-buffer pixel, any format / strides.
__global__ void __launch_bounds__(256) input_process_kernel(const InputParams p)
{
  const int wd = blockIdx.x * blockDim.x + threadIdx.x;
  const int hd = blockIdx.y;
  if (wd >= p.TW) return;
  const int h = hd - p.tile.hDstBegin, w = wd - p.tile.wDstBegin;
  uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
  if (h >= 0 && h < p.tile.H && w >= 0 && w < p.tile.W)
  {
    const int hs = h + p.tile.hSrcBegin, ws = w + p.tile.wSrcBegin;
    const float scale = p.tf.input_scale_ptr ? *p.tf.input_scale_ptr : p.tf.input_scale;
    const bool has_a = p.albedo.ptr != nullptr, has_n = has_a && p.normal.ptr != nullptr;
    const float3 c = img_get3(p.input, hs, ws);
    const float3 a = has_a ? img_get3(p.albedo, hs, ws) : make_float3(0, 0, 0);
    const float3 n = has_n ? img_get3(p.normal, hs, ws) : make_float3(0, 0, 0);
    input_pixel(p, scale, c, a, n, has_a, has_n, lo, hi);
  }
  uint4* d = reinterpret_cast<uint4*>(p.dst + ((size_t)hd * p.TW + wd) * 16);
  d[0] = lo;
  d[1] = hi;
}

// Fast path: packed fp32 RGB images (pixel stride 12 B, 16-B aligned rows), tile origins and
// widths multiples of 4 pixels. A warp owns a span of 128 consecutive tile-buffer pixels of one row
// and moves it with fully coalesced accesses in both directions:
//   * per image the span is 1536 contiguous bytes = 96 16-byte chunks; lane L loads chunks L, L+32,
//     L+64 (LDG.128, consecutive lanes -> consecutive addresses) and parks them in the warp's
//     shared-memory slice;
//   * lane L then owns pixels L, L+32, L+64, L+96 of the span: it reads their 3 floats back from
//     shared memory (word stride 3 across lanes: conflict-free), converts, and writes each pixel's
//     16 fp16 channels with one 32-byte store -- a warp store instruction covers 1 KiB contiguous.
// Validity is per 4-pixel group (48 B = 3 chunks): a group is either inside the tile or zero.
constexpr int kSpanPx = 128;                       // pixels per warp span
constexpr int kSpanBytes = kSpanPx * 12;           // per image
constexpr int kRowWarps = 8;                       // warps per block = spans per block

__device__ __forceinline__ void st_global_v8(void* ptr, const uint4& lo, const uint4& hi)
{
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :: "l"(ptr), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w)
               : "memory");
}

// chunks L, L+32, L+64 of the span's bytes of one image row (zeros outside the valid groups)
__device__ __forceinline__ void span_ldg(const Img& im, bool present, int hs, int ws0, int lane, int g_lo, int g_hi,
                                         float4 (&q)[3])
{
  const uint8_t* row = im.ptr + (size_t)hs * im.rs + (long long)ws0 * 12; // ws0 may be negative: only valid groups are touched
#pragma unroll
  for (int i = 0; i < 3; ++i)
  {
    const int c = i * 32 + lane;
    const int g = c / 3;
    q[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (present && g >= g_lo && g < g_hi) q[i] = __ldg(reinterpret_cast<const float4*>(row + (size_t)c * 16));
  }
}

__device__ __forceinline__ void span_sts(float* slice, int lane, const float4 (&q)[3])
{
#pragma unroll
  for (int i = 0; i < 3; ++i) reinterpret_cast<float4*>(slice)[i * 32 + lane] = q[i];
}

__global__ void __launch_bounds__(kRowWarps * 32) input_process_rows_kernel(const InputParams p)
{
  __shared__ __align__(16) float stage[kRowWarps][3][kSpanBytes / 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wd0 = (blockIdx.x * kRowWarps + warp) * kSpanPx;   // first tile-buffer pixel of the span
  const int hd = blockIdx.y;
  if (wd0 >= p.TW) return;
  const int h = hd - p.tile.hDstBegin;
  __half* drow = p.dst + ((size_t)hd * p.TW + wd0) * 16;
  const int npx = min(kSpanPx, p.TW - wd0);                     // multiple of 4
  // 4-pixel groups of the span that lie inside the tile: [g_lo, g_hi)
  int g_lo = 0, g_hi = 0;
  if (h >= 0 && h < p.tile.H)
  {
    g_lo = max(0, (p.tile.wDstBegin - wd0) >> 2);
    g_hi = min(npx >> 2, (p.tile.wDstBegin + p.tile.W - wd0) >> 2);
  }
  if (g_hi <= g_lo)
  {
    const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (j * 32 + lane < npx) st_global_v8(drow + (size_t)(j * 32 + lane) * 16, z, z);
    return;
  }
  const int hs = h + p.tile.hSrcBegin;
  const int ws0 = wd0 - p.tile.wDstBegin + p.tile.wSrcBegin;   // image column of the span's first pixel
  const bool has_a = p.albedo.ptr != nullptr, has_n = has_a && p.normal.ptr != nullptr;
  // all nine 16-byte loads are in flight before the first one is consumed
  float4 qc[3], qa[3], qn[3];
  span_ldg(p.input, true, hs, ws0, lane, g_lo, g_hi, qc);
  span_ldg(p.albedo, has_a, hs, ws0, lane, g_lo, g_hi, qa);
  span_ldg(p.normal, has_n, hs, ws0, lane, g_lo, g_hi, qn);
  const float scale = p.tf.input_scale_ptr ? *p.tf.input_scale_ptr : p.tf.input_scale;
  asm volatile("" ::: "memory"); // keep the compiler from sinking loads between the shared-memory stores
  span_sts(stage[warp][0], lane, qc);
  span_sts(stage[warp][1], lane, qa);
  span_sts(stage[warp][2], lane, qn);
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 4; ++j)
  {
    const int px = j * 32 + lane;
    if (px >= npx) continue;
    uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
    const int g = px >> 2;
    if (g >= g_lo && g < g_hi)
    {
      const float* c = &stage[warp][0][px * 3];
      const float* qa3 = &stage[warp][1][px * 3];
      const float* qn3 = &stage[warp][2][px * 3];
      const float3 a = make_float3(qa3[0], qa3[1], qa3[2]), n = make_float3(qn3[0], qn3[1], qn3[2]);
      input_pixel(p, scale, make_float3(c[0], c[1], c[2]), a, n, has_a, has_n, lo, hi);
    }
    st_global_v8(drow + (size_t)px * 16, lo, hi);
  }
}

bool packed_rgb32(const Img& im)
{
  return !im.ptr || (!im.is_half && im.C == 3 && im.ps == 12 && im.rs % 16 == 0 &&
                     reinterpret_cast<uintptr_t>(im.ptr) % 16 == 0);
}

// ------------------------------------------------------------------------------------------------
// Output process
// ------------------------------------------------------------------------------------------------
struct OutputParams
{
  const __half* src;
  int TH, TW, C;
  oidnb200_tile tile;
  Transfer tf;
  int hdr, snorm;
  Img dst;
};

__global__ void __launch_bounds__(256) output_process_kernel(const OutputParams p)
{
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  const int h = blockIdx.y;
  if (w >= p.tile.W) return;
  const __half* s = p.src + ((size_t)(h + p.tile.hSrcBegin) * p.TW + (w + p.tile.wSrcBegin)) * p.C;
  const uint2 raw = *reinterpret_cast<const uint2*>(s); // channels 0..3
  const __half2 h01 = *reinterpret_cast<const __half2*>(&raw.x), h23 = *reinterpret_cast<const __half2*>(&raw.y);
  const float3 v = output_pixel(p.tf, p.hdr, p.snorm, p.dst.C == 1, output_scale(p.tf), __low2float(h01), __high2float(h01), __low2float(h23));
  img_set3(p.dst, h + p.tile.hDstBegin, w + p.tile.wDstBegin, v);
}

// Packed fp32 RGB destination, tile origin/width multiples of 4 pixels. The mirror image of the
// input fast path: a warp owns 128 consecutive pixels of one tile row; lane L converts pixels L,
// L+32, L+64, L+96 (8-byte loads at the tensor's 32-byte pixel pitch: 8 lines per warp load instead
// of 32), parks the 3 floats in the warp's shared-memory slice (word stride 3: conflict-free) and
// the warp writes the span's 1536 bytes as 96 consecutive 16-byte chunks.
__global__ void __launch_bounds__(kRowWarps * 32) output_process_rows_kernel(const OutputParams p)
{
  __shared__ __align__(16) float stage[kRowWarps][kSpanBytes / 4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int w0 = (blockIdx.x * kRowWarps + warp) * kSpanPx;    // first tile pixel of the span
  const int h = blockIdx.y;
  if (w0 >= p.tile.W) return;
  const int npx = min(kSpanPx, p.tile.W - w0);                  // multiple of 4
  const float oscale = output_scale(p.tf);
  const __half* s = p.src + ((size_t)(h + p.tile.hSrcBegin) * p.TW + (w0 + p.tile.wSrcBegin)) * p.C;
  uint2 raw[4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
  {
    const int px = j * 32 + lane;
    raw[j] = make_uint2(0, 0);
    if (px < npx) raw[j] = *reinterpret_cast<const uint2*>(s + (size_t)px * p.C); // channels 0..3
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
  {
    const int px = j * 32 + lane;
    if (px >= npx) continue;
    const __half2 h01 = *reinterpret_cast<const __half2*>(&raw[j].x), h23 = *reinterpret_cast<const __half2*>(&raw[j].y);
    const float3 v = output_pixel(p.tf, p.hdr, p.snorm, p.dst.C == 1, oscale, __low2float(h01), __high2float(h01), __low2float(h23));
    float* q = &stage[warp][px * 3];
    q[0] = v.x; q[1] = v.y; q[2] = v.z;
  }
  __syncwarp();
  uint8_t* d = p.dst.ptr + (size_t)(h + p.tile.hDstBegin) * p.dst.rs + (size_t)(w0 + p.tile.wDstBegin) * 12;
  const int nchunks = npx * 3 / 4;                              // 16-byte chunks of the span
#pragma unroll
  for (int i = 0; i < 3; ++i)
  {
    const int c = i * 32 + lane;
    if (c < nchunks) reinterpret_cast<float4*>(d)[c] = reinterpret_cast<const float4*>(stage[warp])[c];
  }
}

// ------------------------------------------------------------------------------------------------
// Autoexposure in two stream-ordered launches:
//   bins:   each warp reduces whole bins (<=16x16 px: lane = column + 16*(row&1), 8 row pairs) of a
//           rectangle of the bin grid and stores log2(mean luminance) per bin (-inf = bin not
//           counted, L <= 1e-8) into the frame's bin array;
//   reduce: one block folds the complete bin array in a fixed order (fp64) and writes the scale.
// The result is a function of the bin array only, so it does not depend on which GPU computed
// which bins: with a frame sharded by tile every rank fills the bins of its own tiles and the
// arrays are summed (x + 0 = x) before the fold -- bit-identical to one GPU.
// ------------------------------------------------------------------------------------------------
struct AeBinsParams
{
  Img src;
  int nbh, nbw;             // bin grid of the whole image
  int bh0, bh1, bw0, bw1;   // rectangle of bins computed by this launch
  float* bins;              // [nbh * nbw]
  int narrow;               // nbh * H and nbw * W fit 32 bits: the bin bounds take 32-bit divisions
};

constexpr int kAeThreads = 256;
constexpr int kAeReduceThreads = 1024;

__global__ void __launch_bounds__(kAeThreads) autoexposure_bins_kernel(const AeBinsParams p)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = kAeThreads / 32;
  const int rw = p.bw1 - p.bw0;
  const int nbins = (p.bh1 - p.bh0) * rw;
  const int col = lane & 15, rpar = lane >> 4;
  for (int bin = blockIdx.x * nwarps + warp; bin < nbins; bin += gridDim.x * nwarps)
  {
    const int bi = p.bh0 + bin / rw, bj = p.bw0 + bin % rw;
    int h0, h1, w0, w1;
    if (p.narrow)
    {
      // (bin index + 1) x image size < 2^32 (every image below 128K pixels a side): 32-bit divisions -- the four
      // 64-bit ones were a third of this kernel's issue slots
      h0 = (int)((unsigned)bi * (unsigned)p.src.H / (unsigned)p.nbh); h1 = (int)((unsigned)(bi + 1) * (unsigned)p.src.H / (unsigned)p.nbh);
      w0 = (int)((unsigned)bj * (unsigned)p.src.W / (unsigned)p.nbw); w1 = (int)((unsigned)(bj + 1) * (unsigned)p.src.W / (unsigned)p.nbw);
    }
    else
    {
      h0 = (int)((long long)bi * p.src.H / p.nbh); h1 = (int)((long long)(bi + 1) * p.src.H / p.nbh);
      w0 = (int)((long long)bj * p.src.W / p.nbw); w1 = (int)((long long)(bj + 1) * p.src.W / p.nbw);
    }
    float L = 0.f;
    const int w = w0 + col;
    // a bin has at most 16 rows = 8 per lane half: all 8 row loads are issued before the first use
    float3 c[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
    {
      const int h = h0 + rpar + 2 * k;
      c[k] = (w < w1 && h < h1) ? img_get3(p.src, h, w) : make_float3(0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k)
    {
      const float r = clampf(nan_to_zero(c[k].x), 0.f, FLT_MAX), g = clampf(nan_to_zero(c[k].y), 0.f, FLT_MAX),
                  b = clampf(nan_to_zero(c[k].z), 0.f, FLT_MAX);
      L += 0.212671f * r + 0.715160f * g + 0.072169f * b; // core/color.h:169-172
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) L += __shfl_xor_sync(0xffffffffu, L, o);
    L /= (float)((h1 - h0) * (w1 - w0));
    if (lane == 0) p.bins[(size_t)bi * p.nbw + bj] = (L > 1e-8f) ? log2f(L) : -INFINITY;
  }
}

__global__ void __launch_bounds__(kAeReduceThreads) autoexposure_reduce_kernel(const float* __restrict__ bins, int nbins,
                                                                                float* __restrict__ dst)
{
  double s = 0.; int c = 0;
  for (int i = threadIdx.x; i < nbins; i += kAeReduceThreads)
  {
    const float v = bins[i];
    if (v > -INFINITY) { s += (double)v; ++c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
  {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  __shared__ double ss[kAeReduceThreads / 32];
  __shared__ int sc[kAeReduceThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { ss[warp] = s; sc[warp] = c; }
  __syncthreads();
  if (warp == 0)
  {
    s = ss[lane]; c = sc[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if (lane == 0) *dst = c > 0 ? 0.18f / exp2f((float)(s / (double)c)) : 1.f;
  }
}

int autoexposure_grid(int nbins)
{
  const int per_block = kAeThreads / 32;
  int g = (nbins + per_block - 1) / per_block;
  const int cap = 148 * 8;
  return g < 1 ? 1 : (g > cap ? cap : g);
}

// ------------------------------------------------------------------------------------------------
// Image copy, pool, upsample
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) image_copy_kernel(const Img src, const Img dst, int px_bytes)
{
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  const int h = blockIdx.y;
  if (w >= dst.W) return;
  const uint8_t* s = src.ptr + (size_t)h * src.rs + (size_t)w * src.ps;
  uint8_t* d = dst.ptr + (size_t)h * dst.rs + (size_t)w * dst.ps;
  if (src.is_half)
    for (int i = 0; i < px_bytes; i += 2) *reinterpret_cast<uint16_t*>(d + i) = *reinterpret_cast<const uint16_t*>(s + i);
  else
    for (int i = 0; i < px_bytes; i += 4) *reinterpret_cast<uint32_t*>(d + i) = *reinterpret_cast<const uint32_t*>(s + i);
}

__device__ __forceinline__ uint32_t hmax2u(uint32_t a, uint32_t b)
{
  uint32_t r;
  asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}

// one thread per 8 channels (16 B) of one output pixel
__global__ void __launch_bounds__(256) pool_kernel(const uint4* __restrict__ src, int H, int W, int C8, uint4* __restrict__ dst)
{
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int Ho = H / 2, Wo = W / 2;
  if (idx >= (long)Ho * Wo * C8) return;
  const int c = (int)(idx % C8);
  const int x = (int)((idx / C8) % Wo);
  const int y = (int)(idx / ((long)C8 * Wo));
  const uint4 a = src[((size_t)(2 * y) * W + 2 * x) * C8 + c], b = src[((size_t)(2 * y) * W + 2 * x + 1) * C8 + c];
  const uint4 d = src[((size_t)(2 * y + 1) * W + 2 * x) * C8 + c], e = src[((size_t)(2 * y + 1) * W + 2 * x + 1) * C8 + c];
  uint4 r;
  r.x = hmax2u(hmax2u(a.x, b.x), hmax2u(d.x, e.x));
  r.y = hmax2u(hmax2u(a.y, b.y), hmax2u(d.y, e.y));
  r.z = hmax2u(hmax2u(a.z, b.z), hmax2u(d.z, e.z));
  r.w = hmax2u(hmax2u(a.w, b.w), hmax2u(d.w, e.w));
  dst[idx] = r;
}

__global__ void __launch_bounds__(256) upsample_kernel(const uint4* __restrict__ src, int H, int W, int C8, uint4* __restrict__ dst)
{
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int Ho = H * 2, Wo = W * 2;
  if (idx >= (long)Ho * Wo * C8) return;
  const int c = (int)(idx % C8);
  const int x = (int)((idx / C8) % Wo);
  const int y = (int)(idx / ((long)C8 * Wo));
  dst[idx] = src[((size_t)(y >> 1) * W + (x >> 1)) * C8 + c];
}

// ------------------------------------------------------------------------------------------------
// Peer flags: the frame-level hand-shake of tile-sharded execution over several processes / GPUs without a
// collective. A rank that has finished a step (its autoexposure bins are in every peer's bin array; its output
// rectangles are in the frame) stores the frame's sequence number into its slot of every rank's flag array -- peer
// stores over NVLink from one tiny block -- and a rank that needs the step waits until every slot of its OWN array has
// reached the number. Both kernels are one block without shared memory: they co-reside with a persistent conv CTA
// (which a collective's kernel, with its shared memory, cannot), so the other frame in flight keeps every SM.
// ------------------------------------------------------------------------------------------------
struct FlagTargets
{
  unsigned int* slot[16];   // this rank's slot in every rank's flag array (own array included)
  int n;
};

__global__ void __launch_bounds__(32) flag_signal_kernel(const FlagTargets t, unsigned int value)
{
  // stream order has completed this rank's copies; make them visible system-wide before the flag
  __threadfence_system();
  if ((int)threadIdx.x < t.n)
  {
    volatile unsigned int* d = t.slot[threadIdx.x];
    *d = value;
  }
  __threadfence_system();
}

__global__ void __launch_bounds__(32) flag_wait_kernel(const unsigned int* flags, int n, unsigned int value, unsigned long long timeout_ns)
{
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  if ((int)threadIdx.x < n)
  {
    const volatile unsigned int* f = flags + threadIdx.x;
    // sequence numbers only grow; compare as a signed difference so a 32-bit wrap does not stall
    while ((int)(*f - value) < 0)
    {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (t - t0 > timeout_ns)
      {
        printf("[oidn_b200] peer flag timeout: slot %d holds %u, waiting for %u\n", (int)threadIdx.x, *f, value);
        __trap();
      }
      __nanosleep(200);
    }
  }
  __threadfence_system();
}

int check_launch(const char* what)
{
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
  {
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

} // namespace
} // namespace oidnb200

using namespace oidnb200;

extern "C" {

int oidnb200_input_process_launch(const oidnb200_image* color, const oidnb200_image* albedo,
                                  const oidnb200_image* normal, const oidnb200_tile* tile,
                                  const oidnb200_transfer* tf, int hdr, int snorm, void* dst, int TH, int TW,
                                  int C, oidnb200_stream stream)
{
  InputParams p;
  if (!make_img(color, p.input) || !make_img(albedo, p.albedo) || !make_img(normal, p.normal) || !p.input.ptr)
  {
    set_error("input_process: missing main input image or unsupported image format");
    return OIDNB200_ERR_INVALID;
  }
  if (!tile || !dst || C != 16 || TH <= 0 || TW <= 0 || !make_transfer(tf, p.tf))
  {
    set_error("input_process: bad arguments (the network input tensor has 16 channels)");
    return OIDNB200_ERR_INVALID;
  }
  if (tile->H < 0 || tile->W < 0 || tile->hDstBegin < 0 || tile->wDstBegin < 0 ||
      tile->hDstBegin + tile->H > TH || tile->wDstBegin + tile->W > TW || tile->hSrcBegin < 0 ||
      tile->wSrcBegin < 0 || tile->hSrcBegin + tile->H > p.input.H || tile->wSrcBegin + tile->W > p.input.W)
  {
    set_error("input_process: tile outside the image or the tile buffer");
    return OIDNB200_ERR_INVALID;
  }
  p.tile = *tile; p.hdr = hdr; p.snorm = snorm;
  p.dst = static_cast<__half*>(dst); p.TH = TH; p.TW = TW; p.C = C;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool vec = packed_rgb32(p.input) && packed_rgb32(p.albedo) && packed_rgb32(p.normal) && TW % 4 == 0 &&
                   tile->W % 4 == 0 && tile->wDstBegin % 4 == 0 && tile->wSrcBegin % 4 == 0;
  if (vec)
  {
    dim3 grid((TW + kSpanPx * kRowWarps - 1) / (kSpanPx * kRowWarps), TH);
    input_process_rows_kernel<<<grid, kRowWarps * 32, 0, st>>>(p);
  }
  else
  {
    const int threads = 256;
    dim3 grid((TW + threads - 1) / threads, TH);
    input_process_kernel<<<grid, threads, 0, st>>>(p);
  }
  return check_launch("input_process");
}

int oidnb200_output_process_launch(const void* src, int TH, int TW, int C, const oidnb200_tile* tile,
                                   const oidnb200_transfer* tf, int hdr, int snorm, const oidnb200_image* dst,
                                   oidnb200_stream stream)
{
  OutputParams p;
  if (!src || !tile || !make_img(dst, p.dst) || !p.dst.ptr || !make_transfer(tf, p.tf) || C < 4 || C % 4)
  {
    set_error("output_process: bad arguments");
    return OIDNB200_ERR_INVALID;
  }
  if (tile->H < 0 || tile->W < 0 || tile->hSrcBegin < 0 || tile->wSrcBegin < 0 || tile->hSrcBegin + tile->H > TH ||
      tile->wSrcBegin + tile->W > TW || tile->hDstBegin < 0 || tile->wDstBegin < 0 ||
      tile->hDstBegin + tile->H > p.dst.H || tile->wDstBegin + tile->W > p.dst.W)
  {
    set_error("output_process: tile outside the tensor or the image");
    return OIDNB200_ERR_INVALID;
  }
  if (tile->H == 0 || tile->W == 0) return 0;
  p.src = static_cast<const __half*>(src); p.TH = TH; p.TW = TW; p.C = C;
  p.tile = *tile; p.hdr = hdr; p.snorm = snorm;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool vec = packed_rgb32(p.dst) && tile->W % 4 == 0 && tile->wDstBegin % 4 == 0;
  if (vec)
  {
    dim3 grid((tile->W + kSpanPx * kRowWarps - 1) / (kSpanPx * kRowWarps), tile->H);
    output_process_rows_kernel<<<grid, kRowWarps * 32, 0, st>>>(p);
  }
  else
  {
    const int threads = 256;
    dim3 grid((tile->W + threads - 1) / threads, tile->H);
    output_process_kernel<<<grid, threads, 0, st>>>(p);
  }
  return check_launch("output_process");
}

void oidnb200_autoexposure_bin_grid(int H, int W, int* num_bins_h, int* num_bins_w)
{
  // core/autoexposure.h:20-24: bins of at most 16x16 pixels
  if (num_bins_h) *num_bins_h = (H + 15) / 16;
  if (num_bins_w) *num_bins_w = (W + 15) / 16;
}

size_t oidnb200_autoexposure_scratch_bytes(int H, int W)
{
  const size_t nbins = (size_t)((H + 15) / 16) * ((W + 15) / 16);
  return (nbins * sizeof(float) + 255) / 256 * 256; // the bin array
}

int oidnb200_autoexposure_bins_launch(const oidnb200_image* src, int bin_h0, int bin_h1, int bin_w0, int bin_w1,
                                      float* bins, oidnb200_stream stream)
{
  AeBinsParams p;
  if (!make_img(src, p.src) || !p.src.ptr || !bins || p.src.H <= 0 || p.src.W <= 0)
  {
    set_error("autoexposure: bad arguments");
    return OIDNB200_ERR_INVALID;
  }
  p.nbh = (p.src.H + 15) / 16; p.nbw = (p.src.W + 15) / 16;
  if (bin_h0 < 0 || bin_w0 < 0 || bin_h1 > p.nbh || bin_w1 > p.nbw || bin_h0 > bin_h1 || bin_w0 > bin_w1)
  {
    set_error("autoexposure: bin rectangle outside the image's bin grid");
    return OIDNB200_ERR_INVALID;
  }
  p.bh0 = bin_h0; p.bh1 = bin_h1; p.bw0 = bin_w0; p.bw1 = bin_w1;
  p.bins = bins;
  p.narrow = ((unsigned long long)p.nbh * p.src.H < (1ull << 32) && (unsigned long long)p.nbw * p.src.W < (1ull << 32)) ? 1 : 0;
  const int n = (bin_h1 - bin_h0) * (bin_w1 - bin_w0);
  if (n == 0) return 0;
  autoexposure_bins_kernel<<<autoexposure_grid(n), kAeThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return check_launch("autoexposure bins");
}

int oidnb200_autoexposure_reduce_launch(const float* bins, int num_bins, float* dst, oidnb200_stream stream)
{
  if (!bins || !dst || num_bins <= 0)
  {
    set_error("autoexposure: bad arguments");
    return OIDNB200_ERR_INVALID;
  }
  autoexposure_reduce_kernel<<<1, kAeReduceThreads, 0, static_cast<cudaStream_t>(stream)>>>(bins, num_bins, dst);
  return check_launch("autoexposure reduce");
}

int oidnb200_autoexposure_launch(const oidnb200_image* src, void* scratch, float* dst, oidnb200_stream stream)
{
  if (!src || !scratch || !dst)
  {
    set_error("autoexposure: bad arguments");
    return OIDNB200_ERR_INVALID;
  }
  const int nbh = (src->H + 15) / 16, nbw = (src->W + 15) / 16;
  const int rc = oidnb200_autoexposure_bins_launch(src, 0, nbh, 0, nbw, static_cast<float*>(scratch), stream);
  if (rc) return rc;
  return oidnb200_autoexposure_reduce_launch(static_cast<float*>(scratch), nbh * nbw, dst, stream);
}

int oidnb200_image_copy_launch(const oidnb200_image* src, const oidnb200_image* dst, oidnb200_stream stream)
{
  Img s, d;
  if (!make_img(src, s) || !make_img(dst, d) || !s.ptr || !d.ptr || s.W != d.W || s.H != d.H ||
      src->format != dst->format)
  {
    set_error("image_copy: images must be set and have the same size and format");
    return OIDNB200_ERR_INVALID;
  }
  if (d.W == 0 || d.H == 0) return 0;
  const int threads = 256;
  dim3 grid((d.W + threads - 1) / threads, d.H);
  image_copy_kernel<<<grid, threads, 0, static_cast<cudaStream_t>(stream)>>>(s, d, s.C * (s.is_half ? 2 : 4));
  return check_launch("image_copy");
}

int oidnb200_flag_signal_launch(void* const* slots, int n, unsigned int value, oidnb200_stream stream)
{
  if (!slots || n < 1 || n > 16)
  {
    set_error("flag_signal: 1..16 slots");
    return OIDNB200_ERR_INVALID;
  }
  FlagTargets t;
  t.n = n;
  for (int i = 0; i < 16; ++i) t.slot[i] = i < n ? static_cast<unsigned int*>(slots[i]) : nullptr;
  flag_signal_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(t, value);
  return check_launch("flag_signal");
}

int oidnb200_flag_wait_launch(const void* flags, int n, unsigned int value, double timeout_s, oidnb200_stream stream)
{
  if (!flags || n < 1 || n > 32)
  {
    set_error("flag_wait: 1..32 slots");
    return OIDNB200_ERR_INVALID;
  }
  const unsigned long long ns = (unsigned long long)((timeout_s > 0 ? timeout_s : 10.0) * 1e9);
  flag_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const unsigned int*>(flags), n, value, ns);
  return check_launch("flag_wait");
}

int oidnb200_pool_launch(const void* src, int H, int W, int C, void* dst, oidnb200_stream stream)
{
  if (!src || !dst || H <= 0 || W <= 0 || (H & 1) || (W & 1) || C <= 0 || C % 8)
  {
    set_error("pool: needs even H, W and C a multiple of 8");
    return OIDNB200_ERR_INVALID;
  }
  const long n = (long)(H / 2) * (W / 2) * (C / 8);
  pool_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
    static_cast<const uint4*>(src), H, W, C / 8, static_cast<uint4*>(dst));
  return check_launch("pool");
}

int oidnb200_upsample_launch(const void* src, int H, int W, int C, void* dst, oidnb200_stream stream)
{
  if (!src || !dst || H <= 0 || W <= 0 || C <= 0 || C % 8)
  {
    set_error("upsample: needs C a multiple of 8");
    return OIDNB200_ERR_INVALID;
  }
  const long n = (long)(H * 2) * (W * 2) * (C / 8);
  upsample_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
    static_cast<const uint4*>(src), H, W, C / 8, static_cast<uint4*>(dst));
  return check_launch("upsample");
}

} // extern "C"
