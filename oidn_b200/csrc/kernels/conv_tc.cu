// 3x3 convolution (pad 1, stride 1, cross-correlation) as an implicit GEMM on tcgen05 / TMEM.
//
// Replaces the reference's CutlassConv (devices/cuda/cutlass_conv.h:150-278) plus its separate
// pool / upsample / concat passes (devices/gpu/gpu_pool.h, gpu_upsample.h,
// core/concat_conv_hwc.cpp:27-31). Semantics follow core/conv.cpp:8-60 and the CPU kernel
// devices/cpu/cpu_conv.ispc:34-127:
//   dst[o,y,x] = act(bias[o] + sum_{i,kh,kw} W[o,i,kh,kw] * src[i,y+kh-1,x+kw-1]), zero padding,
//   fp32 accumulation; optional 2x2 max-pool of the result; src may be the channel concat of two
//   tensors (read in place, never materialised) and src1 may be a half-resolution tensor that is
//   nearest-upsampled on the fly by the loader (a TMA tensor map with a stride-0 "dup" axis).
//
// Mapping (one persistent CTA per SM, warp specialised):
//   * M tile = 128 consecutive pixels of one image row ("strip" of width 128); a work item is a
//     strip x RC rows. Input rows stream through a shared-memory ring one (row, K-chunk) at a
//     time: box = cc channels x 130 pixels (1-px halo each side, TMA zero-fills out of bounds).
//   * The three horizontal taps are three *shifted views* of the same staged row: the UMMA
//     descriptor start address moves by one pixel (= one swizzled row of 32/64/128 B).
//   * The three vertical taps are stacked along N: input row r feeds output rows r+1, r, r-1
//     (kh = 0,1,2), whose fp32 accumulators sit in adjacent TMEM column blocks of a ring, so one
//     tcgen05.mma of N = 3*CoutG covers them. This keeps N large (A-operand smem reads are the
//     limiter for N < 128) and every input row is staged exactly once per item.
//   * Weights for the CTA's output-channel group stay resident in shared memory.
//   * A single thread can issue tcgen05.mma only so fast (a few instructions per MMA at
//     single-warp latency), so a CTA runs up to two independent *streams*: each stream has its
//     own TMA warp, MMA warp, epilogue warpgroup, A ring, TMEM half and staging buffer, works on
//     its own items, and shares the resident weights. With one stream (wide CoutG) the two
//     epilogue warpgroups drain alternate rows of that stream instead.
//   * Epilogue: tcgen05.ld -> +bias -> fp16 -> ReLU -> (2x2 max-pool via a second accumulator +
//     shuffle) -> swizzled smem staging -> TMA store (clipped at the tensor edge).
//
// Warp roles: 0 = TMA stream 0, 1 = MMA stream 0 (+TMEM alloc), 2 = TMA stream 1, 3 = MMA stream 1,
// 4-7 = epilogue warpgroup 0, 8-11 = epilogue warpgroup 1.
#include "conv_common.h"
#include "ptx.cuh"
#include "transfer.cuh"
#include "conv_util.cuh"
#include <cuda_fp16.h>

namespace oidnb200 {

using namespace ptx;

// Optional wait-time tracing (build with -DOIDN_B200_TRACE, tools/probe_conv --trace): every role
// accumulates the cycles it spends blocked at each barrier; lane 0 adds them to p.trace[warp][tag].
#ifdef OIDN_B200_TRACE
#define TRACE_DECL long long tr[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; const long long tr_start = clock64(); long long tr_t = 0;
#define MBAR_WAIT(bar, par, tag) do { const long long t0_ = clock64(); mbar_wait(bar, par, tag); tr[tag] += clock64() - t0_; } while (0)
#define TRACE_WRAP(tag, stmt) do { const long long t0_ = clock64(); stmt; tr[tag] += clock64() - t0_; } while (0)
#define TRACE_FLUSH() do { if (p.trace && lane == 0 && warp < 12) { tr[0] = clock64() - tr_start; \
  for (int i_ = 0; i_ < 16; ++i_) atomicAdd(&p.trace[warp * 16 + i_], (unsigned long long)tr[i_]); } } while (0)
#define TRACE_ARGS , tr, tr_t
#define TRACE_PARAMS , long long* tr, long long& tr_t
#define TRACE_BEGIN() do { tr_t = clock64(); } while (0)
#define TRACE_END(tag) do { const long long t1_ = clock64(); tr[tag] += t1_ - tr_t; tr_t = t1_; } while (0)
#else
#define TRACE_ARGS
#define TRACE_PARAMS
#define TRACE_BEGIN()
#define TRACE_END(tag)
#define TRACE_DECL
#define MBAR_WAIT(bar, par, tag) mbar_wait(bar, par, tag)
#define TRACE_WRAP(tag, stmt) stmt
#define TRACE_FLUSH()
#endif

namespace {

struct SmemLayout
{
  // byte offsets from the 1024-aligned base; [stream][index] arrays
  static constexpr uint32_t full_a     = 0;                                  // kMaxStreams x kMaxStages x 8
  static constexpr uint32_t empty_a    = full_a + kMaxStreams * 8 * kMaxStages;
  static constexpr uint32_t w_full     = empty_a + kMaxStreams * 8 * kMaxStages;
  static constexpr uint32_t tmem_full  = w_full + 8;                         // kMaxStreams x kMaxSlots x 8
  static constexpr uint32_t tmem_empty = tmem_full + kMaxStreams * 8 * kMaxSlots;
  static constexpr uint32_t tmem_ptr   = tmem_empty + kMaxStreams * 8 * kMaxSlots;
  static constexpr uint32_t bias       = 3072;                               // 128 floats
  static constexpr uint32_t a_ring     = kSmemHeader;
};
static_assert(SmemLayout::tmem_ptr + 4 <= SmemLayout::bias && SmemLayout::bias + 512 <= kSmemHeader, "barrier block overflows");

struct Item
{
  int x0, y0, y1;
};

__device__ __forceinline__ Item get_item(const ConvKernelParams& p, int item)
{
  Item it;
  const int strip = item % p.nstrips;
  const int rc    = item / p.nstrips;
  it.x0 = strip * kStripW;
  it.y0 = rc * p.RC;
  it.y1 = min(p.H, it.y0 + p.RC) - 1;
  return it;
}

// All tcgen05.mma of one staged (row, K-chunk): 3 horizontal taps x NK k-steps x (1 or 2) runs,
// fully unrolled so that each MMA costs two descriptor adds plus the issue.
template <int NK, bool TWO>
__device__ __forceinline__ void issue_chunk(uint32_t hi, uint32_t a_lo, uint32_t b_lo,
                                            uint32_t bblk16, uint32_t d0, uint32_t idesc0, uint32_t rb0,
                                            uint32_t d1, uint32_t idesc1, uint32_t rb1, bool skip_first)
{
  constexpr uint32_t row16 = NK * 2; // bytes of one staged pixel / 16
#pragma unroll
  for (int kw = 0; kw < 3; ++kw)
  {
#pragma unroll
    for (int j = 0; j < NK; ++j)
    {
      if (kw == 0 && j == 0 && skip_first) continue;
      const uint64_t adesc = make_desc(hi, a_lo + kw * row16 + 2 * j);
      umma_f16(d0, adesc, make_desc(hi, b_lo + kw * bblk16 + rb0 + 2 * j), idesc0, 1u);
      if (TWO)
        umma_f16(d1, adesc, make_desc(hi, b_lo + kw * bblk16 + rb1 + 2 * j), idesc1, 1u);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Epilogue
// ------------------------------------------------------------------------------------------------
struct EpiCtx
{
  uint32_t sbase, tmem_base, b_region;
  uint8_t* sgen;
  int warp, lane, group, cta, nctas;
  int epi_warp0;   // first epilogue warp of the CTA (4, or 8 in the four-stream kernel)
};

// Bias add (fp32, packed pairs) + fp16 rounding (+ReLU) of W accumulator columns; W = 16 or 32.
template <int W>
__device__ __forceinline__ void bias_cvt(const uint32_t (&v)[W], const float2* bias2, bool relu, uint32_t (&h)[W / 2])
{
#pragma unroll
  for (int i = 0; i < W / 2; ++i)
  {
    const float2 s = __fadd2_rn(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), bias2[i]);
    h[i] = relu ? pack_half2_relu(s.x, s.y) : pack_half2(s.x, s.y);
  }
}

// The same with the bias read from the kernel parameters (constant bank; p.bias_in_params): `c` = first channel.
template <int W>
__device__ __forceinline__ void bias_cvt_c(const uint32_t (&v)[W], const ConvKernelParams& p, int c, bool relu, uint32_t (&h)[W / 2])
{
#pragma unroll
  for (int i = 0; i < W / 2; ++i)
  {
    const float2 s = __fadd2_rn(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])),
                                make_float2(p.bias_c[c + 2 * i], p.bias_c[c + 2 * i + 1]));
    h[i] = relu ? pack_half2_relu(s.x, s.y) : pack_half2(s.x, s.y);
  }
}

// Two warpgroups of 128 threads = the 128 TMEM lanes (pixels) of an accumulator. With two streams,
// warpgroup g drains every row of stream g; with one stream the two warpgroups drain alternate
// rows (row pairs when pooling). Every WARP is its own store pipeline: it owns the 32 pixels of its
// TMEM lane quarter, stages them in its own (double-buffered) swizzled smem slice and issues its
// own TMA store, so a row costs no CTA-level barrier:
//   TMEM -> registers -> (max of the two rows) -> +bias (fp32) -> fp16 (+ReLU) -> (max of the pixel
//   pair) -> smem slice -> TMA store.
// max commutes with the monotonic "+bias, round" so pooling after rounding is bit-identical to
// pooling the rounded full-resolution tensor (what the reference's separate pool pass does).
// WIDE: 32-column TMEM loads where a piece allows (fewer, larger loads); the four-stream kernel
// (80 registers per thread) uses 16-column loads only.
template <int NB, bool POOL, bool WIDE>
__device__ __forceinline__ void epilogue(const ConvKernelParams& p, const EpiCtx& ec TRACE_PARAMS)
{
  constexpr int CoutG = NB * 16;
  constexpr int NP    = out_piece_count(NB);
  constexpr int PROWS = POOL ? 16 : 32;       // staging rows (pixels) of one warp
  const int warp = ec.warp, lane = ec.lane;
  const int NST  = p.nstreams;
  const int R    = p.R;
  const int wg   = (warp - ec.epi_warp0) >> 2; // epilogue warpgroup: 0..1 (0..3 with four streams)
  const int st   = (NST >= 2) ? wg : 0;       // stream drained by this warpgroup
  const bool alternate = (NST == 1);
  const int vcta = ec.cta * NST + st, nv = ec.nctas * NST;
  const int nitems = p.nstrips * p.nrowchunks;
  const int q    = warp & 3;                  // TMEM lane quarter this warp may access
  const bool relu = p.relu != 0;
  const bool direct = p.direct_store != 0;   // registers -> global memory, no staging / TMA store
  const bool cbias = p.bias_in_params != 0;  // bias from the constant bank instead of shared memory
  const int nbuf = p.out_nbuf;
  const uint32_t wregion = ec.b_region + p.b_bytes + (uint32_t)((wg * 4 + q) * nbuf) * p.out_buf_bytes;
  const int gch0 = ec.group * CoutG;           // first output channel of this CTA's group
  __half* const gout = static_cast<__half*>(p.out_ptr) + gch0;
  const uint32_t tfull  = ec.sbase + SmemLayout::tmem_full + 8 * st * kMaxSlots;
  const uint32_t tempty = ec.sbase + SmemLayout::tmem_empty + 8 * st * kMaxSlots;
  const int spix = POOL ? (lane >> 1) : lane; // staging row of this thread's pixel inside the warp slice
  const bool writer = !POOL || ((lane & 1) == 0);
  const uint32_t lane_base = ec.tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)st * (uint32_t)R * CoutG;
  const float2* bias_s = reinterpret_cast<const float2*>(ec.sgen + SmemLayout::bias);

  // this thread's staging row inside each piece and its swizzle term (16-B chunk ^= 128-B line mod chunks)
  uint32_t prow[NP], pxor[NP];
#pragma unroll
  for (int oc = 0; oc < NP; ++oc)
  {
    constexpr int dummy = 0; (void)dummy;
    const uint32_t rowb = (uint32_t)out_piece_cc(NB, oc) * 2u;
    const uint32_t off = (uint32_t)spix * rowb;
    prow[oc] = out_piece_off(NB, oc, PROWS) + off;
    pxor[oc] = ((off >> 7) & ((rowb >> 4) - 1u)) << 4;
  }
  // narrow groups keep their bias in registers
  constexpr bool kBiasRegs = (NB <= 2);
  float2 breg[kBiasRegs ? NB * 8 : 1];
  if (kBiasRegs)
  {
#pragma unroll
    for (int i = 0; i < NB * 8; ++i) breg[i] = bias_s[i];
  }

  uint32_t a_mod = 0, a_par = 0;
  uint32_t buf = 0, rown = 0;
  for (int item = vcta; item < nitems; item += nv)
  {
    const Item it = get_item(p, item);
    const int xo = (POOL ? (it.x0 >> 1) : it.x0) + q * PROWS;
    const bool gwriter = writer && (xo + spix) < p.out_W;   // direct stores clip at the tensor edge themselves
    uint32_t y_mod = a_mod, y_par = a_par;
    for (int y = it.y0; y <= it.y1; y += (POOL ? 2 : 1), ++rown)
    {
      const uint32_t slot0 = (R - 1) - y_mod;
      const uint32_t par0 = y_par;
      if (++y_mod == (uint32_t)R) { y_mod = 0; y_par ^= 1; }
      uint32_t slot1 = 0, par1 = 0;
      if (POOL)
      {
        slot1 = (R - 1) - y_mod; par1 = y_par;
        if (++y_mod == (uint32_t)R) { y_mod = 0; y_par ^= 1; }
      }
      if (alternate && (int)(rown & 1) != wg) continue;

      MBAR_WAIT(tfull + 8 * slot0, par0, 5);
      if (POOL)
        MBAR_WAIT(tfull + 8 * slot1, par1, 6);
      tc_fence_after();
      // this warp's staging slice `buf` must have been read out by the TMA store that used it last
      if (!direct)
      {
        if (lane == 0)
        {
          TRACE_WRAP(7, if (nbuf == 2) bulk_wait_read<1>(); else bulk_wait_read<0>());
        }
        __syncwarp();
      }
      TRACE_BEGIN();
      __half* const gpix = gout + ((size_t)(POOL ? (y >> 1) : y) * p.out_W + (xo + spix)) * p.CoutPad;
      const uint32_t t0 = lane_base + slot0 * CoutG;
      const uint32_t t1 = lane_base + slot1 * CoutG;
      const uint32_t stage_out = wregion + buf * p.out_buf_bytes;
#pragma unroll
      for (int oc = 0; oc < NP; ++oc)
      {
        constexpr int dummy2 = 0; (void)dummy2;
        const int c0 = out_piece_c0(NB, oc), ccw = out_piece_cc(NB, oc);
        const uint32_t rowaddr = stage_out + prow[oc];
        const uint32_t xr = pxor[oc];
#pragma unroll
        for (int j = 0; j < ccw; j += (WIDE ? 32 : 16))
        {
          if (WIDE && ccw - j >= 32)
          {
            uint32_t v[32], h[16];
            tmem_ld32(t0 + c0 + j, v);
            if (POOL)
            {
              uint32_t w[32];
              tmem_ld32(t1 + c0 + j, w);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(fmaxf(__uint_as_float(v[i]), __uint_as_float(w[i])));
            }
            else
              tmem_ld_wait();
            TRACE_END(8);
            if (!kBiasRegs && cbias) bias_cvt_c<32>(v, p, gch0 + c0 + j, relu, h);
            else bias_cvt<32>(v, kBiasRegs ? &breg[(c0 + j) / 2] : bias_s + (c0 + j) / 2, relu, h);
            if (POOL)
            {
#pragma unroll
              for (int i = 0; i < 16; ++i) h[i] = max_half2(h[i], __shfl_xor_sync(0xffffffffu, h[i], 1));
            }
            if (direct)
            {
              // the last group may reach past the tensor's channels (CoutAlloc > CoutPad): clip per 16
              if (gwriter)
              {
                if (gch0 + c0 + j < p.CoutPad)      st_global_32B(gpix + c0 + j, &h[0]);
                if (gch0 + c0 + j + 16 < p.CoutPad) st_global_32B(gpix + c0 + j + 16, &h[8]);
              }
            }
            else if (writer)
            {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                st_shared_v4(rowaddr + (((uint32_t)(j * 2 + k * 16)) ^ xr), h[4 * k], h[4 * k + 1], h[4 * k + 2], h[4 * k + 3]);
            }
            TRACE_END(9);
          }
          else
          {
            uint32_t v[16], h[8];
            tmem_ld16(t0 + c0 + j, v);
            if (POOL)
            {
              uint32_t w[16];
              tmem_ld16(t1 + c0 + j, w);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(fmaxf(__uint_as_float(v[i]), __uint_as_float(w[i])));
            }
            else
              tmem_ld_wait();
            TRACE_END(8);
            if (!kBiasRegs && cbias) bias_cvt_c<16>(v, p, gch0 + c0 + j, relu, h);
            else bias_cvt<16>(v, kBiasRegs ? &breg[(c0 + j) / 2] : bias_s + (c0 + j) / 2, relu, h);
            if (POOL)
            {
#pragma unroll
              for (int i = 0; i < 8; ++i) h[i] = max_half2(h[i], __shfl_xor_sync(0xffffffffu, h[i], 1));
            }
            if (direct)
            {
              if (gwriter && gch0 + c0 + j < p.CoutPad) st_global_32B(gpix + c0 + j, &h[0]);
            }
            else if (writer)
            {
#pragma unroll
              for (int k = 0; k < 2; ++k)
                st_shared_v4(rowaddr + (((uint32_t)(j * 2 + k * 16)) ^ xr), h[4 * k], h[4 * k + 1], h[4 * k + 2], h[4 * k + 3]);
            }
            TRACE_END(9);
          }
        }
      }
      // Release the accumulator slot(s) back to the MMA issuer (one arrive per warp), make the
      // generic-proxy smem writes visible to the TMA engine, then store this warp's pixels.
      tc_fence_before();
      if (!direct) fence_proxy_async();
      __syncwarp();
      TRACE_END(10);
      if (lane == 0)
      {
        mbar_arrive(tempty + 8 * slot0);
        if (POOL)
          mbar_arrive(tempty + 8 * slot1);
        if (!direct)
        {
          const int yo = POOL ? (y >> 1) : y;
#pragma unroll
          for (int oc = 0; oc < NP; ++oc)
            tma_store_3d(&p.omap[oc], stage_out + out_piece_off(NB, oc, PROWS), ec.group * CoutG + out_piece_c0(NB, oc), xo, yo);
          bulk_commit();
        }
      }
      TRACE_END(11);
      if (++buf == (uint32_t)nbuf) buf = 0;
    }
    const uint32_t tot = a_mod + (uint32_t)(it.y1 - it.y0 + 1);
    a_par ^= (tot / R) & 1;
    a_mod = tot % R;
  }
  if (lane == 0) bulk_wait_read<0>();
  __syncwarp();
}


// Epilogue of the network's last convolution with the output process fused in (p.fo.enabled;
// CoutG = 16, no pooling): same accumulator hand-off as epilogue<1, false>, but a thread reads only
// channels 0..3 of its pixel, rounds them to fp16 exactly as the stored tensor would have been
// (so the result is bit-identical to conv + separate output process), applies the output-process
// math and writes the pixel's 3 floats to the image. Consecutive lanes write consecutive 12-byte
// pixels: a warp store covers 384 contiguous bytes. No staging, no TMA store.
__device__ __forceinline__ void epilogue_fused_output(const ConvKernelParams& p, const EpiCtx& ec TRACE_PARAMS)
{
  constexpr int CoutG = 16;
  const int warp = ec.warp, lane = ec.lane;
  const int NST  = p.nstreams;
  const int R    = p.R;
  const int wg   = (warp - ec.epi_warp0) >> 2;
  const int st   = (NST >= 2) ? wg : 0;
  const bool alternate = (NST == 1);
  const int vcta = ec.cta * NST + st, nv = ec.nctas * NST;
  const int nitems = p.nstrips * p.nrowchunks;
  const int q    = warp & 3;
  const bool relu = p.relu != 0;
  const uint32_t tfull  = ec.sbase + SmemLayout::tmem_full + 8 * st * kMaxSlots;
  const uint32_t tempty = ec.sbase + SmemLayout::tmem_empty + 8 * st * kMaxSlots;
  const uint32_t lane_base = ec.tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)st * (uint32_t)R * CoutG;
  const float* bias_s = reinterpret_cast<const float*>(ec.sgen + SmemLayout::bias);
  const float b0 = bias_s[0], b1 = bias_s[1], b2 = bias_s[2];

  const FusedOutput& fo = p.fo;
  Transfer tf;
  tf.type = fo.tf_type; tf.norm = fo.norm; tf.rcp_norm = fo.rcp_norm;
  tf.input_scale = fo.input_scale; tf.input_scale_ptr = fo.input_scale_ptr;
  // The scale may be the autoexposure result written by an earlier grid of the stream: order this
  // grid's first read of it after the prerequisite grids (programmatic dependent launch).
  pdl_wait();
  const float oscale = output_scale(tf);
  const bool hdr = fo.hdr != 0, snorm = fo.snorm != 0;

  uint32_t a_mod = 0, a_par = 0;
  uint32_t rown = 0;
  for (int item = vcta; item < nitems; item += nv)
  {
    const Item it = get_item(p, item);
    const int x  = it.x0 + q * 32 + lane;                 // this thread's pixel column in the tensor
    const int xi = x - fo.wSrc;                           // column inside the output rectangle
    const bool xok = x < p.W && xi >= 0 && xi < fo.W;
    uint32_t y_mod = a_mod, y_par = a_par;
    for (int y = it.y0; y <= it.y1; ++y, ++rown)
    {
      const uint32_t slot0 = (R - 1) - y_mod;
      const uint32_t par0 = y_par;
      if (++y_mod == (uint32_t)R) { y_mod = 0; y_par ^= 1; }
      if (alternate && (int)(rown & 1) != wg) continue;

      MBAR_WAIT(tfull + 8 * slot0, par0, 5);
      tc_fence_after();
      uint32_t v[4];
      tmem_ld4(lane_base + slot0 * CoutG, v);
      tmem_ld_wait();
      // the accumulator slot goes back to the MMA issuer before the (long) per-pixel math
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty + 8 * slot0);

      const int yi = y - fo.hSrc;
      if (xok && yi >= 0 && yi < fo.H)
      {
        const float s0 = __uint_as_float(v[0]) + b0, s1 = __uint_as_float(v[1]) + b1, s2 = __uint_as_float(v[2]) + b2;
        const uint32_t h01 = relu ? pack_half2_relu(s0, s1) : pack_half2(s0, s1);
        const uint32_t h2x = relu ? pack_half2_relu(s2, 0.f) : pack_half2(s2, 0.f);
        const __half2 q01 = *reinterpret_cast<const __half2*>(&h01), q2x = *reinterpret_cast<const __half2*>(&h2x);
        const float3 o = output_pixel(tf, hdr, snorm, false, oscale, __low2float(q01), __high2float(q01), __low2float(q2x));
        float* d = reinterpret_cast<float*>(fo.ptr + (long long)(yi + fo.hDst) * fo.rs) + (size_t)(xi + fo.wDst) * 3;
        d[0] = o.x; d[1] = o.y; d[2] = o.z;
      }
    }
    const uint32_t tot = a_mod + (uint32_t)(it.y1 - it.y0 + 1);
    a_par ^= (tot / R) & 1;
    a_mod = tot % R;
  }
}

} // namespace

// NSTMAX = 2: 384 threads (warps 0-3 producers, 4-11 epilogue); NSTMAX = 4: 768 threads (warps 0-7
// producers, 8-23 epilogue), launched only with p.nstreams == 4 (CoutG <= 32: four accumulator rings
// of >= 4 slots fit the 512 TMEM columns). More streams = more threads issuing tcgen05.mma: the narrow
// layers are bound by how fast ONE thread can issue (a 128x96x16 MMA is 48 tensor-pipe cycles).
template <int NSTMAX>
__device__ __forceinline__ void conv3x3_tc_body(const ConvKernelParams& p)
{
  constexpr int kProducerWarps = 2 * NSTMAX;
  constexpr int kThreads = (kProducerWarps + 4 * NSTMAX) * 32;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));

  // warp index through a shuffle: the compiler can then prove it warp-uniform and keep the role
  // loops' state in uniform registers
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  TRACE_DECL

  const int NST     = p.nstreams;               // 1 or 2
  const int group   = blockIdx.x % p.ngroups;
  const int cta     = blockIdx.x / p.ngroups;
  const int nctas   = gridDim.x / p.ngroups;
  const int nitems  = p.nstrips * p.nrowchunks;
  const int NS      = p.nstages;                // A stages per stream
  const int R       = p.R;                      // accumulator ring slots per stream
  const uint32_t ring_bytes = p.ring_bytes;   // one stream's A ring
  const uint32_t b_region = sbase + SmemLayout::a_ring + (uint32_t)NST * ring_bytes;

  // ---------------------------------------------------------------- setup
  if (warp == 0 && lane == 0)
  {
    for (int st = 0; st < NST; ++st)
    {
      for (int s = 0; s < NS; ++s)
      {
        mbar_init(sbase + SmemLayout::full_a + 8 * (st * kMaxStages + s), 1);
        mbar_init(sbase + SmemLayout::empty_a + 8 * (st * kMaxStages + s), 1);
      }
      for (int s = 0; s < R; ++s)
      {
        mbar_init(sbase + SmemLayout::tmem_full + 8 * (st * kMaxSlots + s), 1);
        mbar_init(sbase + SmemLayout::tmem_empty + 8 * (st * kMaxSlots + s), 4); // one arrive per draining warp
      }
    }
    mbar_init(sbase + SmemLayout::w_full, 1);
    fence_mbar_init();
    for (int c = 0; c < p.nchunks; ++c)
    {
      prefetch_tmap(&p.amap[c]);
      prefetch_tmap(&p.wmap[c]);
    }
  }
  if (warp == 1)
  {
    tmem_alloc(sbase + SmemLayout::tmem_ptr, kTmemCols);
    tmem_relinquish();
  }
  if (warp >= kProducerWarps)
  {
    float* bias_s = reinterpret_cast<float*>(sgen + SmemLayout::bias);
    for (int i = threadIdx.x - kProducerWarps * 32; i < p.CoutG; i += kThreads - kProducerWarps * 32)
      bias_s[i] = p.bias[group * p.CoutG + i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_launch_dependents();   // this CTA's resources are the gate for the next grid's CTAs anyway
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(sgen + SmemLayout::tmem_ptr);

  // Both single-issuer roles keep warp-uniform control flow (all 32 lanes walk the loops, one
  // elected lane issues): values stay in uniform registers instead of being broadcast per use.
  if (warp < kProducerWarps)
  {
    const int st = warp >> 1;                     // stream of this TMA / MMA warp
    if (st < NST)
    {
      const int vcta = cta * NST + st, nv = nctas * NST; // streams are independent "virtual CTAs"
      const uint32_t full_a  = sbase + SmemLayout::full_a + 8 * st * kMaxStages;
      const uint32_t empty_a = sbase + SmemLayout::empty_a + 8 * st * kMaxStages;
      const uint32_t a_ring  = sbase + SmemLayout::a_ring + (uint32_t)st * ring_bytes;
      const bool leader = elect_one();
      // -------------------------------------------------------------- TMA producer
      if ((warp & 1) == 0)
      {
        // Resident weights of this CTA's output-channel group: [kw][chunk] blocks of 3*CoutG rows.
        if (st == 0 && leader)
        {
          mbar_arrive_expect_tx(sbase + SmemLayout::w_full, p.w_bytes);
          for (int c = 0; c < p.nchunks; ++c)
            for (int kw = 0; kw < 3; ++kw)
              tma_load_4d(b_region + p.chunk_boff[c] + kw * p.chunk_bblk[c], &p.wmap[c],
                          sbase + SmemLayout::w_full, p.chunk_wc0[c], group * p.CoutG, 0, kw);
        }
        // The launch may overlap the tail of the previous kernel in the stream (programmatic
        // dependent launch): everything above touched only weights; activations written by that
        // kernel are first read here, and this kernel's stores are ordered after these loads.
        pdl_wait();
        // in-frame start stamp: the prerequisite grid has completed, this grid's first activation load follows
        if (p.stamps && st == 0 && leader) atomicMin(&p.stamps[0], globaltimer_ns());
        uint32_t s = 0, ph = 0;            // cursor of the ring (of its src2 part when folded)
        uint32_t su = 0, phu = 0;          // folded: cursor of the src1 part (stages [0, NU))
        const uint32_t NU = (uint32_t)p.nstages_u, NS2 = (uint32_t)NS - NU;
        const int PF = p.prefetch_rows;
        for (int item = vcta; item < nitems; item += nv)
        {
          const Item it = get_item(p, item);
          // Shallow rings (wide layers: the resident weights take most of the shared memory) cannot
          // cover HBM latency with stages in flight, so the rows ahead are pulled into L2 first.
          if (PF > 0 && leader)
          {
            for (int r = it.y0 - 1; r < min(it.y0 - 1 + PF, it.y1 + 2); ++r)
              for (int c = 0; c < p.nchunks; ++c)
              {
                if (p.chunk_up[c]) tma_prefetch_4d(&p.amap[c], p.chunk_c0[c], 0, it.x0 / 2 - 1, r >> 1);
                else               tma_prefetch_3d(&p.amap[c], p.chunk_c0[c], it.x0 - 1, r);
              }
          }
          for (int r = it.y0 - 1; r <= it.y1 + 1; ++r)
          {
            if (PF > 0 && leader && r + PF <= it.y1 + 1)
            {
              const int rp = r + PF;
              for (int c = 0; c < p.nchunks; ++c)
              {
                if (p.chunk_up[c]) { if (!(rp & 1)) tma_prefetch_4d(&p.amap[c], p.chunk_c0[c], 0, it.x0 / 2 - 1, rp >> 1); }
                else               tma_prefetch_3d(&p.amap[c], p.chunk_c0[c], it.x0 - 1, rp);
              }
            }
            for (int c = 0; c < p.nchunks; ++c)
            {
              // row folding: an upsampled source is staged once per LOW-RES row (virtual rows 2Y and 2Y+1 are the same
              // data): at even rows, and at the item's first row whatever its parity
              const bool useU = p.up_fold && p.chunk_up[c];
              if (useU && (r & 1) && r != it.y0 - 1) continue;
              const uint32_t bi = useU ? su : NU + s;
              const uint32_t full = full_a + 8 * bi;
              const uint32_t dst  = a_ring + p.stage_off[bi];
              MBAR_WAIT(empty_a + 8 * bi, (useU ? phu : ph) ^ 1, 1);
              const int cc = p.chunk_cc[c];
              if (leader)
              {
                if (p.chunk_up[c])
                {
                  // 132 virtual pixels starting at x0-2: (dup 2, stride 0) x (66 half-res pixels);
                  // virtual row r is half-res row r>>1 (arithmetic shift keeps r=-1 out of bounds,
                  // which TMA zero-fills).
                  mbar_arrive_expect_tx(full, 132u * cc * 2u);
                  tma_load_4d(dst, &p.amap[c], full, p.chunk_c0[c], 0, it.x0 / 2 - 1, r >> 1);
                }
                else
                {
                  mbar_arrive_expect_tx(full, 130u * cc * 2u);
                  tma_load_3d(dst, &p.amap[c], full, p.chunk_c0[c], it.x0 - 1, r);
                }
              }
              if (useU) { if (++su == NU) { su = 0; phu ^= 1; } }
              else      { if (++s == NS2) { s = 0; ph ^= 1; } }
            }
          }
        }
      }
      // -------------------------------------------------------------- MMA issuer
      else
      {
        // One lane issues every tcgen05.mma of the stream, so this loop must cost only a handful
        // of instructions per MMA and per row: everything below is warp-uniform (uniform
        // registers), descriptors are (per-chunk constant high word | running low word), the
        // per-chunk constants come precomputed from the planner, a row is at most two runs of
        // accumulators contiguous in TMEM, and the per-chunk issue sequence is fully unrolled.
        const uint32_t tfull  = sbase + SmemLayout::tmem_full + 8 * st * kMaxSlots;
        const uint32_t tempty = sbase + SmemLayout::tmem_empty + 8 * st * kMaxSlots;
        const uint32_t CoutG = p.CoutG;
        const uint32_t tbase = tmem_base + (uint32_t)st * (uint32_t)R * CoutG; // this stream's columns
        const uint32_t max_run = min(3u, 256u / CoutG);
        const uint32_t idesc1 = umma_idesc_f16(CoutG);
        const uint32_t sbase16 = (sbase & 0x3FFFFu) >> 4;
        const uint32_t ring16  = sbase16 + ((SmemLayout::a_ring + (uint32_t)st * ring_bytes) >> 4);
        const uint32_t breg16  = sbase16 + ((SmemLayout::a_ring + (uint32_t)NST * ring_bytes) >> 4);
        const int nchunks = p.nchunks;
        MBAR_WAIT(sbase + SmemLayout::w_full, 0, 2);
        tc_fence_after();
        uint32_t stage = 0, sphase = 0;       // A ring position / parity (of the ring's src2 part when folded)
        uint32_t stage_u = 0, sphase_u = 0;   // folded: position / parity of the src1 part (stages [0, NU))
        uint32_t held_u = 0;                  // folded: first src1 stage of the low-res row staged at the last even row
        const uint32_t NU = (uint32_t)p.nstages_u, NS2 = (uint32_t)NS - NU;
        const bool fold = p.up_fold != 0;
        uint32_t a_mod = 0, a_par = 0;        // (first accumulator index of the item) % R, parity of / R
        for (int item = vcta; item < nitems; item += nv)
        {
          const Item it = get_item(p, item);
          uint32_t top_mod = a_mod, top_par = a_par; // accumulator fed by kh=0 of the current row
          for (int r = it.y0 - 1; r <= it.y1 + 1; ++r)
          {
            // Input row r feeds output row y = r - kh + 1 for every kh with y inside the item.
            const int kh_lo = max(0, r + 1 - it.y1);
            const int kh_hi = min(2, r + 1 - it.y0);
            const bool fresh = (kh_lo == 0);
            if (fresh)
            {
              // kh=0 opens a fresh accumulator: its ring slot must have been drained.
              MBAR_WAIT(tempty + 8 * ((R - 1) - top_mod), top_par ^ 1, 3);
              tc_fence_after();
            }
            // Accumulators of kh_lo..kh_hi sit in consecutive ring slots S, S+1, .. (mod R): at most
            // two runs contiguous in TMEM (ring wrap, N <= 256).
            int m = (int)top_mod - kh_lo;
            if (m < 0) m += R;
            const uint32_t S  = (uint32_t)((R - 1) - m);
            const uint32_t n  = (uint32_t)(kh_hi - kh_lo + 1);
            const uint32_t n0 = min(min(n, max_run), (uint32_t)R - S);
            const uint32_t n1 = n - n0;
            uint32_t S1 = S + n0;
            if (S1 >= (uint32_t)R) S1 -= R;
            const uint32_t d0 = tbase + S * CoutG, d1 = tbase + S1 * CoutG;
            const uint32_t brow0 = (uint32_t)kh_lo * CoutG, brow1 = brow0 + n0 * CoutG; // weight rows
            const uint32_t idesc_r0 = umma_idesc_f16(n0 * CoutG);
            const uint32_t idesc_r1 = umma_idesc_f16(n1 * CoutG);

            uint32_t ui = 0;   // folded: ordinal of the src1 chunk inside this row
            for (int c = 0; c < nchunks; ++c)
            {
              const uint32_t nk     = p.chunk_nk[c];                 // 16-channel k-steps: 1, 2 or 4
              const uint32_t row16  = nk * 2;                        // row bytes / 16
              const uint32_t hi     = p.chunk_hi[c];                 // SBO, version, swizzle
              // Which stage, which weights, which accumulators. Folded src1 chunk (ConvKernelParams::up_fold): the
              // low-res row is staged at the even row (or the item's first row) and serves the odd row after it too --
              // there with the O weights (block 3), into the fresh output row r+1 only.
              uint32_t bi = NU + stage;       // index into the stage table / barrier arrays
              bool wait_stage = true, release = true, useU = false;
              uint32_t cn0 = n0, cn1 = n1, crb0 = brow0 * row16, crb1 = brow1 * row16, cd0 = d0, cid0 = idesc_r0;
              if (fold && p.chunk_up[c])
              {
                useU = true;
                const bool odd = (r & 1) != 0, start = (r == it.y0 - 1);
                if (odd && !start)
                {
                  const uint32_t my = ui++;
                  if (kh_lo > 0) continue;                      // output row r+1 is outside the item (stage released at the even row)
                  bi = held_u + my; if (bi >= NU) bi -= NU;
                  wait_stage = false;
                }
                else
                {
                  if (ui == 0) held_u = stage_u;
                  ++ui;
                  bi = stage_u;
                  if (!odd && r + 2 <= it.y1) release = false;    // the odd row that follows still needs it
                }
                if (odd)
                {
                  // kh = 0 only: one row block, the O weights
                  cn0 = 1; cn1 = 0; crb0 = 3u * CoutG * row16; crb1 = 0;
                  cd0 = tbase + ((uint32_t)(R - 1) - top_mod) * CoutG; cid0 = idesc1;
                  if (kh_lo > 0) { if (wait_stage) { /* item start at an odd row always has kh_lo == 0 */ } }
                }
              }
              const uint32_t a_lo   = ring16 + (p.stage_off[bi] >> 4) + p.chunk_a16[c]; // upsampled rows start at x0-2
              const uint32_t b_lo   = breg16 + p.chunk_b16[c];
              const uint32_t bblk16 = p.chunk_bblk16[c];
              const bool first = fresh && c == 0;
              if (wait_stage)
              {
                MBAR_WAIT(full_a + 8 * bi, useU ? sphase_u : sphase, 4);
                tc_fence_after();
              }
              TRACE_BEGIN();
              if (leader)
              {
                if (first)
                {
                  // the first contribution to the fresh accumulator (kh=0 block of run 0) overwrites it
                  umma_f16(cd0, make_desc(hi, a_lo), make_desc(hi, b_lo + crb0), idesc1, 0u);
                  if (cn0 > 1)
                    umma_f16(cd0 + CoutG, make_desc(hi, a_lo), make_desc(hi, b_lo + crb0 + CoutG * row16),
                             umma_idesc_f16((cn0 - 1) * CoutG), 1u);
                  if (cn1)
                    umma_f16(d1, make_desc(hi, a_lo), make_desc(hi, b_lo + crb1), idesc_r1, 1u);
                }
                if (cn1)
                {
                  if (nk == 4)      issue_chunk<4, true>(hi, a_lo, b_lo, bblk16, cd0, cid0, crb0, d1, idesc_r1, crb1, first);
                  else if (nk == 2) issue_chunk<2, true>(hi, a_lo, b_lo, bblk16, cd0, cid0, crb0, d1, idesc_r1, crb1, first);
                  else              issue_chunk<1, true>(hi, a_lo, b_lo, bblk16, cd0, cid0, crb0, d1, idesc_r1, crb1, first);
                }
                else
                {
                  if (nk == 4)      issue_chunk<4, false>(hi, a_lo, b_lo, bblk16, cd0, cid0, crb0, d1, idesc_r1, crb1, first);
                  else if (nk == 2) issue_chunk<2, false>(hi, a_lo, b_lo, bblk16, cd0, cid0, crb0, d1, idesc_r1, crb1, first);
                  else              issue_chunk<1, false>(hi, a_lo, b_lo, bblk16, cd0, cid0, crb0, d1, idesc_r1, crb1, first);
                }
                if (release) umma_commit(empty_a + 8 * bi); // stage reusable once these MMAs retire
              }
              TRACE_END(8);
              if (useU) { if (wait_stage) { if (++stage_u == NU) { stage_u = 0; sphase_u ^= 1; } } }
              else      { if (++stage == NS2) { stage = 0; sphase ^= 1; } }
            }
            // Output row r-1 has now received kh=0,1,2.
            if (r - 1 >= it.y0 && r - 1 <= it.y1)
            {
              int m2 = (int)top_mod - 2;
              if (m2 < 0) m2 += R;
              if (leader)
                umma_commit(tfull + 8 * ((R - 1) - m2));
            }
            if (++top_mod == (uint32_t)R) { top_mod = 0; top_par ^= 1; }
          }
          const uint32_t tot = a_mod + (uint32_t)(it.y1 - it.y0 + 1);
          a_par ^= (tot / R) & 1;
          a_mod = tot % R;
        }
      }
    }
    __syncwarp();
  }
  // ---------------------------------------------------------------- epilogue
  else
  {
    // Specialised per (CoutG / 16, pool): the row loop is the instruction-bound part of the narrow,
    // HBM-bound layers (one warp retires a 32-pixel row segment in ~80 instructions).
    EpiCtx ec;
    ec.sbase = sbase; ec.tmem_base = tmem_base; ec.b_region = b_region; ec.sgen = sgen;
    ec.warp = warp; ec.lane = lane; ec.group = group; ec.cta = cta; ec.nctas = nctas;
    ec.epi_warp0 = kProducerWarps;
    if (p.fo.enabled)
      epilogue_fused_output(p, ec TRACE_ARGS);
    else
    switch ((p.CoutG >> 4) * 2 + (p.post_op == POST_POOL ? 1 : 0))
    {
#define EPI_CASE(NB) case (NB) * 2: epilogue<NB, false, NSTMAX == 2>(p, ec TRACE_ARGS); break; case (NB) * 2 + 1: epilogue<NB, true, NSTMAX == 2>(p, ec TRACE_ARGS); break;
      EPI_CASE(1) EPI_CASE(2)
#undef EPI_CASE
      default:
        if constexpr (NSTMAX == 2)
        {
          switch ((p.CoutG >> 4) * 2 + (p.post_op == POST_POOL ? 1 : 0))
          {
#define EPI_CASE(NB) case (NB) * 2: epilogue<NB, false, true>(p, ec TRACE_ARGS); break; case (NB) * 2 + 1: epilogue<NB, true, true>(p, ec TRACE_ARGS); break;
            EPI_CASE(3) EPI_CASE(4) EPI_CASE(5) EPI_CASE(6) EPI_CASE(7) EPI_CASE(8)
#undef EPI_CASE
            default: break;
          }
        }
        break;
    }
  }

  // ---------------------------------------------------------------- teardown
  TRACE_FLUSH();
  tc_fence_before();
  __syncthreads();
  if (p.stamps && threadIdx.x == 0) atomicMax(&p.stamps[1], globaltimer_ns());
  if (warp == 1)
  {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

__global__ void __launch_bounds__(kConvThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ ConvKernelParams p)
{
  conv3x3_tc_body<2>(p);
}

__global__ void __launch_bounds__(kConvThreads4, 1)
conv3x3_tc_kernel_s4(const __grid_constant__ ConvKernelParams p)
{
  conv3x3_tc_body<4>(p);
}

cudaError_t conv3x3_tc_launch(const ConvKernelParams& p, int grid, size_t smem_bytes, cudaStream_t stream)
{
  // the opt-in shared-memory attribute is per device
  static bool attr_set[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 64 && !attr_set[dev])
  {
    e = cudaFuncSetAttribute(conv3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(conv3x3_tc_kernel_s4, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  const bool four = p.nstreams == 4;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(four ? kConvThreads4 : kConvThreads);
  cfg.dynamicSmemBytes = smem_bytes; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return four ? cudaLaunchKernelEx(&cfg, conv3x3_tc_kernel_s4, p) : cudaLaunchKernelEx(&cfg, conv3x3_tc_kernel, p);
}

} // namespace oidnb200
