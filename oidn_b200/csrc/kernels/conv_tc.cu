// 3x3 convolution (pad 1, stride 1, cross-correlation) as an implicit GEMM on tcgen05 / TMEM.
//
// Replaces the reference's CutlassConv (devices/cuda/cutlass_conv.h:150-278) plus its separate
// pool / concat passes (devices/gpu/gpu_pool.h, core/concat_conv_hwc.cpp:27-31).
// Semantics follow core/conv.cpp:8-60 and the CPU kernel devices/cpu/cpu_conv.ispc:34-127:
//   dst[o,y,x] = act(bias[o] + sum_{i,kh,kw} W[o,i,kh,kw] * src[i,y+kh-1,x+kw-1]), zero padding,
//   fp32 accumulation; optional 2x2 max-pool of the result; src may be the channel concat of two
//   tensors (read in place, never materialised) and src1 may be a half-resolution tensor that is
//   nearest-upsampled on the fly by the loader.
//
// Mapping (one persistent CTA per SM, warp specialised: warp0 = TMA producer, warp1 = MMA issuer,
// warps 2-5 = epilogue):
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
//   * Epilogue: tcgen05.ld -> +bias -> ReLU -> (2x2 max-pool via a second accumulator + shuffle)
//     -> fp16 -> global.
#include "conv_common.h"
#include "ptx.cuh"
#include <cuda_fp16.h>

namespace oidnb200 {

using namespace ptx;

namespace {

struct SmemLayout
{
  // byte offsets from the 1024-aligned base
  static constexpr uint32_t full_a     = 0;                       // kMaxStages x 8
  static constexpr uint32_t empty_a    = full_a + 8 * kMaxStages;
  static constexpr uint32_t w_full     = empty_a + 8 * kMaxStages;
  static constexpr uint32_t tmem_full  = w_full + 8;              // kMaxSlots x 8
  static constexpr uint32_t tmem_empty = tmem_full + 8 * kMaxSlots;
  static constexpr uint32_t tmem_ptr   = tmem_empty + 8 * kMaxSlots;
  static constexpr uint32_t bias       = 1024;                    // 128 floats
  static constexpr uint32_t a_ring     = kSmemHeader;
};
static_assert(SmemLayout::tmem_ptr + 4 <= 1024, "barrier block overflows");

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

__device__ __forceinline__ uint32_t pack_half2(float a, float b)
{
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

} // namespace

__global__ void __launch_bounds__(192, 1)
conv3x3_tc_kernel(const __grid_constant__ ConvKernelParams p)
{
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int group   = blockIdx.x % p.ngroups;
  const int cta     = blockIdx.x / p.ngroups;
  const int nctas   = gridDim.x / p.ngroups;
  const int nitems  = p.nstrips * p.nrowchunks;
  const int NS      = p.nstages;
  const int R       = p.R;
  const uint32_t stage_bytes = (p.shift_mode == 2) ? 3u * 16384u : (uint32_t)kStageBytes;
  const uint32_t b_region = sbase + SmemLayout::a_ring + NS * stage_bytes;

  // ---------------------------------------------------------------- setup
  if (warp == 0 && lane == 0)
  {
    for (int s = 0; s < NS; ++s)
    {
      mbar_init(sbase + SmemLayout::full_a + 8 * s, 1);
      mbar_init(sbase + SmemLayout::empty_a + 8 * s, 1);
    }
    mbar_init(sbase + SmemLayout::w_full, 1);
    for (int s = 0; s < R; ++s)
    {
      mbar_init(sbase + SmemLayout::tmem_full + 8 * s, 1);
      mbar_init(sbase + SmemLayout::tmem_empty + 8 * s, 4); // one arrive per epilogue warp
    }
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
  if (warp >= 2)
  {
    float* bias_s = reinterpret_cast<float*>(sgen + SmemLayout::bias);
    for (int i = threadIdx.x - 64; i < p.CoutG; i += 128)
      bias_s[i] = p.bias[group * p.CoutG + i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(sgen + SmemLayout::tmem_ptr);

  // ---------------------------------------------------------------- TMA producer
  if (warp == 0)
  {
    if (lane == 0)
    {
      // Resident weights of this CTA's output-channel group: [kw][chunk] blocks of 3*CoutG rows.
      mbar_arrive_expect_tx(sbase + SmemLayout::w_full, p.w_bytes);
      for (int c = 0; c < p.nchunks; ++c)
        for (int kw = 0; kw < 3; ++kw)
          tma_load_4d(b_region + p.chunk_boff[c] + kw * p.chunk_bblk[c], &p.wmap[c],
                      sbase + SmemLayout::w_full, p.chunk_wc0[c], group * p.CoutG, 0, kw);

      uint32_t k = 0;
      for (int item = cta; item < nitems; item += nctas)
      {
        const Item it = get_item(p, item);
        for (int r = it.y0 - 1; r <= it.y1 + 1; ++r)
        {
          for (int c = 0; c < p.nchunks; ++c, ++k)
          {
            const uint32_t s  = k % NS;
            const uint32_t ph = (k / NS) & 1;
            const uint32_t full = sbase + SmemLayout::full_a + 8 * s;
            const uint32_t dst  = sbase + SmemLayout::a_ring + s * stage_bytes;
            mbar_wait(sbase + SmemLayout::empty_a + 8 * s, ph ^ 1, 1);
            const int cc = p.chunk_cc[c];
            if (p.chunk_up[c])
            {
              // 132 virtual pixels starting at x0-2: (dup 2, stride 0) x (66 half-res pixels);
              // virtual row r is half-res row r>>1 (arithmetic shift keeps r=-1 out of bounds,
              // which TMA zero-fills).
              mbar_arrive_expect_tx(full, 132u * cc * 2u);
              tma_load_4d(dst, &p.amap[c], full, p.chunk_c0[c], 0, it.x0 / 2 - 1, r >> 1);
            }
            else if (p.shift_mode == 2)
            {
              mbar_arrive_expect_tx(full, 3u * 128u * cc * 2u);
              for (int kw = 0; kw < 3; ++kw)
                tma_load_3d(dst + kw * 16384u, &p.amap[c], full, p.chunk_c0[c], it.x0 - 1 + kw, r);
            }
            else
            {
              mbar_arrive_expect_tx(full, 130u * cc * 2u);
              tma_load_3d(dst, &p.amap[c], full, p.chunk_c0[c], it.x0 - 1, r);
            }
          }
        }
      }
    }
    __syncwarp();
  }
  // ---------------------------------------------------------------- MMA issuer
  else if (warp == 1)
  {
    if (lane == 0)
    {
      mbar_wait(sbase + SmemLayout::w_full, 0, 2);
      tc_fence_after();
      const int CoutG = p.CoutG;
      const int max_run = min(3, 256 / CoutG);
      uint32_t k = 0;
      uint32_t accbase = 0;
      for (int item = cta; item < nitems; item += nctas)
      {
        const Item it = get_item(p, item);
        for (int r = it.y0 - 1; r <= it.y1 + 1; ++r)
        {
          // Input row r feeds output row y = r - kh + 1 for every kh with y inside the item.
          const int kh_lo = max(0, r + 1 - it.y1);
          const int kh_hi = min(2, r + 1 - it.y0);
          const uint32_t a_top = accbase + (uint32_t)(r + 1 - it.y0); // accumulator index of kh=0
          if (kh_lo == 0)
          {
            // kh=0 opens a fresh accumulator: its ring slot must have been drained.
            const uint32_t slot = (R - 1) - (a_top % R);
            mbar_wait(sbase + SmemLayout::tmem_empty + 8 * slot, ((a_top / R) & 1) ^ 1, 3);
            tc_fence_after();
          }
          // Split kh_lo..kh_hi into runs that are contiguous in TMEM (ring wrap) and N <= 256.
          int run_kh[3], run_n[3], nruns = 0;
          for (int kh = kh_lo; kh <= kh_hi; ++kh)
          {
            const uint32_t slot = (R - 1) - ((a_top - kh) % R);
            if (nruns > 0 && slot != 0 && run_n[nruns - 1] < max_run)
              run_n[nruns - 1]++;
            else
            {
              run_kh[nruns] = kh;
              run_n[nruns]  = 1;
              nruns++;
            }
          }
          for (int c = 0; c < p.nchunks; ++c, ++k)
          {
            const uint32_t s  = k % NS;
            const uint32_t ph = (k / NS) & 1;
            mbar_wait(sbase + SmemLayout::full_a + 8 * s, ph, 4);
            tc_fence_after();
            const uint32_t cc = p.chunk_cc[c];
            const uint32_t row_bytes = cc * 2;
            const uint32_t a_stage = sbase + SmemLayout::a_ring + s * stage_bytes;
            const uint32_t px_off  = p.chunk_up[c] ? 1u : 0u; // upsampled rows start at x0-2
            for (int kw = 0; kw < 3; ++kw)
            {
              const uint32_t a_tap = (p.shift_mode == 2 && !p.chunk_up[c])
                                       ? a_stage + kw * 16384u
                                       : a_stage + (kw + px_off) * row_bytes;
              const uint32_t b_tap = b_region + p.chunk_boff[c] + kw * p.chunk_bblk[c];
              for (uint32_t j = 0; j < cc / 16; ++j)
              {
                const uint32_t a_addr = a_tap + j * 32;
                const uint32_t a_bo   = (p.shift_mode == 1) ? ((a_addr >> 7) & 7u) : 0u;
                const uint64_t adesc  = umma_desc(a_addr, row_bytes, a_bo);
                const bool first = (c == 0 && kw == 0 && j == 0);
                for (int q = 0; q < nruns; ++q)
                {
                  int kh = run_kh[q], n = run_n[q];
                  const uint32_t slot = (R - 1) - ((a_top - kh) % R);
                  uint32_t d_addr = tmem_base + slot * CoutG;
                  uint32_t b_addr = b_tap + kh * CoutG * row_bytes + j * 32;
                  if (first && kh == 0)
                  {
                    // first contribution to the fresh accumulator overwrites it
                    umma_f16(d_addr, adesc, umma_desc(b_addr, row_bytes, 0),
                             umma_idesc_f16(CoutG), 0u);
                    kh++; n--;
                    d_addr += CoutG;
                    b_addr += CoutG * row_bytes;
                  }
                  if (n > 0)
                    umma_f16(d_addr, adesc, umma_desc(b_addr, row_bytes, 0),
                             umma_idesc_f16(n * CoutG), 1u);
                }
              }
            }
            umma_commit(sbase + SmemLayout::empty_a + 8 * s); // stage reusable once these MMAs retire
          }
          // Output row r-1 has now received kh=0,1,2.
          if (r - 1 >= it.y0 && r - 1 <= it.y1)
          {
            const uint32_t a_done = accbase + (uint32_t)(r - 1 - it.y0);
            umma_commit(sbase + SmemLayout::tmem_full + 8 * ((R - 1) - (a_done % R)));
          }
        }
        accbase += (uint32_t)(it.y1 - it.y0 + 1);
      }
    }
    __syncwarp();
  }
  // ---------------------------------------------------------------- epilogue
  else
  {
    const int q    = warp & 3;                  // TMEM lane quarter this warp may access
    const int lpix = q * 32 + lane;             // pixel within the strip
    const float* bias_s = reinterpret_cast<const float*>(sgen + SmemLayout::bias);
    const int CoutG = p.CoutG;
    const int cbase = group * CoutG;
    __half* dst = reinterpret_cast<__half*>(p.dst);
    const bool pool = (p.post_op == POST_POOL);
    const int ystep = pool ? 2 : 1;
    uint32_t accbase = 0;
    for (int item = cta; item < nitems; item += nctas)
    {
      const Item it = get_item(p, item);
      const int x = it.x0 + lpix;
      for (int y = it.y0; y <= it.y1; y += ystep)
      {
        const uint32_t a0 = accbase + (uint32_t)(y - it.y0);
        const uint32_t slot0 = (R - 1) - (a0 % R);
        const uint32_t slot1 = (R - 1) - ((a0 + 1) % R);
        mbar_wait(sbase + SmemLayout::tmem_full + 8 * slot0, (a0 / R) & 1, 5);
        if (pool)
          mbar_wait(sbase + SmemLayout::tmem_full + 8 * slot1, ((a0 + 1) / R) & 1, 6);
        tc_fence_after();
        const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + slot0 * CoutG;
        const uint32_t t1 = tmem_base + ((uint32_t)(q * 32) << 16) + slot1 * CoutG;
        for (int j = 0; j < CoutG; j += 16)
        {
          uint32_t v[16];
          tmem_ld16(t0 + j, v);
          float f[16];
          if (pool)
          {
            uint32_t w[16];
            tmem_ld16(t1 + j, w);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i)
            {
              float m = fmaxf(__uint_as_float(v[i]), __uint_as_float(w[i]));
              m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
              f[i] = m;
            }
          }
          else
          {
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i)
              f[i] = __uint_as_float(v[i]);
          }
          uint32_t h[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
          {
            float a = f[2 * i] + bias_s[j + 2 * i];
            float b = f[2 * i + 1] + bias_s[j + 2 * i + 1];
            if (p.relu)
            {
              a = fmaxf(a, 0.f);
              b = fmaxf(b, 0.f);
            }
            h[i] = pack_half2(a, b);
          }
          const int co = cbase + j;
          if (x < p.W && co < p.CoutPad)
          {
            const uint4 lo = make_uint4(h[0], h[1], h[2], h[3]);
            const uint4 hi = make_uint4(h[4], h[5], h[6], h[7]);
            if (pool)
            {
              if ((lane & 1) == 0)
              {
                const size_t o = ((size_t)(y >> 1) * (p.W >> 1) + (x >> 1)) * p.dstC + co;
                *reinterpret_cast<uint4*>(dst + o)     = lo;
                *reinterpret_cast<uint4*>(dst + o + 8) = hi;
              }
            }
            else if (p.post_op == POST_UPSAMPLE)
            {
#pragma unroll
              for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx)
                {
                  const size_t o = ((size_t)(2 * y + dy) * (2 * p.W) + (2 * x + dx)) * p.dstC + co;
                  *reinterpret_cast<uint4*>(dst + o)     = lo;
                  *reinterpret_cast<uint4*>(dst + o + 8) = hi;
                }
            }
            else
            {
              const size_t o = ((size_t)y * p.W + x) * p.dstC + co;
              *reinterpret_cast<uint4*>(dst + o)     = lo;
              *reinterpret_cast<uint4*>(dst + o + 8) = hi;
            }
          }
        }
        // Release the accumulator slot(s) back to the MMA issuer.
        tc_fence_before();
        __syncwarp();
        if (lane == 0)
        {
          mbar_arrive(sbase + SmemLayout::tmem_empty + 8 * slot0);
          if (pool)
            mbar_arrive(sbase + SmemLayout::tmem_empty + 8 * slot1);
        }
      }
      accbase += (uint32_t)(it.y1 - it.y0 + 1);
    }
  }

  // ---------------------------------------------------------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 1)
  {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

cudaError_t conv3x3_tc_launch(const ConvKernelParams& p, int grid, size_t smem_bytes, cudaStream_t stream)
{
  static bool attr_set = false;
  if (!attr_set)
  {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kSmemBudget);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  conv3x3_tc_kernel<<<grid, 192, smem_bytes, stream>>>(p);
  return cudaGetLastError();
}

} // namespace oidnb200
