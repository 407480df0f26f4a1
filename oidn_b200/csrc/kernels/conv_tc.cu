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
  const uint32_t stage_bytes = (uint32_t)kStageBytes;
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
  // Both single-issuer roles keep warp-uniform control flow (all 32 lanes walk the loops, one
  // elected lane issues): values stay in uniform registers instead of being broadcast per use.
  if (warp == 0)
  {
    const bool leader = elect_one();
    // Resident weights of this CTA's output-channel group: [kw][chunk] blocks of 3*CoutG rows.
    if (leader)
    {
      mbar_arrive_expect_tx(sbase + SmemLayout::w_full, p.w_bytes);
      for (int c = 0; c < p.nchunks; ++c)
        for (int kw = 0; kw < 3; ++kw)
          tma_load_4d(b_region + p.chunk_boff[c] + kw * p.chunk_bblk[c], &p.wmap[c],
                      sbase + SmemLayout::w_full, p.chunk_wc0[c], group * p.CoutG, 0, kw);
    }
    uint32_t s = 0, ph = 0;
    for (int item = cta; item < nitems; item += nctas)
    {
      const Item it = get_item(p, item);
      for (int r = it.y0 - 1; r <= it.y1 + 1; ++r)
      {
        for (int c = 0; c < p.nchunks; ++c)
        {
          const uint32_t full = sbase + SmemLayout::full_a + 8 * s;
          const uint32_t dst  = sbase + SmemLayout::a_ring + s * stage_bytes;
          mbar_wait(sbase + SmemLayout::empty_a + 8 * s, ph ^ 1, 1);
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
          if (++s == (uint32_t)NS) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  }
  // ---------------------------------------------------------------- MMA issuer
  else if (warp == 1)
  {
    // One lane issues every tcgen05.mma of the CTA, so this loop must cost only a handful of
    // instructions per MMA: descriptors are (constant high word | running low word), ring/slot
    // arithmetic is hoisted to once per input row, and there are at most two runs per row.
    const bool leader = elect_one();
    mbar_wait(sbase + SmemLayout::w_full, 0, 2);
    tc_fence_after();
    const uint32_t CoutG = p.CoutG;
    const uint32_t max_run = min(3u, 256u / CoutG);
    const uint32_t idesc1 = umma_idesc_f16(CoutG);
    uint32_t stage = 0, sphase = 0;       // A ring position / parity
    uint32_t a_mod = 0, a_par = 0;        // (first accumulator index of the item) % R, parity of / R
    for (int item = cta; item < nitems; item += nctas)
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
          mbar_wait(sbase + SmemLayout::tmem_empty + 8 * ((R - 1) - top_mod), top_par ^ 1, 3);
          tc_fence_after();
        }
        // Split kh_lo..kh_hi into (at most two) runs contiguous in TMEM (ring wrap) with N <= 256.
        uint32_t d0 = 0, n0 = 0, brow0 = 0, d1 = 0, n1 = 0, brow1 = 0;
        for (int kh = kh_lo; kh <= kh_hi; ++kh)
        {
          int m = (int)top_mod - kh;
          if (m < 0) m += R;
          const uint32_t slot = (R - 1) - m;
          if (n0 == 0)
          {
            d0 = tmem_base + slot * CoutG; brow0 = kh * CoutG; n0 = 1;
          }
          else if (n1 == 0 && slot != 0 && n0 < max_run)
            n0++;
          else if (n1 == 0)
          {
            d1 = tmem_base + slot * CoutG; brow1 = kh * CoutG; n1 = 1;
          }
          else
            n1++;
        }
        const uint32_t idesc_r0 = umma_idesc_f16(n0 * CoutG);
        const uint32_t idesc_r1 = umma_idesc_f16(n1 * CoutG);

        for (int c = 0; c < p.nchunks; ++c)
        {
          mbar_wait(sbase + SmemLayout::full_a + 8 * stage, sphase, 4);
          tc_fence_after();
          const uint32_t cc     = p.chunk_cc[c];
          const uint32_t row16  = cc >> 3;                      // row bytes / 16
          const uint32_t hi     = (uint32_t)(umma_desc(0, cc * 2, 0) >> 32); // SBO, version, swizzle
          const uint32_t a_base = sbase + SmemLayout::a_ring + stage * stage_bytes
                                  + (p.chunk_up[c] ? cc * 2 : 0); // upsampled rows start at x0-2
          const uint32_t a_lo0  = (a_base & 0x3FFFFu) >> 4;
          const uint32_t b_lo0  = ((b_region + p.chunk_boff[c]) & 0x3FFFFu) >> 4;
          const uint32_t bblk16 = p.chunk_bblk[c] >> 4;
          const uint32_t rb0 = brow0 * row16, rb1 = brow1 * row16;
          const uint32_t nk = cc >> 4;
          uint32_t j0 = 0;
          if (fresh && c == 0)
          {
            // the first contribution to the fresh accumulator (kh=0 block of run 0) overwrites it
            if (leader)
            {
              umma_f16(d0, make_desc(hi, a_lo0), make_desc(hi, b_lo0 + rb0), idesc1, 0u);
              if (n0 > 1)
                umma_f16(d0 + CoutG, make_desc(hi, a_lo0), make_desc(hi, b_lo0 + rb0 + CoutG * row16),
                         umma_idesc_f16((n0 - 1) * CoutG), 1u);
              if (n1)
                umma_f16(d1, make_desc(hi, a_lo0), make_desc(hi, b_lo0 + rb1), idesc_r1, 1u);
            }
            j0 = 1;
          }
          for (uint32_t kw = 0; kw < 3; ++kw)
          {
            const uint32_t a_kw = a_lo0 + kw * row16;
            const uint32_t b_kw = b_lo0 + kw * bblk16;
            for (uint32_t j = (kw == 0 ? j0 : 0u); j < nk; ++j)
            {
              if (leader)
              {
                umma_f16(d0, make_desc(hi, a_kw + 2 * j), make_desc(hi, b_kw + rb0 + 2 * j), idesc_r0, 1u);
                if (n1)
                  umma_f16(d1, make_desc(hi, a_kw + 2 * j), make_desc(hi, b_kw + rb1 + 2 * j), idesc_r1, 1u);
              }
            }
          }
          if (leader)
            umma_commit(sbase + SmemLayout::empty_a + 8 * stage); // stage reusable once these MMAs retire
          if (++stage == (uint32_t)NS) { stage = 0; sphase ^= 1; }
        }
        // Output row r-1 has now received kh=0,1,2.
        if (r - 1 >= it.y0 && r - 1 <= it.y1)
        {
          int m = (int)top_mod - 2;
          if (m < 0) m += R;
          if (leader)
            umma_commit(sbase + SmemLayout::tmem_full + 8 * ((R - 1) - m));
        }
        if (++top_mod == (uint32_t)R) { top_mod = 0; top_par ^= 1; }
      }
      const uint32_t tot = a_mod + (uint32_t)(it.y1 - it.y0 + 1);
      a_par ^= (tot / R) & 1;
      a_mod = tot % R;
    }
    __syncwarp();
  }
  // ---------------------------------------------------------------- epilogue
  else
  {
    // 4 warps = 128 threads = the 128 TMEM lanes (pixels) of an accumulator. Per output row:
    // TMEM -> registers -> +bias, ReLU, (pool) -> fp16 -> swizzled smem staging -> TMA store.
    const int q    = warp & 3;                  // TMEM lane quarter this warp may access
    const int lpix = q * 32 + lane;             // pixel within the strip
    const bool issuer = (threadIdx.x == 64);    // first epilogue thread issues the TMA stores
    const float* bias_s = reinterpret_cast<const float*>(sgen + SmemLayout::bias);
    const int CoutG = p.CoutG;
    const bool pool = (p.post_op == POST_POOL);
    const int ystep = pool ? 2 : 1;
    const uint32_t out_region = b_region + p.b_bytes;
    const int nbuf = p.out_nbuf;
    const int spix = pool ? (lpix >> 1) : lpix; // staging row of this thread's pixel
    const bool writer = !pool || ((lane & 1) == 0);
    uint32_t a_mod = 0, a_par = 0;
    uint32_t buf = 0;
    for (int item = cta; item < nitems; item += nctas)
    {
      const Item it = get_item(p, item);
      uint32_t y_mod = a_mod, y_par = a_par;
      for (int y = it.y0; y <= it.y1; y += ystep)
      {
        const uint32_t slot0 = (R - 1) - y_mod;
        const uint32_t par0 = y_par;
        if (++y_mod == (uint32_t)R) { y_mod = 0; y_par ^= 1; }
        const uint32_t slot1 = (R - 1) - y_mod;
        const uint32_t par1 = y_par;
        if (pool) { if (++y_mod == (uint32_t)R) { y_mod = 0; y_par ^= 1; } }

        // staging buffer `buf` must have been read out by its previous TMA store
        if (issuer)
        {
          if (nbuf == 2) bulk_wait_read<1>(); else bulk_wait_read<0>();
        }
        named_bar_sync(1, 128);

        mbar_wait(sbase + SmemLayout::tmem_full + 8 * slot0, par0, 5);
        if (pool)
          mbar_wait(sbase + SmemLayout::tmem_full + 8 * slot1, par1, 6);
        tc_fence_after();
        const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + slot0 * CoutG;
        const uint32_t t1 = tmem_base + ((uint32_t)(q * 32) << 16) + slot1 * CoutG;
        const uint32_t stage_out = out_region + buf * p.out_buf_bytes;
        for (int oc = 0; oc < p.nout; ++oc)
        {
          const int c0 = p.out_c0[oc], ccw = p.out_cc[oc];
          const uint32_t rowb = ccw * 2;
          const uint32_t mask = (rowb >> 4) - 1;              // 16-B chunks per row - 1 (1,3,7)
          const uint32_t piece = stage_out + p.out_off[oc];
          for (int j = 0; j < ccw; j += 16)
          {
            uint32_t v[16];
            float f[16];
            tmem_ld16(t0 + c0 + j, v);
            if (pool)
            {
              uint32_t w[16];
              tmem_ld16(t1 + c0 + j, w);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i)
              {
                float m = fmaxf(__uint_as_float(v[i]), __uint_as_float(w[i]));
                f[i] = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
              }
            }
            else
            {
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
            }
            const float4* b4 = reinterpret_cast<const float4*>(bias_s + c0 + j);
            uint32_t h[8];
#pragma unroll
            for (int i = 0; i < 4; ++i)
            {
              const float4 bb = b4[i];
              float a0 = f[4 * i] + bb.x, a1 = f[4 * i + 1] + bb.y, a2 = f[4 * i + 2] + bb.z, a3 = f[4 * i + 3] + bb.w;
              if (p.relu)
              {
                a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); a2 = fmaxf(a2, 0.f); a3 = fmaxf(a3, 0.f);
              }
              h[2 * i] = pack_half2(a0, a1);
              h[2 * i + 1] = pack_half2(a2, a3);
            }
            if (writer)
            {
              // swizzled staging row: 16-B chunk index XOR (128-B line index mod chunks-per-atom-row)
              const uint32_t off0 = (uint32_t)spix * rowb + (uint32_t)j * 2;
              const uint32_t sw0 = off0 ^ (((off0 >> 7) & mask) << 4);
              const uint32_t off1 = off0 + 16;
              const uint32_t sw1 = off1 ^ (((off1 >> 7) & mask) << 4);
              st_shared_v4(piece + sw0, h[0], h[1], h[2], h[3]);
              st_shared_v4(piece + sw1, h[4], h[5], h[6], h[7]);
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
        // Make the generic-proxy smem writes visible to the TMA engine, then store the row.
        fence_proxy_async();
        named_bar_sync(2, 128);
        if (issuer)
        {
          const int yo = pool ? (y >> 1) : y;
          const int xo = pool ? (it.x0 >> 1) : it.x0;
          for (int oc = 0; oc < p.nout; ++oc)
            tma_store_3d(&p.omap[oc], stage_out + p.out_off[oc], group * CoutG + p.out_c0[oc], xo, yo);
          bulk_commit();
        }
        if (++buf == (uint32_t)nbuf) buf = 0;
      }
      const uint32_t tot = a_mod + (uint32_t)(it.y1 - it.y0 + 1);
      a_par ^= (tot / R) & 1;
      a_mod = tot % R;
    }
    if (issuer) bulk_wait_read<0>();
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
