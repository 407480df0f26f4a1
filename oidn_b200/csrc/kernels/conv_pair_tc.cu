// Two chained 3x3 convolutions in ONE kernel: B(A(x)), the tensor between them never leaves the SM.
//
// The UNet's full-resolution pairs -- enc_conv0 -> enc_conv1 (+2x2 max-pool) and dec_conv1b -> dec_conv0
// (+output process), and their counterparts in the small / large nets (core/unet_filter.cpp:468-531) -- are
// HBM-bound as separate launches: the 32/64-channel full-resolution tensor between them costs a write and a
// re-read of 64-128 B per pixel each. Here conv A's epilogue writes its rows (bias, ReLU, fp16 rounding -- the
// bits the stored tensor would have had, so the result is bit-identical to the two-launch path) straight into a
// shared-memory ring in the swizzled K-major layout conv B's tcgen05.mma reads as its A operand.
//
// Mapping (persistent CTA, up to two independent row streams, as conv_tc.cu):
//   * a work item is a strip of 126 output pixels of B x RC rows. A's M tile are the 128 pixels x0-1 .. x0+126
//     of a row (B's horizontal taps need one pixel either side), fed by the usual 130-pixel TMA box at x0-2;
//     vertically A produces the RC+2 rows y0-1 .. y1+1 from the RC+4 input rows y0-2 .. y1+2. Recompute: 2/128
//     columns + 2/RC rows of conv A.
//   * B's zero padding applies to A's OUTPUT tensor: A's epilogue writes zeros for pixels outside the image.
//   * per stream: TMA warp -> input ring -> A's MMA warp -> A-epilogue warpgroup (TMEM -> mid ring) -> B's MMA
//     warp -> B-epilogue warpgroup (TMEM -> global: plain / pooled tensor with 32-byte stores, or the output image
//     through the fused output process). Both convs stack the three vertical taps along N and take the horizontal
//     taps as shifted descriptor views. A and B have their OWN issuing threads: four issuers per CTA.
// Warp roles: 0/2 = TMA stream 0/1, 1/3 = A's MMA stream 0/1 (warp 1 allocates TMEM), 4-7 / 8-11 = A epilogue of
// stream 0/1, 12-15 / 16-19 = B epilogue of stream 0/1, 20/21 = B's MMA stream 0/1.
#include "conv_common.h"
#include "ptx.cuh"
#include "transfer.cuh"
#include "conv_util.cuh"
#include <cuda_fp16.h>

namespace oidnb200 {

using namespace ptx;

// Optional wait-time tracing (-DOIDN_B200_TRACE, tools/probe_pair trace build): per warp, cycles blocked per tag.
#ifdef OIDN_B200_TRACE
#define PTR_DECL long long tr[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; const long long tr_start = clock64();
#define PWAIT(bar, par, tag) do { const long long t0_ = clock64(); mbar_wait(bar, par, tag); tr[tag] += clock64() - t0_; } while (0)
#define PTIME(tag, stmt) do { const long long t0_ = clock64(); stmt; tr[tag] += clock64() - t0_; } while (0)
#define PTR_FLUSH() do { if (p.trace && lane == 0 && warp < 22) { tr[0] = clock64() - tr_start; \
  for (int i_ = 0; i_ < 16; ++i_) atomicAdd(&p.trace[warp * 16 + i_], (unsigned long long)tr[i_]); } } while (0)
#define PTR_ARG , tr
#define PTR_PARAM , long long* tr
#else
#define PTR_DECL
#define PWAIT(bar, par, tag) mbar_wait(bar, par, tag)
#define PTIME(tag, stmt) stmt
#define PTR_FLUSH()
#define PTR_ARG
#define PTR_PARAM
#endif

namespace {

struct PairSmem
{
  // byte offsets from the 1024-aligned base; per-stream arrays are [stream][index]
  static constexpr uint32_t full_a    = 0;                                      // 2 x kPairMaxStages x 8
  static constexpr uint32_t empty_a   = full_a + 2 * 8 * kPairMaxStages;
  static constexpr uint32_t mid_full  = empty_a + 2 * 8 * kPairMaxStages;       // 2 x kPairMaxMid x 8
  static constexpr uint32_t mid_empty = mid_full + 2 * 8 * kPairMaxMid;
  static constexpr uint32_t acca_full = mid_empty + 2 * 8 * kPairMaxMid;        // 2 x kMaxSlots x 8
  static constexpr uint32_t acca_empty = acca_full + 2 * 8 * kMaxSlots;
  static constexpr uint32_t accb_full = acca_empty + 2 * 8 * kMaxSlots;
  static constexpr uint32_t accb_empty = accb_full + 2 * 8 * kMaxSlots;
  static constexpr uint32_t w_full    = accb_empty + 2 * 8 * kMaxSlots;
  static constexpr uint32_t tmem_ptr  = w_full + 8;
  static constexpr uint32_t bias_a    = 3072;                                   // 64 floats
  static constexpr uint32_t bias_b    = 3072 + 256;                             // 64 floats
  static constexpr uint32_t xch       = 3072 + 512;                             // tap-packed B: [stream][row parity][warp][8] floats
};
static_assert(PairSmem::tmem_ptr + 4 <= PairSmem::bias_a && PairSmem::xch + 2 * 2 * 4 * 8 * 4 <= kSmemHeader, "barrier block overflows");

struct PItem { int x0, y0, y1; };

__device__ __forceinline__ PItem pair_item(const PairKernelParams& p, int item)
{
  PItem it;
  const int strip = item % p.nstrips;
  const int rc = item / p.nstrips;
  it.x0 = strip * kPairStrip;
  it.y0 = rc * p.RC;
  it.y1 = min(p.H, it.y0 + p.RC) - 1;
  return it;
}

// One conv's constants for the MMA issuer (descriptor arithmetic in 16-byte units, as conv_tc.cu)
struct Side
{
  uint32_t Cout, R, tbase, max_run, idesc1, hi, nk, row16, b_lo, bblk16, full_bar, empty_bar, ntaps;
};

template <int NK, bool TWO>
__device__ __forceinline__ void issue_taps(uint32_t hi, uint32_t a_lo, uint32_t b_lo, uint32_t bblk16, uint32_t d0, uint32_t idesc0,
                                           uint32_t rb0, uint32_t d1, uint32_t idesc1, uint32_t rb1, bool skip_first, uint32_t ntaps)
{
  constexpr uint32_t row16 = NK * 2;
#pragma unroll
  for (int kw = 0; kw < 3; ++kw)
  {
    if (kw >= (int)ntaps) break;   // tap-packed B: the horizontal taps are columns of the one unshifted view
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

// All MMAs of one staged source row r (stage start a_lo) into the accumulators of the output rows [o0, o1] it
// touches: out row y = r - kh + 1. (top_mod, top_par) = ring position of the accumulator kh = 0 feeds. Commits
// `release_bar` (the stage may be refilled once these MMAs retire) and, when r completes output row r-1, that
// row's "accumulator full" barrier. Warp-uniform; only `leader` issues.
__device__ __forceinline__ void issue_row(const Side& c, bool leader, uint32_t a_lo, int r, int o0, int o1, uint32_t release_bar,
                                          uint32_t& top_mod, uint32_t& top_par, int tag PTR_PARAM)
{
  const int R = (int)c.R;
  const int kh_lo = max(0, r + 1 - o1);
  const int kh_hi = min(2, r + 1 - o0);
  const bool fresh = (kh_lo == 0);
  if (fresh)
  {
    PWAIT(c.empty_bar + 8 * ((R - 1) - top_mod), top_par ^ 1, tag);
    tc_fence_after();
  }
#ifdef OIDN_B200_TRACE
  const long long ti0_ = clock64();
#endif
  int m = (int)top_mod - kh_lo;
  if (m < 0) m += R;
  const uint32_t S  = (uint32_t)((R - 1) - m);
  const uint32_t n  = (uint32_t)(kh_hi - kh_lo + 1);
  const uint32_t n0 = min(min(n, c.max_run), (uint32_t)R - S);
  const uint32_t n1 = n - n0;
  uint32_t S1 = S + n0;
  if (S1 >= (uint32_t)R) S1 -= R;
  const uint32_t d0 = c.tbase + S * c.Cout, d1 = c.tbase + S1 * c.Cout;
  const uint32_t rb0 = (uint32_t)kh_lo * c.Cout * c.row16, rb1 = rb0 + n0 * c.Cout * c.row16;
  const uint32_t idesc_r0 = umma_idesc_f16(n0 * c.Cout), idesc_r1 = umma_idesc_f16(n1 * c.Cout);
  if (leader)
  {
    if (fresh)
    {
      // the first contribution to the fresh accumulator (kh = 0 block of run 0) overwrites it
      umma_f16(d0, make_desc(c.hi, a_lo), make_desc(c.hi, c.b_lo + rb0), c.idesc1, 0u);
      if (n0 > 1)
        umma_f16(d0 + c.Cout, make_desc(c.hi, a_lo), make_desc(c.hi, c.b_lo + rb0 + c.Cout * c.row16),
                 umma_idesc_f16((n0 - 1) * c.Cout), 1u);
      if (n1)
        umma_f16(d1, make_desc(c.hi, a_lo), make_desc(c.hi, c.b_lo + rb1), idesc_r1, 1u);
    }
    if (n1)
    {
      if (c.nk == 4)      issue_taps<4, true>(c.hi, a_lo, c.b_lo, c.bblk16, d0, idesc_r0, rb0, d1, idesc_r1, rb1, fresh, c.ntaps);
      else if (c.nk == 2) issue_taps<2, true>(c.hi, a_lo, c.b_lo, c.bblk16, d0, idesc_r0, rb0, d1, idesc_r1, rb1, fresh, c.ntaps);
      else                issue_taps<1, true>(c.hi, a_lo, c.b_lo, c.bblk16, d0, idesc_r0, rb0, d1, idesc_r1, rb1, fresh, c.ntaps);
    }
    else
    {
      if (c.nk == 4)      issue_taps<4, false>(c.hi, a_lo, c.b_lo, c.bblk16, d0, idesc_r0, rb0, d1, idesc_r1, rb1, fresh, c.ntaps);
      else if (c.nk == 2) issue_taps<2, false>(c.hi, a_lo, c.b_lo, c.bblk16, d0, idesc_r0, rb0, d1, idesc_r1, rb1, fresh, c.ntaps);
      else                issue_taps<1, false>(c.hi, a_lo, c.b_lo, c.bblk16, d0, idesc_r0, rb0, d1, idesc_r1, rb1, fresh, c.ntaps);
    }
    umma_commit(release_bar);
  }
  if (r - 1 >= o0 && r - 1 <= o1)
  {
    int m2 = (int)top_mod - 2;
    if (m2 < 0) m2 += R;
    if (leader) umma_commit(c.full_bar + 8 * ((R - 1) - m2));
  }
#ifdef OIDN_B200_TRACE
  tr[tag + 8] += clock64() - ti0_;
#endif
  if (++top_mod == (uint32_t)R) { top_mod = 0; top_par ^= 1; }
}

// Ring bookkeeping of an epilogue: (slot, parity) of consecutive output rows, carried across items.
struct RingPos
{
  uint32_t a_mod = 0, a_par = 0, y_mod = 0, y_par = 0;
  int R;
  __device__ __forceinline__ void begin_item() { y_mod = a_mod; y_par = a_par; }
  __device__ __forceinline__ void next(uint32_t& slot, uint32_t& par)
  {
    slot = (uint32_t)(R - 1) - y_mod; par = y_par;
    if (++y_mod == (uint32_t)R) { y_mod = 0; y_par ^= 1; }
  }
  __device__ __forceinline__ void end_item(int rows)
  {
    const uint32_t tot = a_mod + (uint32_t)rows;
    a_par ^= (tot / (uint32_t)R) & 1;
    a_mod = tot % (uint32_t)R;
  }
};

} // namespace

__global__ void __launch_bounds__(kPairThreads, 1)
conv3x3_pair_kernel(const __grid_constant__ PairKernelParams p)
{
  extern __shared__ uint8_t smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  PTR_DECL

  const int NST = p.nstreams;
  const int nitems = p.nstrips * p.nrowchunks;
  const int cta = blockIdx.x, nctas = gridDim.x;
  const int NA = p.NA, NM = p.NM, RA = p.RA, RB = p.RB;
  const uint32_t w_region = sbase + kSmemHeader;                       // A weights, then B weights
  const uint32_t s_region = w_region + p.w_bytes_smem;                 // per stream: input ring, mid ring
  const uint32_t stream_bytes = (uint32_t)NA * p.a_stage_bytes + (uint32_t)NM * p.mid_stage_bytes;

  // ---------------------------------------------------------------- setup
  if (warp == 0 && lane == 0)
  {
    for (int st = 0; st < NST; ++st)
    {
      for (int s = 0; s < NA; ++s)
      {
        mbar_init(sbase + PairSmem::full_a + 8 * (st * kPairMaxStages + s), 1);
        mbar_init(sbase + PairSmem::empty_a + 8 * (st * kPairMaxStages + s), 1);
      }
      for (int s = 0; s < NM; ++s)
      {
        mbar_init(sbase + PairSmem::mid_full + 8 * (st * kPairMaxMid + s), 4);   // one arrive per A-epilogue warp
        mbar_init(sbase + PairSmem::mid_empty + 8 * (st * kPairMaxMid + s), 1);
      }
      for (int s = 0; s < RA; ++s)
      {
        mbar_init(sbase + PairSmem::acca_full + 8 * (st * kMaxSlots + s), 1);
        mbar_init(sbase + PairSmem::acca_empty + 8 * (st * kMaxSlots + s), 4);
      }
      for (int s = 0; s < RB; ++s)
      {
        mbar_init(sbase + PairSmem::accb_full + 8 * (st * kMaxSlots + s), 1);
        mbar_init(sbase + PairSmem::accb_empty + 8 * (st * kMaxSlots + s), 4);
      }
    }
    mbar_init(sbase + PairSmem::w_full, 1);
    fence_mbar_init();
    prefetch_tmap(&p.amap); prefetch_tmap(&p.wmapA); prefetch_tmap(&p.wmapB);
  }
  if (warp == 1)
  {
    tmem_alloc(sbase + PairSmem::tmem_ptr, kTmemCols);
    tmem_relinquish();
  }
  if (warp >= 4)
  {
    float* ba = reinterpret_cast<float*>(sgen + PairSmem::bias_a);
    float* bb = reinterpret_cast<float*>(sgen + PairSmem::bias_b);
    for (int i = threadIdx.x - 128; i < p.CA; i += kPairThreads - 128) ba[i] = p.biasA[i];
    for (int i = threadIdx.x - 128; i < p.CB; i += kPairThreads - 128) bb[i] = p.biasB[i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_launch_dependents();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(sgen + PairSmem::tmem_ptr);
  const uint32_t tmem_stream_cols = (uint32_t)(RA * p.CA + RB * p.CB);

  if (warp < 4)
  {
    const int st = warp >> 1;
    if (st < NST)
    {
      const int vcta = cta * NST + st, nv = nctas * NST;
      const uint32_t full_a  = sbase + PairSmem::full_a + 8 * st * kPairMaxStages;
      const uint32_t empty_a = sbase + PairSmem::empty_a + 8 * st * kPairMaxStages;
      const uint32_t a_ring  = s_region + (uint32_t)st * stream_bytes;
      const uint32_t m_ring  = a_ring + (uint32_t)NA * p.a_stage_bytes;
      const bool leader = elect_one();
      // -------------------------------------------------------------- TMA producer
      if ((warp & 1) == 0)
      {
        if (st == 0 && leader)
        {
          mbar_arrive_expect_tx(sbase + PairSmem::w_full, p.w_bytes);
          for (int kw = 0; kw < 3; ++kw)
          {
            tma_load_4d(w_region + kw * p.wA_blk, &p.wmapA, sbase + PairSmem::w_full, 0, 0, 0, kw);
            if (kw == 0 || !p.tapB)
              tma_load_4d(w_region + p.wB_off + kw * p.wB_blk, &p.wmapB, sbase + PairSmem::w_full, 0, 0, 0, kw);
          }
        }
        pdl_wait();
        if (p.stamps && st == 0 && leader) atomicMin(&p.stamps[0], globaltimer_ns());
        uint32_t s = 0, ph = 0;
        const uint32_t row_tx = 130u * (uint32_t)p.ccA * 2u;
        for (int item = vcta; item < nitems; item += nv)
        {
          const PItem it = pair_item(p, item);
          for (int r = it.y0 - 2; r <= it.y1 + 2; ++r)
          {
            PWAIT(empty_a + 8 * s, ph ^ 1, 1);
            if (leader)
            {
              mbar_arrive_expect_tx(full_a + 8 * s, row_tx);
              tma_load_3d(a_ring + s * p.a_stage_bytes, &p.amap, full_a + 8 * s, 0, it.x0 - 2, r);
            }
            if (++s == (uint32_t)NA) { s = 0; ph ^= 1; }
          }
        }
      }
      // -------------------------------------------------------------- A's MMA issuer
      else
      {
        const uint32_t sbase16 = (sbase & 0x3FFFFu) >> 4;
        const uint32_t col0 = tmem_base + (uint32_t)st * tmem_stream_cols;
        Side A;
        A.Cout = (uint32_t)p.CA; A.R = (uint32_t)RA; A.tbase = col0; A.max_run = min(3u, 256u / A.Cout);
        A.idesc1 = umma_idesc_f16(A.Cout); A.hi = p.hiA; A.nk = (uint32_t)p.ccA / 16u; A.row16 = A.nk * 2;
        A.b_lo = sbase16 + (kSmemHeader >> 4); A.bblk16 = p.wA_blk >> 4;
        A.ntaps = 3;
        A.full_bar = sbase + PairSmem::acca_full + 8 * st * kMaxSlots;
        A.empty_bar = sbase + PairSmem::acca_empty + 8 * st * kMaxSlots;
        const uint32_t a_ring16 = sbase16 + ((kSmemHeader + p.w_bytes_smem + (uint32_t)st * stream_bytes) >> 4);
        (void)m_ring;

        mbar_wait(sbase + PairSmem::w_full, 0, 2);
        tc_fence_after();

        uint32_t stage = 0, sphase = 0, a_mod = 0, a_par = 0;
        for (int item = vcta; item < nitems; item += nv)
        {
          const PItem it = pair_item(p, item);
          const int ma0 = it.y0 - 1, ma1 = it.y1 + 1;      // mid rows of this item
          uint32_t top_mod = a_mod, top_par = a_par;
          for (int r = ma0 - 1; r <= ma1 + 1; ++r)
          {
            PWAIT(full_a + 8 * stage, sphase, 4);
            tc_fence_after();
            issue_row(A, leader, a_ring16 + ((stage * p.a_stage_bytes) >> 4), r, ma0, ma1, empty_a + 8 * stage, top_mod, top_par, 3 PTR_ARG);
            if (++stage == (uint32_t)NA) { stage = 0; sphase ^= 1; }
          }
          const uint32_t tot = a_mod + (uint32_t)(ma1 - ma0 + 1);
          a_par ^= (tot / (uint32_t)RA) & 1;
          a_mod = tot % (uint32_t)RA;
        }
      }
    }
    __syncwarp();
  }
  // ---------------------------------------------------------------- B's MMA issuer (own warp: a second issuing thread
  // per stream -- a thread issues one tcgen05 instruction per ~70-90 cycles, the narrow MMAs of these layers
  // take 44-56 tensor-pipe cycles, so one issuer per conv keeps the pipe fed where one per stream does not)
  else if (warp >= 20)
  {
    const int st = warp - 20;
    if (st < NST)
    {
      const int vcta = cta * NST + st, nv = nctas * NST;
      const bool leader = elect_one();
      const uint32_t sbase16 = (sbase & 0x3FFFFu) >> 4;
      Side B;
      B.Cout = (uint32_t)p.CB; B.R = (uint32_t)RB;
      B.tbase = tmem_base + (uint32_t)st * tmem_stream_cols + (uint32_t)(RA * p.CA);
      B.max_run = min(3u, 256u / B.Cout);
      B.idesc1 = umma_idesc_f16(B.Cout); B.hi = p.hiB; B.nk = (uint32_t)p.CA / 16u; B.row16 = B.nk * 2;
      B.b_lo = sbase16 + ((kSmemHeader + p.wB_off) >> 4); B.bblk16 = p.wB_blk >> 4;
      B.ntaps = p.tapB ? 1u : 3u;
      B.full_bar = sbase + PairSmem::accb_full + 8 * st * kMaxSlots;
      B.empty_bar = sbase + PairSmem::accb_empty + 8 * st * kMaxSlots;
      const uint32_t m_ring16 = sbase16 + ((kSmemHeader + p.w_bytes_smem + (uint32_t)st * stream_bytes + (uint32_t)NA * p.a_stage_bytes) >> 4);
      const uint32_t mid_full  = sbase + PairSmem::mid_full + 8 * st * kPairMaxMid;
      const uint32_t mid_empty = sbase + PairSmem::mid_empty + 8 * st * kPairMaxMid;
      PWAIT(sbase + PairSmem::w_full, 0, 2);
      tc_fence_after();
      uint32_t ms = 0, mph = 0, b_mod = 0, b_par = 0;
      for (int item = vcta; item < nitems; item += nv)
      {
        const PItem it = pair_item(p, item);
        uint32_t top_mod = b_mod, top_par = b_par;
        for (int m = it.y0 - 1; m <= it.y1 + 1; ++m)
        {
          PWAIT(mid_full + 8 * ms, mph, 6);
          tc_fence_after();
          issue_row(B, leader, m_ring16 + ((ms * p.mid_stage_bytes) >> 4), m, it.y0, it.y1, mid_empty + 8 * ms, top_mod, top_par, 7 PTR_ARG);
          if (++ms == (uint32_t)NM) { ms = 0; mph ^= 1; }
        }
        const uint32_t tot = b_mod + (uint32_t)(it.y1 - it.y0 + 1);
        b_par ^= (tot / (uint32_t)RB) & 1;
        b_mod = tot % (uint32_t)RB;
      }
    }
    __syncwarp();
  }
  // ---------------------------------------------------------------- A epilogue: TMEM -> mid ring
  else if (warp >= 4 && warp < 12)
  {
    const int st = (warp - 4) >> 2;
    if (st < NST)
    {
      const int q = warp & 3;
      const int vcta = cta * NST + st, nv = nctas * NST;
      const int CA = p.CA;
      const uint32_t tfull  = sbase + PairSmem::acca_full + 8 * st * kMaxSlots;
      const uint32_t tempty = sbase + PairSmem::acca_empty + 8 * st * kMaxSlots;
      const uint32_t mid_full  = sbase + PairSmem::mid_full + 8 * st * kPairMaxMid;
      const uint32_t mid_empty = sbase + PairSmem::mid_empty + 8 * st * kPairMaxMid;
      const uint32_t m_ring = s_region + (uint32_t)st * stream_bytes + (uint32_t)NA * p.a_stage_bytes;
      const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)st * tmem_stream_cols;
      const float2* bias2 = reinterpret_cast<const float2*>(sgen + PairSmem::bias_a);
      const int pix = q * 32 + lane;                         // A pixel of this thread inside the strip: x = x0 - 1 + pix
      const uint32_t rowb = (uint32_t)CA * 2u;
      const uint32_t row_off = (uint32_t)pix * rowb;
      const uint32_t xr = ((row_off >> 7) & ((rowb >> 4) - 1u)) << 4;   // swizzle term of this row (16-B chunks)
      const bool relu = p.reluA != 0;
      const bool cbias = p.bias_in_params != 0;   // bias through the constant bank: the shared-memory pipe is the bottleneck
      RingPos ring; ring.R = RA;
      uint32_t ms = 0, mph = 0;
      for (int item = vcta; item < nitems; item += nv)
      {
        const PItem it = pair_item(p, item);
        const int x = it.x0 - 1 + pix;
        const bool xin = x >= 0 && x < p.W;
        ring.begin_item();
        for (int m = it.y0 - 1; m <= it.y1 + 1; ++m)
        {
          uint32_t slot, par;
          ring.next(slot, par);
          PWAIT(tfull + 8 * slot, par, 5);
          tc_fence_after();
          const bool inside = xin && m >= 0 && m < p.H;
          const uint32_t t0 = lane_base + slot * (uint32_t)CA;
          const uint32_t rowaddr = m_ring + ms * p.mid_stage_bytes + row_off;
          // the mid stage must have been read by B's MMAs of kPairMaxMid rows ago
          PWAIT(mid_empty + 8 * ms, mph ^ 1, 8);
#pragma unroll 1
          for (int j = 0; j < CA; j += 32)
          {
            uint32_t v[32], h[16];
            tmem_ld32(t0 + j, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i)
            {
              const float2 bv = cbias ? make_float2(p.biasA_c[j + 2 * i], p.biasA_c[j + 2 * i + 1]) : bias2[j / 2 + i];
              const float2 s = __fadd2_rn(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), bv);
              const uint32_t hv = relu ? pack_half2_relu(s.x, s.y) : pack_half2(s.x, s.y);
              h[i] = inside ? hv : 0u;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
              st_shared_v4(rowaddr + (((uint32_t)(j * 2 + k * 16)) ^ xr), h[4 * k], h[4 * k + 1], h[4 * k + 2], h[4 * k + 3]);
          }
          tc_fence_before();
          fence_proxy_async();       // generic-proxy smem writes -> visible to tcgen05.mma (async proxy)
          __syncwarp();
          if (lane == 0)
          {
            mbar_arrive(tempty + 8 * slot);
            mbar_arrive(mid_full + 8 * ms);
          }
          if (++ms == (uint32_t)NM) { ms = 0; mph ^= 1; }
        }
        ring.end_item(it.y1 - it.y0 + 3);
      }
    }
  }
  // ---------------------------------------------------------------- B epilogue: TMEM -> global
  else if (warp >= 12 && warp < 20)
  {
    const int st = (warp - 12) >> 2;
    if (st < NST)
    {
      const int q = warp & 3;
      const int vcta = cta * NST + st, nv = nctas * NST;
      const int CB = p.CB;
      const uint32_t tfull  = sbase + PairSmem::accb_full + 8 * st * kMaxSlots;
      const uint32_t tempty = sbase + PairSmem::accb_empty + 8 * st * kMaxSlots;
      const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)st * tmem_stream_cols + (uint32_t)(RA * p.CA);
      const float* bias_s = reinterpret_cast<const float*>(sgen + PairSmem::bias_b);
      const float2* bias2 = reinterpret_cast<const float2*>(bias_s);
      const int pix = q * 32 + lane;                         // B pixel inside the strip: x = x0 + pix, valid for pix < 126
      const bool relu = p.reluB != 0;
      const bool cbias = p.bias_in_params != 0;
      RingPos ring; ring.R = RB;
      if (p.fo.enabled)
      {
        // output process in the epilogue (as conv_tc.cu's epilogue_fused_output): channels 0..2 only
        const FusedOutput& fo = p.fo;
        Transfer tf;
        tf.type = fo.tf_type; tf.norm = fo.norm; tf.rcp_norm = fo.rcp_norm;
        tf.input_scale = fo.input_scale; tf.input_scale_ptr = fo.input_scale_ptr;
        pdl_wait();
        const float oscale = output_scale(tf);
        const bool hdr = fo.hdr != 0, snorm = fo.snorm != 0;
        const float b0 = bias_s[0], b1 = bias_s[1], b2 = bias_s[2];
        if (p.tapB)
        {
          // Tap-packed last conv (3 output channels): ONE unshifted view, column kw * 3 + c of a row's accumulator holds
          // the tap-kw partial sum P[i][kw][c] of the mid pixel i = this TMEM lane. Output pixel i (x = x0 - 1 + i,
          // i = 1 .. 126) = P[i-1][0] + P[i][1] + P[i+1][2]: neighbours by warp shuffles, across the warps of the
          // strip through 3 floats of shared memory per side (double-buffered by row, one named barrier per row).
          float* const xbase = reinterpret_cast<float*>(sgen + PairSmem::xch) + st * 64;
          uint32_t rowpar = 0;
          for (int item = vcta; item < nitems; item += nv)
          {
            const PItem it = pair_item(p, item);
            const int x = it.x0 - 1 + pix;
            const int xi = x - fo.wSrc;
            const bool xok = pix >= 1 && pix <= kPairStrip && x < p.W && xi >= 0 && xi < fo.W;
            ring.begin_item();
            for (int y = it.y0; y <= it.y1; ++y)
            {
              uint32_t slot, par;
              ring.next(slot, par);
              PWAIT(tfull + 8 * slot, par, 9);
              tc_fence_after();
              uint32_t v[16];
              tmem_ld16(lane_base + slot * (uint32_t)CB, v);
              tmem_ld_wait();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(tempty + 8 * slot);
              float l0 = __shfl_up_sync(0xffffffffu, __uint_as_float(v[0]), 1);
              float l1 = __shfl_up_sync(0xffffffffu, __uint_as_float(v[1]), 1);
              float l2 = __shfl_up_sync(0xffffffffu, __uint_as_float(v[2]), 1);
              float r0 = __shfl_down_sync(0xffffffffu, __uint_as_float(v[6]), 1);
              float r1 = __shfl_down_sync(0xffffffffu, __uint_as_float(v[7]), 1);
              float r2 = __shfl_down_sync(0xffffffffu, __uint_as_float(v[8]), 1);
              float* const xb = xbase + rowpar * 32;
              if (lane == 31) { xb[q * 8 + 0] = __uint_as_float(v[0]); xb[q * 8 + 1] = __uint_as_float(v[1]); xb[q * 8 + 2] = __uint_as_float(v[2]); }
              if (lane == 0)  { xb[q * 8 + 4] = __uint_as_float(v[6]); xb[q * 8 + 5] = __uint_as_float(v[7]); xb[q * 8 + 6] = __uint_as_float(v[8]); }
              named_bar_sync(1 + st, 128);
              if (lane == 0 && q > 0)  { l0 = xb[(q - 1) * 8 + 0]; l1 = xb[(q - 1) * 8 + 1]; l2 = xb[(q - 1) * 8 + 2]; }
              if (lane == 31 && q < 3) { r0 = xb[(q + 1) * 8 + 4]; r1 = xb[(q + 1) * 8 + 5]; r2 = xb[(q + 1) * 8 + 6]; }
              rowpar ^= 1;
              const int yi = y - fo.hSrc;
              if (xok && yi >= 0 && yi < fo.H)
              {
                const float s0 = ((l0 + __uint_as_float(v[3])) + r0) + b0, s1 = ((l1 + __uint_as_float(v[4])) + r1) + b1,
                            s2 = ((l2 + __uint_as_float(v[5])) + r2) + b2;
                const uint32_t h01 = relu ? pack_half2_relu(s0, s1) : pack_half2(s0, s1);
                const uint32_t h2x = relu ? pack_half2_relu(s2, 0.f) : pack_half2(s2, 0.f);
                const __half2 q01 = *reinterpret_cast<const __half2*>(&h01), q2x = *reinterpret_cast<const __half2*>(&h2x);
                const float3 o = output_pixel(tf, hdr, snorm, false, oscale, __low2float(q01), __high2float(q01), __low2float(q2x));
                float* d = reinterpret_cast<float*>(fo.ptr + (long long)(yi + fo.hDst) * fo.rs) + (size_t)(xi + fo.wDst) * 3;
                d[0] = o.x; d[1] = o.y; d[2] = o.z;
              }
            }
            ring.end_item(it.y1 - it.y0 + 1);
          }
        }
        else
        {
          for (int item = vcta; item < nitems; item += nv)
          {
            const PItem it = pair_item(p, item);
            const int x = it.x0 + pix;
            const int xi = x - fo.wSrc;
            const bool xok = pix < kPairStrip && x < p.W && xi >= 0 && xi < fo.W;
            ring.begin_item();
            for (int y = it.y0; y <= it.y1; ++y)
            {
              uint32_t slot, par;
              ring.next(slot, par);
              PWAIT(tfull + 8 * slot, par, 9);
              tc_fence_after();
              uint32_t v[4];
              tmem_ld4(lane_base + slot * (uint32_t)CB, v);
              tmem_ld_wait();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(tempty + 8 * slot);
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
            ring.end_item(it.y1 - it.y0 + 1);
          }
        }
      }
      else
      {
        const bool pool = p.poolB != 0;
        __half* const gout = static_cast<__half*>(p.out_ptr);
        for (int item = vcta; item < nitems; item += nv)
        {
          const PItem it = pair_item(p, item);
          // pooled: lanes (2i, 2i+1) hold one output pixel, the even lane writes it
          const int xo = pool ? ((it.x0 + pix) >> 1) : (it.x0 + pix);
          const bool writer = pix < kPairStrip && (it.x0 + pix) < p.W && (!pool || (lane & 1) == 0);
          ring.begin_item();
          for (int y = it.y0; y <= it.y1; y += (pool ? 2 : 1))
          {
            uint32_t slot0, par0, slot1 = 0, par1 = 0;
            ring.next(slot0, par0);
            if (pool) ring.next(slot1, par1);
            PWAIT(tfull + 8 * slot0, par0, 9);
            if (pool) PWAIT(tfull + 8 * slot1, par1, 10);
            tc_fence_after();
            __half* const gpix = gout + ((size_t)(pool ? (y >> 1) : y) * p.out_W + xo) * p.CoutPadB;
#pragma unroll 1
            for (int j = 0; j < CB; j += 16)
            {
              uint32_t v[16], h[8];
              tmem_ld16(lane_base + slot0 * (uint32_t)CB + j, v);
              if (pool)
              {
                uint32_t w[16];
                tmem_ld16(lane_base + slot1 * (uint32_t)CB + j, w);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(fmaxf(__uint_as_float(v[i]), __uint_as_float(w[i])));
              }
              else
                tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 8; ++i)
              {
                const float2 bv = cbias ? make_float2(p.biasB_c[j + 2 * i], p.biasB_c[j + 2 * i + 1]) : bias2[j / 2 + i];
                const float2 s = __fadd2_rn(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), bv);
                h[i] = relu ? pack_half2_relu(s.x, s.y) : pack_half2(s.x, s.y);
              }
              if (pool)
              {
#pragma unroll
                for (int i = 0; i < 8; ++i) h[i] = max_half2(h[i], __shfl_xor_sync(0xffffffffu, h[i], 1));
              }
              if (writer && j < p.CoutPadB) st_global_32B(gpix + j, h);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0)
            {
              mbar_arrive(tempty + 8 * slot0);
              if (pool) mbar_arrive(tempty + 8 * slot1);
            }
          }
          ring.end_item(it.y1 - it.y0 + 1);
        }
      }
    }
  }

  // ---------------------------------------------------------------- teardown
  PTR_FLUSH();
  tc_fence_before();
  __syncthreads();
  if (p.stamps && threadIdx.x == 0) atomicMax(&p.stamps[1], globaltimer_ns());
  if (warp == 1)
  {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

cudaError_t conv3x3_pair_launch(const PairKernelParams& p, int grid, size_t smem_bytes, cudaStream_t stream)
{
  static bool attr_set[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 64 && !attr_set[dev])
  {
    e = cudaFuncSetAttribute(conv3x3_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
    if (e != cudaSuccess) return e;
    attr_set[dev] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kPairThreads);
  cfg.dynamicSmemBytes = smem_bytes; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, conv3x3_pair_kernel, p);
}

} // namespace oidnb200
