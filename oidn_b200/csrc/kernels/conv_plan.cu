// Host side of the convolution op: work decomposition, shared-memory budget, TMA tensor maps,
// weight packing, and the C ABI (include/oidn_b200_kernels.h). Also holds the plain SIMT version
// of the same op that the GPU tests use as a second witness.
#include "conv_common.h"
#include "common.h"
#include "transfer.cuh"
#include "../../../include/oidn_b200_kernels.h"
#include <cuda_fp16.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace oidnb200 {

cudaError_t conv3x3_tc_launch(const ConvKernelParams& p, int grid, size_t smem_bytes, cudaStream_t stream);
cudaError_t conv3x3_pair_launch(const PairKernelParams& p, int grid, size_t smem_bytes, cudaStream_t stream);

// ------------------------------------------------------------------------------------------------
// Driver entry point for tensor-map encoding (no link-time libcuda dependency, so the library
// still loads on a machine without a driver).
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode()
{
  static PFN_encodeTiled fn = nullptr;
  if (!fn)
  {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

static CUtensorMapSwizzle swizzle_for(int cc)
{
  return cc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (cc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

int encode_tmap(CUtensorMap* m, const void* base, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes /* rank-1 */, const uint32_t* box, int cc)
{
  PFN_encodeTiled enc = get_encode();
  if (!enc)
  {
    set_error("cuTensorMapEncodeTiled driver entry point not available");
    return OIDNB200_ERR_DRIVER;
  }
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i)
  {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
  }
  for (int i = 0; i < rank - 1; ++i) gstr[i] = strides_bytes[i];
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim,
                   gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(cc),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
  {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return OIDNB200_ERR_DRIVER;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Plan
// ------------------------------------------------------------------------------------------------
struct ConvPlan
{
  oidnb200_conv_desc desc;
  ConvKernelParams kp;
  int grid = 0;
  size_t smem = 0;
  int CinTot = 0, CoutAlloc = 0;
  int chunk_src[kMaxChunks];
  const void *src1 = nullptr, *src2 = nullptr, *weights = nullptr;
  void* dst = nullptr;
  bool bound = false;
  CUtensorMap wmap_tap;      // the tap-packed copy of the weights (can_tap_pack), valid once bound
};

// A 3-channel last conv (dec_conv0 / dec_conv1c) inside a fused pair can take its three horizontal taps as COLUMNS of
// one MMA instead of three shifted views: per vertical tap a 16-column block holds column kw * 3 + c. A third of the
// MMAs and of their shared-memory operand reads (the pair kernel's bottleneck); the pair's epilogue adds the three
// partial sums of neighbouring pixels. The packed weights carry that copy after the regular layout: [kh][16][C1].
static bool can_tap_pack(const ConvPlan& pl)
{
  const oidnb200_conv_desc& d = pl.desc;
  return d.Cout == 16 && pl.CoutAlloc == 16 && d.C2 == 0 && !d.src1_upsampled && d.post_op == POST_NONE &&
         pl.kp.nchunks == 1 && pl.kp.ngroups == 1 && (d.C1 == 32 || d.C1 == 64);
}

static int num_sms()
{
  static int n = 0;
  if (!n)
  {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
    {
      cudaGetLastError();
      n = 148; // B200; lets the planner run (and be unit-tested) without a GPU
    }
  }
  return n;
}

static uint32_t align_up(uint32_t v, uint32_t a) { return (v + a - 1) / a * a; }

static int plan_create_impl(const oidnb200_conv_desc& d, ConvPlan& pl, int want_streams, bool allow_fold);

// Row folding of an upsampled src1 (ConvKernelParams::up_fold) pays when the larger resident weights (4 stacked rows
// per tap instead of 3) neither shrink the output-channel group, nor cost the second stream, nor leave the input rings
// shallow: plan both ways and keep the folded plan only then (dec_conv1a of the three nets).
static int plan_create(const oidnb200_conv_desc& d, ConvPlan& pl, int want_streams = 0 /* 0 = the planner's rule */)
{
  const int rc = plan_create_impl(d, pl, want_streams, false);
  if (rc || !d.src1_upsampled || getenv("OIDN_B200_NO_UPFOLD")) return rc;
  ConvPlan folded;
  if (plan_create_impl(d, folded, want_streams, true) == 0 && folded.kp.up_fold &&
      (getenv("OIDN_B200_FORCE_UPFOLD") ||
       (folded.kp.CoutG == pl.kp.CoutG && folded.kp.nstreams == pl.kp.nstreams)))
    pl = folded;
  return 0;
}

static int plan_create_impl(const oidnb200_conv_desc& d, ConvPlan& pl, int want_streams, bool allow_fold)
{
  if (d.H <= 0 || d.W <= 0 || d.C1 <= 0 || d.C1 % 16 || d.C2 < 0 || d.C2 % 16 || d.Cout <= 0 || d.Cout % 16)
  {
    set_error("conv: channel counts must be positive multiples of 16 and H,W > 0");
    return OIDNB200_ERR_INVALID;
  }
  if (d.post_op == POST_POOL && ((d.H & 1) || (d.W & 1)))
  {
    set_error("conv: PostOp::Pool needs even H and W");
    return OIDNB200_ERR_INVALID;
  }
  if (d.shift_mode != 0)
  {
    set_error("conv: shift_mode 1/2 were hardware-probe variants (profiles/probe_r01.md); only 0 exists");
    return OIDNB200_ERR_INVALID;
  }
  if (d.src1_upsampled && ((d.H & 1) || (d.W & 1)))
  {
    set_error("conv: upsampled source needs even H and W");
    return OIDNB200_ERR_INVALID;
  }
  pl.desc = d;
  ConvKernelParams& kp = pl.kp;
  memset(&kp, 0, sizeof(kp));

  // K chunks: each source is cut into 64/32/16-channel pieces (row bytes 128/64/32).
  int n = 0;
  const int Cs[2] = {d.C1, d.C2};
  int wc0 = 0;
  for (int s = 0; s < 2; ++s)
  {
    int c0 = 0, rem = Cs[s];
    while (rem > 0)
    {
      const int cc = rem >= 64 ? 64 : (rem >= 32 ? 32 : 16);
      if (n >= kMaxChunks)
      {
        set_error("conv: too many K chunks");
        return OIDNB200_ERR_UNSUPPORTED;
      }
      kp.chunk_c0[n] = c0;
      kp.chunk_wc0[n] = wc0;
      kp.chunk_cc[n] = cc;
      kp.chunk_up[n] = (s == 0 && d.src1_upsampled) ? 1 : 0;
      pl.chunk_src[n] = s;
      c0 += cc; wc0 += cc; rem -= cc; ++n;
    }
  }
  kp.nchunks = n;
  pl.CinTot = d.C1 + d.C2;

  // Output-channel group: the largest one whose resident weights + output staging leave room for
  // at least two A stages.
  // One A stage holds one input row of one K chunk: 132 pixels (130 + the 2 extra of an upsampled
  // source) x the widest chunk. Narrow layers get small stages, hence a deep TMA prefetch ring:
  // keeping HBM busy needs ~10 MB of loads in flight chip-wide (bandwidth x latency).
  int maxcc = 16;
  for (int c = 0; c < n; ++c) maxcc = std::max(maxcc, kp.chunk_cc[c]);
  const uint32_t stage_bytes = align_up(132u * maxcc * 2u, 1024);
  kp.stage_bytes = stage_bytes;
  const int min_stages = 2;
  const uint32_t avail = kSmemBudget - 1024 /*alignment slack*/ - kSmemHeader;
  const bool fold = d.src1_upsampled && allow_fold;   // ConvKernelParams::up_fold
  auto b_bytes = [&](int CoutG) {
    uint32_t off = 0;
    for (int c = 0; c < n; ++c)
      off = align_up(off, 1024) + 3u * ((fold && kp.chunk_up[c] ? 4u : 3u) * CoutG * kp.chunk_cc[c] * 2u);
    return align_up(off, 1024);
  };
  // Output staging: every epilogue warp (2 warpgroups x 4) owns a slice for its 32 pixels (16 after
  // pooling), cut into pieces of 64/32/16 channels, each piece 1024-aligned (swizzle atoms).
  const uint32_t out_rows = (d.post_op == POST_POOL) ? 16u : 32u;
  auto out_bytes = [&](int CoutG) { // one staging slice of one warp
    uint32_t off = 0;
    for (int rem = CoutG; rem > 0;)
    {
      const int cc = rem >= 64 ? 64 : (rem >= 32 ? 32 : 16);
      off = align_up(off, 1024) + out_rows * cc * 2u;
      rem -= cc;
    }
    return align_up(off, 1024);
  };
  constexpr uint32_t kEpiWarps = 8;   // sizing of CoutG assumes the two-stream kernel with staged stores
  int CoutG = std::min(d.Cout, 128);
  while (CoutG > 16 && b_bytes(CoutG) + kEpiWarps * out_bytes(CoutG) + min_stages * stage_bytes > avail) CoutG -= 16;
  if (b_bytes(CoutG) + kEpiWarps * out_bytes(CoutG) + min_stages * stage_bytes > avail)
  {
    set_error("conv: weights for 16 output channels do not fit in shared memory");
    return OIDNB200_ERR_UNSUPPORTED;
  }
  const int ngroups = (d.Cout + CoutG - 1) / CoutG;
  CoutG = ((d.Cout + ngroups - 1) / ngroups + 15) / 16 * 16; // rebalance
  kp.CoutG = CoutG;
  kp.ngroups = ngroups;
  kp.CoutPad = d.Cout;
  pl.CoutAlloc = CoutG * ngroups;

  uint32_t off = 0;
  for (int c = 0; c < n; ++c)
  {
    off = align_up(off, 1024);
    kp.chunk_boff[c] = off;
    kp.chunk_bblk[c] = (fold && kp.chunk_up[c] ? 4u : 3u) * CoutG * kp.chunk_cc[c] * 2u;
    {
      const uint32_t rowb = (uint32_t)kp.chunk_cc[c] * 2u;   // bytes of one staged pixel / weight row
      const uint32_t layout = rowb == 128 ? 2u : (rowb == 64 ? 4u : 6u); // 128B / 64B / 32B swizzle
      kp.chunk_nk[c] = (uint32_t)kp.chunk_cc[c] / 16u;
      // sm_100 descriptor high word: SBO = 8 rows (bits 32..45), version 1 (bit 46), swizzle (bits 61..63)
      kp.chunk_hi[c] = (((8u * rowb) >> 4) & 0x3FFFu) | (1u << 14) | (layout << 29);
      kp.chunk_a16[c] = kp.chunk_up[c] ? rowb >> 4 : 0u;
      kp.chunk_b16[c] = off >> 4;
      kp.chunk_bblk16[c] = kp.chunk_bblk[c] >> 4;
    }
    off += 3u * kp.chunk_bblk[c];
  }
  const uint32_t bbytes = align_up(off, 1024);
  kp.b_bytes = bbytes;
  kp.w_bytes = 0;
  for (int c = 0; c < n; ++c) kp.w_bytes += 3u * kp.chunk_bblk[c];
  kp.up_fold = fold ? 1 : 0;

  // output pieces (same constexpr decomposition the kernel's epilogue is specialised on)
  const int NB = CoutG / 16;
  kp.nout = out_piece_count(NB);
  uint32_t ooff = 0;
  for (int i = 0; i < kp.nout; ++i)
  {
    kp.out_c0[i] = out_piece_c0(NB, i);
    kp.out_cc[i] = out_piece_cc(NB, i);
    kp.out_off[i] = out_piece_off(NB, i, (int)out_rows);
    ooff = kp.out_off[i] + out_rows * kp.out_cc[i] * 2u;
  }
  kp.out_buf_bytes = align_up(ooff, 1024);

  // What is left after the weights and the epilogue's staging slices goes to A stages.
  // A stage of chunk c: 132 pixels x its channels, 1024-aligned. A ring whose stage count is a multiple
  // of the chunk count gives every stage a fixed chunk, so narrow chunks take narrow stages.
  uint32_t chunk_stage[kMaxChunks], row_bytes = 0;
  for (int c = 0; c < n; ++c)
  {
    chunk_stage[c] = align_up(132u * kp.chunk_cc[c] * 2u, 1024);
    row_bytes += chunk_stage[c];
  }
  // folded: two rings in one table -- U stages (src1 chunks: one set per LOW-RES row) then S stages (src2 chunks: one
  // set per row); sized for the same look-ahead in rows: NU = n_up * k, NS2 = n_s2 * 2k, plus what still fits
  int n_up = 0, n_s2 = 0;
  uint32_t su = 1024, ss = 1024;
  for (int c = 0; c < n; ++c)
  {
    if (kp.chunk_up[c]) { ++n_up; su = std::max(su, chunk_stage[c]); }
    else { ++n_s2; ss = std::max(ss, chunk_stage[c]); }
  }
  int fold_nu = 0, fold_ns = 0;
  auto ring = [&](uint32_t left, int& nstages, bool& by_row) { // one stream's ring inside `left` bytes
    if (fold)
    {
      by_row = false; nstages = 0; fold_nu = fold_ns = 0;
      int k = 0;
      while ((uint32_t)((k + 1) * (n_up * su + 2 * n_s2 * ss)) <= left && (k + 1) * (n_up + 2 * n_s2) <= kMaxStages) ++k;
      if (k < 1) return;
      fold_nu = n_up * k; fold_ns = 2 * n_s2 * k;
      uint32_t used = (uint32_t)fold_nu * su + (uint32_t)fold_ns * ss;
      while (n_s2 && used + ss <= left && fold_nu + fold_ns < kMaxStages) { ++fold_ns; used += ss; }
      while (used + su <= left && fold_nu + fold_ns < kMaxStages) { ++fold_nu; used += su; }
      nstages = fold_nu + fold_ns;
      return;
    }
    const int uni = std::min((int)(left / stage_bytes), kMaxStages);
    const int rows = std::min((int)(left / row_bytes), kMaxStages / n);
    by_row = rows * n > uni;
    nstages = by_row ? rows * n : uni;
  };
  // Candidates, best first. Streams: a stream serialises "wait for the row, issue its MMAs, commit"
  // on one issuing thread, so two independent streams per CTA keep the tensor pipe busy (CoutG <= 64:
  // two >= 4-slot accumulator rings fit TMEM); every stream needs >= 2 A stages. Measured on the 4K
  // shapes (profiles/r01_probe4_direct_streams.log, r01_probe5_four_streams.log): a second stream is
  // worth 8-10 % on dec_conv2a/3a even with a 2-3 stage ring each; FOUR streams (CoutG <= 32, the
  // 768-thread kernel) gain 6 % on enc_conv0 but lose 3-8 % on enc_conv1 / dec_conv1b / dec_conv0
  // (4-slot rings, 80 registers per thread), so they are opt-in: OIDN_B200_STREAMS=4.
  // Epilogue stores: staged through swizzled shared memory + TMA store, double-buffered if that
  // leaves >= 3 stages per stream, else single-buffered, else straight from registers (frees all
  // staging memory; measured 0-5 % slower at equal stream count, so it is only used to afford
  // another stream). OIDN_B200_STREAMS (max streams) and OIDN_B200_DIRECT_STORE (0/1) override.
  int max_streams = CoutG <= 64 ? 2 : 1;
  if (want_streams == 4 && CoutG <= 32) max_streams = 4;
  if (const char* e = getenv("OIDN_B200_STREAMS"))
    max_streams = std::min(CoutG <= 32 ? 4 : max_streams, std::max(1, atoi(e)));
  int force_direct = -1;
  if (const char* e = getenv("OIDN_B200_DIRECT_STORE")) force_direct = atoi(e);
  bool by_row = false, found = false;
  for (int ns = max_streams; ns >= 1 && !found; ns >>= 1)
  {
    if (ns == 3) continue;
    const uint32_t epi_warps = ns == 4 ? 16u : 8u;
    for (int nbuf = 2; nbuf >= 0 && !found; --nbuf)   // nbuf 0 = direct stores
    {
      if (force_direct == 1 && nbuf != 0) continue;
      if (force_direct == 0 && nbuf == 0) continue;
      const uint32_t fixed = bbytes + epi_warps * nbuf * kp.out_buf_bytes;
      if (fixed + (uint32_t)ns * 2u * stage_bytes > avail) continue;
      int nst; bool br;
      ring(((avail - fixed) / ns) & ~1023u, nst, br);
      const int want = fold ? 2 * n_up + 2 * n_s2 : (nbuf == 2 ? 3 : 2);   // folded: two low-res rows + two rows in flight
      if (nst < want) continue;
      kp.nstreams = ns; kp.nstages = nst; kp.out_nbuf = nbuf; by_row = br; found = true;
    }
  }
  if (!found)
  {
    set_error("conv: no shared-memory configuration fits");
    return OIDNB200_ERR_UNSUPPORTED;
  }
  kp.direct_store = kp.out_nbuf == 0 ? 1 : 0;
  const uint32_t epi_warps = kp.nstreams == 4 ? 16u : 8u;
  {
    uint32_t off = 0;
    for (int s2 = 0; s2 < kp.nstages; ++s2)
    {
      kp.stage_off[s2] = off;
      off += fold ? (s2 < fold_nu ? su : ss) : (by_row ? chunk_stage[s2 % n] : stage_bytes);
    }
    kp.ring_bytes = off;
    kp.nstages_u = fold ? fold_nu : 0;
  }
  kp.R = std::min(kMaxSlots, (kTmemCols / kp.nstreams) / CoutG);
  // A ring that cannot hold even one input row (all K chunks) does not hide HBM latency: the TMA
  // producer then pulls the rows ahead into L2 first. Measured (profiles/r01_prefetch_ab.log):
  // dec_conv2a 0.272 -> 0.250 ms, dec_conv3a 0.144 -> 0.139 ms; deeper rings lose ~1 %.
  // OIDN_B200_PREFETCH overrides the choice (hardware probing only).
  kp.prefetch_rows = (kp.nstages < n && !fold) ? 2 : 0;
  if (const char* e = getenv("OIDN_B200_PREFETCH")) kp.prefetch_rows = atoi(e);
  pl.smem = 1024 + kSmemHeader + (size_t)kp.nstreams * kp.ring_bytes + bbytes +
            (size_t)epi_warps * kp.out_nbuf * kp.out_buf_bytes;

  // Work decomposition: strips of 128 px x RC rows; pick RC minimising the critical path.
  kp.H = d.H; kp.W = d.W;
  kp.nstrips = (d.W + kStripW - 1) / kStripW;
  const int Pc = std::max(1, num_sms() / ngroups);      // physical CTAs per output-channel group
  const int P = Pc * kp.nstreams;                        // independent row streams per group
  const int step = (d.post_op == POST_POOL) ? 2 : 1;
  int bestRC = step;
  long bestCost = -1;
  for (int RC = step; RC <= d.H + step - 1; RC += step)
  {
    const long items = (long)kp.nstrips * ((d.H + RC - 1) / RC);
    const long waves = (items + P - 1) / P;
    const long cost = waves * (RC + 2);
    if (bestCost < 0 || cost < bestCost || (cost == bestCost && RC > bestRC))
    {
      bestCost = cost;
      bestRC = RC;
    }
  }
  kp.RC = bestRC;
  kp.nrowchunks = (d.H + bestRC - 1) / bestRC;
  const int items = kp.nstrips * kp.nrowchunks;
  pl.grid = std::min(Pc, (items + kp.nstreams - 1) / kp.nstreams) * ngroups;

  kp.relu = d.relu;
  kp.post_op = d.post_op;
  return 0;
}

static int plan_bind(ConvPlan& pl, const void* src1, const void* src2, const void* weights,
                     const void* bias, void* dst)
{
  const oidnb200_conv_desc& d = pl.desc;
  ConvKernelParams& kp = pl.kp;
  if (!src1 || (d.C2 > 0 && !src2) || !weights || !bias || !dst)
  {
    set_error("conv: null pointer in bind");
    return OIDNB200_ERR_INVALID;
  }
  for (int c = 0; c < kp.nchunks; ++c)
  {
    const int s = pl.chunk_src[c];
    const int C = s == 0 ? d.C1 : d.C2;
    const void* base = s == 0 ? src1 : src2;
    const int cc = kp.chunk_cc[c];
    int rc;
    if (kp.chunk_up[c])
    {
      const uint64_t Wl = d.W / 2, Hl = d.H / 2;
      const uint64_t dims[4] = {(uint64_t)C, 2, Wl, Hl};
      const uint64_t str[3] = {0, (uint64_t)C * 2, Wl * C * 2};
      const uint32_t box[4] = {(uint32_t)cc, 2, 66, 1};
      rc = encode_tmap(&kp.amap[c], base, 4, dims, str, box, cc);
    }
    else
    {
      const uint64_t dims[3] = {(uint64_t)C, (uint64_t)d.W, (uint64_t)d.H};
      const uint64_t str[2] = {(uint64_t)C * 2, (uint64_t)d.W * C * 2};
      const uint32_t box[3] = {(uint32_t)cc, 130u, 1};
      rc = encode_tmap(&kp.amap[c], base, 3, dims, str, box, cc);
    }
    if (rc) return rc;
    if (kp.up_fold && kp.chunk_up[c])
    {
      // the folded weights of src1 follow the regular ones: [kw][E0, E1, E2, O][CoutAlloc][C1]
      const uint8_t* wf = static_cast<const uint8_t*>(weights) + (size_t)9 * pl.CoutAlloc * pl.CinTot * 2;
      const uint64_t wd[4] = {(uint64_t)d.C1, (uint64_t)pl.CoutAlloc, 4, 3};
      const uint64_t ws[3] = {(uint64_t)d.C1 * 2, (uint64_t)pl.CoutAlloc * d.C1 * 2, (uint64_t)4 * pl.CoutAlloc * d.C1 * 2};
      const uint32_t wb[4] = {(uint32_t)cc, (uint32_t)kp.CoutG, 4, 1};
      rc = encode_tmap(&kp.wmap[c], wf, 4, wd, ws, wb, cc);
    }
    else
    {
      if (can_tap_pack(pl))
      {
        const uint8_t* wt = static_cast<const uint8_t*>(weights) + (size_t)9 * pl.CoutAlloc * pl.CinTot * 2;
        const uint64_t td[4] = {(uint64_t)d.C1, 16, 3, 1};
        const uint64_t ts[3] = {(uint64_t)d.C1 * 2, (uint64_t)16 * d.C1 * 2, (uint64_t)48 * d.C1 * 2};
        const uint32_t tb[4] = {(uint32_t)cc, 16, 3, 1};
        rc = encode_tmap(&pl.wmap_tap, wt, 4, td, ts, tb, cc);
        if (rc) return rc;
      }
      const uint64_t wd[4] = {(uint64_t)pl.CinTot, (uint64_t)pl.CoutAlloc, 3, 3};
      const uint64_t ws[3] = {(uint64_t)pl.CinTot * 2, (uint64_t)pl.CoutAlloc * pl.CinTot * 2,
                              (uint64_t)3 * pl.CoutAlloc * pl.CinTot * 2};
      const uint32_t wb[4] = {(uint32_t)cc, (uint32_t)kp.CoutG, 3, 1};
      rc = encode_tmap(&kp.wmap[c], weights, 4, wd, ws, wb, cc);
    }
    if (rc) return rc;
  }
  {
    const uint64_t Wd = d.post_op == POST_POOL ? d.W / 2 : d.W, Hd = d.post_op == POST_POOL ? d.H / 2 : d.H;
    const uint64_t od[3] = {(uint64_t)d.Cout, Wd, Hd};
    const uint64_t os[2] = {(uint64_t)d.Cout * 2, Wd * d.Cout * 2};
    for (int oc = 0; oc < kp.nout; ++oc)
    {
      const uint32_t ob[3] = {(uint32_t)kp.out_cc[oc], d.post_op == POST_POOL ? 16u : 32u, 1}; // one warp's pixels
      const int rc = encode_tmap(&kp.omap[oc], dst, 3, od, os, ob, kp.out_cc[oc]);
      if (rc) return rc;
    }
  }
  pl.dst = dst;
  kp.out_ptr = dst;
  kp.out_W = d.post_op == POST_POOL ? d.W / 2 : d.W;
  kp.bias = static_cast<const float*>(bias);
  pl.src1 = src1; pl.src2 = src2; pl.weights = weights;
  pl.bound = true;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// SIMT witness: one thread per (pixel, output channel), fp32 accumulation in kh,kw,ci order.
// ------------------------------------------------------------------------------------------------
__global__ void conv3x3_simt_kernel(const __half* __restrict__ src1, const __half* __restrict__ src2,
                                    int C1, int C2, int up1, int H, int W,
                                    const __half* __restrict__ wts, int CinTot, int CoutAlloc,
                                    const float* __restrict__ bias, int Cout, int relu,
                                    __half* __restrict__ out /* [H][W][Cout] */)
{
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)H * W * Cout) return;
  const int co = (int)(idx % Cout);
  const int x = (int)((idx / Cout) % W);
  const int y = (int)(idx / ((long)Cout * W));
  float acc = bias[co];
  for (int kh = 0; kh < 3; ++kh)
    for (int kw = 0; kw < 3; ++kw)
    {
      const int yy = y + kh - 1, xx = x + kw - 1;
      if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
      const __half* w = wts + ((size_t)(kw * 3 + kh) * CoutAlloc + co) * CinTot;
      const __half* s1 = up1 ? src1 + ((size_t)(yy >> 1) * (W >> 1) + (xx >> 1)) * C1
                             : src1 + ((size_t)yy * W + xx) * C1;
      for (int ci = 0; ci < C1; ++ci) acc += __half2float(w[ci]) * __half2float(s1[ci]);
      if (C2 > 0)
      {
        const __half* s2 = src2 + ((size_t)yy * W + xx) * C2;
        for (int ci = 0; ci < C2; ++ci) acc += __half2float(w[C1 + ci]) * __half2float(s2[ci]);
      }
    }
  if (relu) acc = fmaxf(acc, 0.f);
  out[idx] = __float2half_rn(acc);
}

__global__ void post_simt_kernel(const __half* __restrict__ in, int H, int W, int C, int post_op,
                                 __half* __restrict__ dst)
{
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (post_op == POST_POOL)
  {
    const int Ho = H / 2, Wo = W / 2;
    if (idx >= (long)Ho * Wo * C) return;
    const int c = (int)(idx % C);
    const int x = (int)((idx / C) % Wo);
    const int y = (int)(idx / ((long)C * Wo));
    float m = -INFINITY;
    for (int dy = 0; dy < 2; ++dy)
      for (int dx = 0; dx < 2; ++dx)
        m = fmaxf(m, __half2float(in[((size_t)(2 * y + dy) * W + 2 * x + dx) * C + c]));
    dst[idx] = __float2half_rn(m);
  }
  else if (post_op == POST_UPSAMPLE)
  {
    const int Ho = H * 2, Wo = W * 2;
    if (idx >= (long)Ho * Wo * C) return;
    const int c = (int)(idx % C);
    const int x = (int)((idx / C) % Wo);
    const int y = (int)(idx / ((long)C * Wo));
    dst[idx] = in[((size_t)(y >> 1) * W + (x >> 1)) * C + c];
  }
  else
  {
    if (idx >= (long)H * W * C) return;
    dst[idx] = in[idx];
  }
}

} // namespace oidnb200

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
using namespace oidnb200;

struct oidnb200_conv
{
  ConvPlan plan;
  // The network's last conv (16 padded output channels) with the output process in its epilogue is
  // bound by the per-pixel math of the epilogue warps, not by MMA issue or HBM: the four-stream
  // kernel has 16 epilogue warps instead of 8 (measured in the 4K frame: dec_conv0 + output 0.166 ->
  // 0.143 ms, while the UNfused conv is 7 % slower with four streams). So a conv that can be fused
  // keeps a second, four-stream plan that oidnb200_conv_launch uses while a fusion is set.
  std::unique_ptr<ConvPlan> fused_plan;
  // the packed bias as oidnb200_conv_pack_bias produced it (host copy): goes into the kernel parameters at bind
  mutable std::vector<float> host_bias;
  mutable int tap_O = 0;     // 3 once oidnb200_conv_pack_weights wrote the tap-packed copy (can_tap_pack, 3 output channels)
};

// Two chained convs as one launch (conv_pair_tc.cu). Refers to the two single-conv ops, which own descriptors,
// packed weights and bindings.
struct oidnb200_conv_pair
{
  const oidnb200_conv* a = nullptr;
  const oidnb200_conv* b = nullptr;
  PairKernelParams kp;
  int grid = 0;
  size_t smem = 0;
  bool bound = false;
  bool tap_ok = false;       // conv B has a tap-packed weight copy: used when its output process is fused
};

static int pair_plan(const oidnb200_conv_desc& da, const ConvPlan& A, const oidnb200_conv_desc& db, const ConvPlan& B,
                     oidnb200_conv_pair& pr)
{
  auto no = [](const char* why) { set_error(std::string("conv pair: ") + why); return OIDNB200_ERR_UNSUPPORTED; };
  if (getenv("OIDN_B200_NO_PAIRS")) return no("disabled by OIDN_B200_NO_PAIRS");
  if (da.H != db.H || da.W != db.W) return no("resolutions differ");
  if (da.C2 != 0 || db.C2 != 0 || da.src1_upsampled || db.src1_upsampled) return no("concat / upsampled sources are not fused");
  if (da.post_op != POST_NONE || (db.post_op != POST_NONE && db.post_op != POST_POOL)) return no("post-op not supported");
  if (A.kp.nchunks != 1 || B.kp.nchunks != 1 || A.kp.ngroups != 1 || B.kp.ngroups != 1) return no("more than one K chunk / channel group");
  if (db.C1 != da.Cout || (da.Cout != 32 && da.Cout != 64) || db.Cout > 64) return no("channel counts not covered");
  PairKernelParams& kp = pr.kp;
  memset(&kp, 0, sizeof(kp));
  kp.H = da.H; kp.W = da.W;
  kp.ccA = da.C1; kp.CA = da.Cout; kp.CB = db.Cout;
  kp.poolB = db.post_op == POST_POOL; kp.reluA = da.relu; kp.reluB = db.relu;
  kp.a_stage_bytes = align_up(130u * kp.ccA * 2u, 1024);
  kp.mid_stage_bytes = align_up(130u * kp.CA * 2u, 1024);
  kp.wA_blk = 3u * kp.CA * kp.ccA * 2u;
  kp.wB_blk = 3u * kp.CB * kp.CA * 2u;
  if (kp.wA_blk % 1024 || kp.wB_blk % 1024) return no("weight blocks are not 1024-byte multiples");
  kp.wB_off = 3u * kp.wA_blk;
  kp.w_bytes = 3u * (kp.wA_blk + kp.wB_blk);
  kp.w_bytes_smem = align_up(kp.w_bytes, 1024);
  kp.hiA = A.kp.chunk_hi[0];
  kp.hiB = B.kp.chunk_hi[0];
  const uint32_t avail = kSmemBudget - 1024 - kSmemHeader - kp.w_bytes_smem;
  // Streams x accumulator rings x shared-memory rings (measured on the 4K shapes, tools/run_probe_pair_sweep.sh,
  // profiles/r02_probe_pair_sweep.log: two streams beat one deep stream by 4-6 %; a pooling B wants one more slot
  // than A -- its epilogue drains two rows at a time -- so A runs on three; ring depths beyond that change < 1 %).
  // OIDN_B200_PAIR_{STREAMS,RA,RB,NM,NA} override (hardware probing only).
  auto envi = [](const char* n, int dflt) { const char* e = getenv(n); return e ? atoi(e) : dflt; };
  int nst = 0;
  for (int cand = envi("OIDN_B200_PAIR_STREAMS", 2); cand >= 1 && !nst; --cand)
  {
    const int cols = kTmemCols / cand;
    int ra = envi("OIDN_B200_PAIR_RA", (cand == 2 && (kp.poolB || kp.CA == 64)) ? 3 : 4);
    int rb = std::min(envi("OIDN_B200_PAIR_RB", 8), (cols - ra * kp.CA) / kp.CB);
    if (rb < 4) continue;
    for (int nm = envi("OIDN_B200_PAIR_NM", 4); nm >= 3 && !nst; --nm)
    {
      const long left = (long)avail / cand - (long)nm * kp.mid_stage_bytes;
      const int na = (int)std::min<long>(envi("OIDN_B200_PAIR_NA", kPairMaxStages), left / (long)kp.a_stage_bytes);
      if (na >= 3) { nst = cand; kp.nstreams = cand; kp.RA = ra; kp.RB = rb; kp.NM = nm; kp.NA = na; }
    }
  }
  if (!nst) return no("accumulator / shared-memory rings do not fit");
  pr.smem = 1024 + kSmemHeader + kp.w_bytes_smem + (size_t)nst * ((size_t)kp.NA * kp.a_stage_bytes + (size_t)kp.NM * kp.mid_stage_bytes);
  kp.nstrips = (da.W + kPairStrip - 1) / kPairStrip;
  const int P = num_sms() * nst;
  const int step = kp.poolB ? 2 : 1;
  int bestRC = step; long bestCost = -1;
  for (int RC = step; RC <= da.H + step - 1; RC += step)
  {
    const long items = (long)kp.nstrips * ((da.H + RC - 1) / RC);
    const long waves = (items + P - 1) / P;
    const long cost = waves * (RC + 4);
    if (bestCost < 0 || cost < bestCost || (cost == bestCost && RC > bestRC)) { bestCost = cost; bestRC = RC; }
  }
  kp.RC = bestRC;
  kp.nrowchunks = (da.H + bestRC - 1) / bestRC;
  const int items = kp.nstrips * kp.nrowchunks;
  pr.grid = std::min(num_sms(), (items + nst - 1) / nst);
  return 0;
}

extern "C" {

int oidnb200_conv_pair_create(const oidnb200_conv* a, const oidnb200_conv* b, oidnb200_conv_pair** out)
{
  if (!a || !b || !out)
  {
    set_error("conv_pair_create: null argument");
    return OIDNB200_ERR_INVALID;
  }
  oidnb200_conv_pair* pr = new oidnb200_conv_pair();
  pr->a = a; pr->b = b;
  const int rc = pair_plan(a->plan.desc, a->plan, b->plan.desc, b->plan, *pr);
  if (rc)
  {
    delete pr;
    return rc;
  }
  *out = pr;
  return 0;
}

void oidnb200_conv_pair_destroy(oidnb200_conv_pair* pair) { delete pair; }

int oidnb200_conv_pair_bind(oidnb200_conv_pair* pr)
{
  const ConvPlan& A = pr->a->plan;
  const ConvPlan& B = pr->b->plan;
  if (!A.bound || !B.bound)
  {
    set_error("conv_pair_bind: both convolutions must be bound first");
    return OIDNB200_ERR_INVALID;
  }
  PairKernelParams& kp = pr->kp;
  kp.amap = A.kp.amap[0];
  kp.wmapA = A.kp.wmap[0];
  kp.wmapB = B.kp.wmap[0];
  kp.biasA = A.kp.bias;
  kp.biasB = B.kp.bias;
  kp.bias_in_params = 0;
  if (A.kp.bias_in_params && B.kp.bias_in_params)
  {
    memcpy(kp.biasA_c, A.kp.bias_c, sizeof(float) * kp.CA);
    memcpy(kp.biasB_c, B.kp.bias_c, sizeof(float) * kp.CB);
    kp.bias_in_params = 1;
  }
  kp.out_ptr = B.kp.out_ptr;
  kp.out_W = B.kp.out_W;
  kp.CoutPadB = B.kp.CoutPad;
  pr->tap_ok = can_tap_pack(B) && pr->b->tap_O == 3 && !kp.poolB && !getenv("OIDN_B200_NO_TAP_PACK");
  pr->bound = true;
  return 0;
}

int oidnb200_conv_pair_launch(oidnb200_conv_pair* pr, oidnb200_stream stream)
{
  if (!pr->bound)
  {
    set_error("conv_pair_launch: op not bound");
    return OIDNB200_ERR_INVALID;
  }
  pr->kp.fo = pr->b->plan.kp.fo;            // conv B's fused output process (oidnb200_conv_set_output_process)
  {
    PairKernelParams& kp = pr->kp;
    kp.tapB = (pr->tap_ok && kp.fo.enabled) ? 1 : 0;
    kp.wmapB = kp.tapB ? pr->b->plan.wmap_tap : pr->b->plan.kp.wmap[0];
    kp.w_bytes = 3u * kp.wA_blk + (kp.tapB ? 1u : 3u) * kp.wB_blk;
  }
  pr->kp.stamps = pr->b->plan.kp.stamps;    // in-frame interval of the pair under conv B's entry
  pr->kp.trace = pr->b->plan.kp.trace;
  const cudaError_t e = conv3x3_pair_launch(pr->kp, pr->grid, pr->smem, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess)
  {
    set_error(std::string("conv_pair_launch: ") + cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

int oidnb200_conv_pair_get_info(const oidnb200_conv_pair* pr, oidnb200_conv_info* info)
{
  memset(info, 0, sizeof(*info));
  // ngroups / nchunks / out_nbuf carry the pair's own ring sizes: mid stages, A's accumulator slots
  info->grid = pr->grid; info->smem_bytes = (int)pr->smem; info->ngroups = pr->kp.NM; info->cout_group = pr->kp.CB;
  info->nchunks = pr->kp.RA; info->nstages = pr->kp.NA; info->ring_slots = pr->kp.RB; info->rows_per_item = pr->kp.RC;
  info->nstrips = pr->kp.nstrips; info->nrowchunks = pr->kp.nrowchunks; info->nstreams = pr->kp.nstreams; info->out_nbuf = 0;
  return 0;
}

int oidnb200_conv_create(const oidnb200_conv_desc* desc, oidnb200_conv** out)
{
  if (!desc || !out)
  {
    set_error("conv_create: null argument");
    return OIDNB200_ERR_INVALID;
  }
  oidnb200_conv* c = new oidnb200_conv();
  const int rc = plan_create(*desc, c->plan);
  if (rc)
  {
    delete c;
    return rc;
  }
  if (desc->Cout == 16 && desc->post_op == POST_NONE && c->plan.kp.ngroups == 1)
  {
    c->fused_plan.reset(new ConvPlan());
    if (plan_create(*desc, *c->fused_plan, 4) != 0 || c->fused_plan->kp.nstreams != 4)
      c->fused_plan.reset();   // does not fit: the fused conv runs on the default plan
  }
  *out = c;
  return 0;
}

void oidnb200_conv_destroy(oidnb200_conv* conv) { delete conv; }

size_t oidnb200_conv_weight_bytes(const oidnb200_conv* conv)
{
  const ConvPlan& pl = conv->plan;
  size_t n = (size_t)9 * pl.CoutAlloc * pl.CinTot;
  if (pl.kp.up_fold) n += (size_t)12 * pl.CoutAlloc * pl.desc.C1;   // [kw][4][CoutAlloc][C1]: vertically pre-summed src1 weights
  if (can_tap_pack(pl)) n += (size_t)48 * pl.desc.C1;               // [kh][16][C1]: horizontal taps as columns
  return n * sizeof(uint16_t);
}

size_t oidnb200_conv_bias_bytes(const oidnb200_conv* conv)
{
  return (size_t)conv->plan.CoutAlloc * sizeof(float);
}

int oidnb200_conv_pack_weights(const oidnb200_conv* conv, const uint16_t* w_oihw, int O, int I1, int I2,
                               void* dst_weights)
{
  const ConvPlan& pl = conv->plan;
  if (!w_oihw || !dst_weights || O > pl.desc.Cout || I1 > pl.desc.C1 || I2 > pl.desc.C2 || O <= 0 || I1 <= 0 || I2 < 0)
  {
    set_error("conv_pack_weights: logical dims exceed padded dims");
    return OIDNB200_ERR_INVALID;
  }
  uint16_t* dst = static_cast<uint16_t*>(dst_weights);
  memset(dst, 0, oidnb200_conv_weight_bytes(conv));
  const int I = I1 + I2;
  for (int o = 0; o < O; ++o)
    for (int i = 0; i < I; ++i)
    {
      const int ci = i < I1 ? i : pl.desc.C1 + (i - I1); // src2 channels follow src1's padded ones
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw)
          dst[((size_t)(kw * 3 + kh) * pl.CoutAlloc + o) * pl.CinTot + ci] =
            w_oihw[(((size_t)o * I + i) * 3 + kh) * 3 + kw];
    }
  conv->tap_O = 0;
  if (can_tap_pack(pl) && O == 3 && I2 == 0)
  {
    uint16_t* f = dst + (size_t)9 * pl.CoutAlloc * pl.CinTot;
    for (int o = 0; o < O; ++o)
      for (int i = 0; i < I1; ++i)
        for (int kh = 0; kh < 3; ++kh)
          for (int kw = 0; kw < 3; ++kw)
            f[((size_t)kh * 16 + (kw * 3 + o)) * pl.desc.C1 + i] = w_oihw[(((size_t)o * I + i) * 3 + kh) * 3 + kw];
    conv->tap_O = 3;
  }
  if (pl.kp.up_fold)
  {
    // Row folding of the upsampled src1: upsampled rows 2Y and 2Y+1 are both low-res row Y. At the even virtual row
    // r = 2Y the stack (output rows r+1, r, r-1) takes E = [w0+w1; w1+w2; w2] (kh taps that land on row Y), at the odd
    // row r = 2Y+1 the fresh output row r+1 takes O = [w0]. Sums in fp32, rounded to fp16 once.
    uint16_t* f = dst + (size_t)9 * pl.CoutAlloc * pl.CinTot;
    const int C1 = pl.desc.C1;
    for (int o = 0; o < O; ++o)
      for (int i = 0; i < I1; ++i)
        for (int kw = 0; kw < 3; ++kw)
        {
          float w[3];
          for (int kh = 0; kh < 3; ++kh) w[kh] = half_bits_to_float(w_oihw[(((size_t)o * I + i) * 3 + kh) * 3 + kw]);
          const float st[4] = {w[0] + w[1], w[1] + w[2], w[2], w[0]};
          for (int j = 0; j < 4; ++j)
            f[((size_t)(kw * 4 + j) * pl.CoutAlloc + o) * C1 + i] = float_to_half_bits(st[j]);
        }
  }
  return 0;
}

int oidnb200_conv_pack_bias(const oidnb200_conv* conv, const uint16_t* b_x, int O, void* dst_bias)
{
  const ConvPlan& pl = conv->plan;
  if (!b_x || !dst_bias || O > pl.desc.Cout || O <= 0)
  {
    set_error("conv_pack_bias: bad arguments");
    return OIDNB200_ERR_INVALID;
  }
  float* dst = static_cast<float*>(dst_bias);
  for (int o = 0; o < pl.CoutAlloc; ++o) dst[o] = o < O ? half_bits_to_float(b_x[o]) : 0.f;
  conv->host_bias.assign(dst, dst + pl.CoutAlloc);   // bind() copies it into the kernel parameters
  return 0;
}

int oidnb200_conv_bind(oidnb200_conv* conv, const void* src1, const void* src2, const void* weights,
                       const void* bias, void* dst)
{
  const int rc = plan_bind(conv->plan, src1, src2, weights, bias, dst);
  auto embed = [&](ConvPlan& pl) {
    pl.kp.bias_in_params = 0;
    if (!getenv("OIDN_B200_SMEM_BIAS") && (int)conv->host_bias.size() == pl.CoutAlloc && pl.CoutAlloc <= kMaxBiasParams)
    {
      memcpy(pl.kp.bias_c, conv->host_bias.data(), conv->host_bias.size() * sizeof(float));
      pl.kp.bias_in_params = 1;
    }
  };
  if (rc == 0) embed(conv->plan);
  if (rc == 0 && conv->fused_plan)
  {
    const int rc2 = plan_bind(*conv->fused_plan, src1, src2, weights, bias, dst);
    if (rc2 == 0) embed(*conv->fused_plan);
    return rc2;
  }
  return rc;
}

int oidnb200_conv_set_output_process(oidnb200_conv* conv, const oidnb200_tile* tile, const oidnb200_transfer* tf,
                                     int hdr, int snorm, const oidnb200_image* dst)
{
  ConvPlan& pl = conv->plan;
  FusedOutput& fo = pl.kp.fo;
  if (!dst)
  {
    fo.enabled = 0;
    if (conv->fused_plan) conv->fused_plan->kp.fo.enabled = 0;
    return 0;
  }
  const oidnb200_conv_desc& d = pl.desc;
  Transfer t;
  if (!tile || !make_transfer(tf, t))
  {
    set_error("conv_set_output_process: bad arguments");
    return OIDNB200_ERR_INVALID;
  }
  if (d.Cout != 16 || d.post_op != POST_NONE || pl.kp.ngroups != 1)
  {
    set_error("conv_set_output_process: only the last convolution (16 padded output channels, no post-op) can be fused");
    return OIDNB200_ERR_UNSUPPORTED;
  }
  if (dst->format != OIDNB200_FORMAT_FLOAT3 || dst->pixel_stride != 12 || dst->row_stride % 4 != 0 ||
      reinterpret_cast<uintptr_t>(dst->ptr) % 4 != 0 || !dst->ptr)
  {
    set_error("conv_set_output_process: the fused path writes packed fp32 RGB images only");
    return OIDNB200_ERR_UNSUPPORTED;
  }
  if (tile->H < 0 || tile->W < 0 || tile->hSrcBegin < 0 || tile->wSrcBegin < 0 || tile->hSrcBegin + tile->H > d.H ||
      tile->wSrcBegin + tile->W > d.W || tile->hDstBegin < 0 || tile->wDstBegin < 0 ||
      tile->hDstBegin + tile->H > dst->H || tile->wDstBegin + tile->W > dst->W)
  {
    set_error("conv_set_output_process: tile outside the tensor or the image");
    return OIDNB200_ERR_INVALID;
  }
  fo.enabled = 1;
  fo.ptr = static_cast<unsigned char*>(dst->ptr);
  fo.rs = (long long)dst->row_stride;
  fo.hSrc = tile->hSrcBegin; fo.wSrc = tile->wSrcBegin; fo.hDst = tile->hDstBegin; fo.wDst = tile->wDstBegin;
  fo.H = tile->H; fo.W = tile->W;
  fo.tf_type = t.type; fo.norm = t.norm; fo.rcp_norm = t.rcp_norm;
  fo.input_scale = t.input_scale; fo.input_scale_ptr = t.input_scale_ptr;
  fo.hdr = hdr; fo.snorm = snorm;
  if (conv->fused_plan) conv->fused_plan->kp.fo = fo;
  return 0;
}

int oidnb200_conv_launch(const oidnb200_conv* conv, oidnb200_stream stream)
{
  const ConvPlan& pl = (conv->plan.kp.fo.enabled && conv->fused_plan) ? *conv->fused_plan : conv->plan;
  if (!pl.bound)
  {
    set_error("conv_launch: op not bound");
    return OIDNB200_ERR_INVALID;
  }
  if (pl.desc.post_op == POST_UPSAMPLE)
  {
    set_error("conv_launch: PostOp::Upsample is fused into the consumer (src1_upsampled); "
              "the tensor-core kernel never materialises an upsampled tensor");
    return OIDNB200_ERR_UNSUPPORTED;
  }
  cudaError_t e = conv3x3_tc_launch(pl.kp, pl.grid, pl.smem, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess)
  {
    set_error(std::string("conv_launch: ") + cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

int oidnb200_conv_launch_simt(const oidnb200_conv* conv, void* scratch, oidnb200_stream stream)
{
  const ConvPlan& pl = conv->plan;
  const oidnb200_conv_desc& d = pl.desc;
  if (!pl.bound || !scratch)
  {
    set_error("conv_launch_simt: op not bound or no scratch");
    return OIDNB200_ERR_INVALID;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long n = (long)d.H * d.W * d.Cout;
  conv3x3_simt_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
    static_cast<const __half*>(pl.src1), static_cast<const __half*>(pl.src2), d.C1, d.C2,
    d.src1_upsampled, d.H, d.W, static_cast<const __half*>(pl.weights), pl.CinTot, pl.CoutAlloc,
    pl.kp.bias, d.Cout, d.relu, static_cast<__half*>(scratch));
  long m = n;
  if (d.post_op == POST_POOL) m = n / 4;
  if (d.post_op == POST_UPSAMPLE) m = n * 4;
  post_simt_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(static_cast<const __half*>(scratch), d.H,
                                                               d.W, d.Cout, d.post_op,
                                                               static_cast<__half*>(pl.dst));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
  {
    set_error(std::string("conv_launch_simt: ") + cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

int oidnb200_conv_set_trace(oidnb200_conv* conv, void* trace_counters)
{
  conv->plan.kp.trace = static_cast<unsigned long long*>(trace_counters);
  return 0;
}

int oidnb200_conv_set_stamps(oidnb200_conv* conv, void* stamps)
{
  conv->plan.kp.stamps = static_cast<unsigned long long*>(stamps);
  if (conv->fused_plan) conv->fused_plan->kp.stamps = static_cast<unsigned long long*>(stamps);
  return 0;
}

int oidnb200_conv_get_info(const oidnb200_conv* conv, oidnb200_conv_info* info)
{
  const ConvPlan& pl = conv->plan;
  info->grid = pl.grid;
  info->smem_bytes = (int)pl.smem;
  info->ngroups = pl.kp.ngroups;
  info->cout_group = pl.kp.CoutG;
  info->nchunks = pl.kp.nchunks;
  info->nstages = pl.kp.nstages;
  info->ring_slots = pl.kp.R;
  info->rows_per_item = pl.kp.RC;
  info->nstrips = pl.kp.nstrips;
  info->nrowchunks = pl.kp.nrowchunks;
  info->nstreams = pl.kp.nstreams;
  info->out_nbuf = pl.kp.out_nbuf;
  return 0;
}

} // extern "C"
