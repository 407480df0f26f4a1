// Shared host/device definitions for the 3x3 convolution kernels.
//
// Tensor layout in HBM (all activations): NHWC, fp16, N=1, channels padded to a multiple of 16
// (one tcgen05 K step). The reference CUDA device uses the same "hwc" layout with blockC=8
// (devices/cuda/cuda_device.cpp:215-219); 16 is the fp16 UMMA K granularity.
//
// Packed weight layout in HBM: [kw][kh][CoutAlloc][CinTot] fp16, CinTot = C1p + C2p (the padded
// channels of src1 followed by the padded channels of src2: the in-place concat of
// core/graph.cpp:205-208), CoutAlloc = ngroups*CoutG >= CoutPad, zero padded.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace oidnb200 {

constexpr int kMaxChunks   = 8;     // K chunks (<=64 channels each) over both sources
constexpr int kMaxOutChunks = 3;    // output-channel pieces (64/32/16) of one CoutG group
constexpr int kMaxStages   = 24;    // A-operand pipeline depth (per stream)
constexpr int kConvThreads = 384;   // 2 x (TMA warp + MMA warp) + 2 x 4 epilogue warps
constexpr int kConvThreads4 = 768;  // four-stream variant (CoutG <= 32): 4 x (TMA + MMA warp) + 4 x 4 epilogue warps
constexpr int kMaxStreams  = 4;
constexpr int kStripW      = 128;   // output pixels per MMA tile (UMMA M)
constexpr int kMaxStageBytes = 17408; // one A stage: up to 132 px x 128 B, 1024-aligned (64-channel chunk)
constexpr int kSmemHeader  = 4096;  // barriers + TMEM pointer (3 KB) + bias (512 B)
constexpr int kTmemCols    = 512;
constexpr int kMaxSlots    = 16;
constexpr int kSmemBudget  = 232448; // 227 KB opt-in dynamic shared memory per CTA
constexpr int kMaxBiasParams = 256; // output channels whose bias travels in the kernel parameters

// Output-channel pieces of a group of nb*16 channels: 64-channel pieces, then 32, then 16 (the TMA
// store box / swizzle widths). Shared by the planner (tensor maps) and the kernel (staging layout).
__host__ __device__ constexpr int out_piece_count(int nb) { return nb / 4 + ((nb >> 1) & 1) + (nb & 1); }
__host__ __device__ constexpr int out_piece_cc(int nb, int i)
{
  return i < nb / 4 ? 64 : ((i == nb / 4 && (nb & 2)) ? 32 : 16);
}
__host__ __device__ constexpr int out_piece_c0(int nb, int i)
{
  return i <= nb / 4 ? i * 64 : (nb / 4) * 64 + 32;
}
// byte offset of piece i inside one staging slice of `rows` pixels (every piece 1024-aligned)
__host__ __device__ constexpr uint32_t out_piece_off(int nb, int i, int rows)
{
  uint32_t off = 0;
  for (int k = 0; k < i; ++k) off += ((uint32_t)(rows * out_piece_cc(nb, k) * 2) + 1023u) / 1024u * 1024u;
  return off;
}

enum PostOp : int { POST_NONE = 0, POST_POOL = 1, POST_UPSAMPLE = 2 /* SIMT witness only */ };

// Output process fused into the epilogue of the network's last convolution (16 padded output
// channels, 3 used): instead of storing the tensor, every pixel of the tile's output rectangle goes
// through the output-process math (devices/gpu/gpu_output_process.h:35-73) and is written to the
// packed fp32 RGB output image. Removes the tensor write, its re-read and one launch.
struct FusedOutput
{
  int      enabled;
  unsigned char* ptr;               // output image: 3 x fp32 per pixel, pixel stride 12 bytes
  long long rs;                     // row stride in bytes
  int      hSrc, wSrc, hDst, wDst, H, W; // core/tile.h: rectangle in the tensor -> origin in the image
  int      tf_type;                 // OIDNB200_TF_*
  float    norm, rcp_norm;          // PU / Log normalisation
  float    input_scale;             // used when input_scale_ptr == nullptr
  const float* input_scale_ptr;     // autoexposure result (device memory)
  int      hdr, snorm;
};

struct ConvKernelParams
{
  CUtensorMap amap[kMaxChunks]; // activations, one per K chunk (3D, or 4D with a stride-0 dup axis for a virtually upsampled source)
  CUtensorMap wmap[kMaxChunks]; // packed weights, one per K chunk (4D)
  CUtensorMap omap[kMaxOutChunks]; // destination, one per output-channel piece of the group (3D)
  int      nchunks;
  int      chunk_c0[kMaxChunks];    // first channel of the chunk inside its source tensor
  int      chunk_wc0[kMaxChunks];   // first channel of the chunk on the packed-weight Cin axis
  int      chunk_cc[kMaxChunks];    // channels in the chunk: 16, 32 or 64 (row bytes 32/64/128)
  int      chunk_up[kMaxChunks];    // 1: source stored at half resolution, read through the 4D dup map
  uint32_t chunk_boff[kMaxChunks];  // byte offset of the chunk's kw=0 weight block in the B region
  uint32_t chunk_bblk[kMaxChunks];  // bytes of one kw block (3*CoutG rows)
  // the MMA issuer's per-chunk constants (descriptor arithmetic is in 16-byte units)
  uint32_t chunk_nk[kMaxChunks];    // 16-channel k-steps in the chunk (1, 2 or 4)
  uint32_t chunk_hi[kMaxChunks];    // high word of the chunk's UMMA shared-memory descriptors
  uint32_t chunk_a16[kMaxChunks];   // start of the strip inside a stage: one pixel in for an upsampled source
  uint32_t chunk_b16[kMaxChunks];   // chunk_boff / 16
  uint32_t chunk_bblk16[kMaxChunks]; // chunk_bblk / 16
  int      H, W;                    // conv resolution (= resolution of the unpooled output)
  int      CoutG, ngroups, CoutPad; // output channels per CTA group / groups / padded total
  int      nstreams;                // 1, 2 or 4 independent row streams per CTA
  int      R;                       // TMEM accumulator ring slots per stream (nstreams*R*CoutG <= 512)
  int      RC, nstrips, nrowchunks; // rows per work item, strips across W, row chunks down H
  int      nstages;                 // A pipeline stages per stream
  int      up_fold;                 // 1: a virtually upsampled src1 is convolved at ITS OWN row resolution. Upsampled rows
                                    // 2Y and 2Y+1 are the same low-res row Y, so row Y is staged ONCE (its own ring of
                                    // nstages_u stages) and used by two consecutive virtual rows with vertically
                                    // pre-summed weights (conv_plan.cu packs them behind the regular ones): at the even
                                    // row r = 2Y the usual stack (output rows r+1, r, r-1) takes E = [w0+w1; w1+w2; w2],
                                    // at the odd row r = 2Y+1 only the fresh row r+1 takes O = [w0]: N = 3+1 row blocks
                                    // per low-res row instead of 3+3 (two instead of three row taps per output row)
  int      nstages_u;               // folded: stages [0, nstages_u) of the ring hold src1 chunks, the rest src2 chunks
  int      prefetch_rows;           // > 0: the TMA producer prefetches input rows this far ahead into L2
  uint32_t stage_bytes;             // bytes of the widest A stage: 132 px x the widest K chunk, 1024-aligned
  uint32_t ring_bytes;              // bytes of one stream's A ring
  uint32_t stage_off[kMaxStages];   // byte offset of stage s inside the stream's ring. When nstages is a
                                    // multiple of nchunks, stage s always holds chunk s % nchunks and is
                                    // sized for it (narrow chunks take less room: deeper ring); otherwise
                                    // every stage has stage_bytes
  uint32_t w_bytes;                 // total weight bytes TMA-loaded per CTA
  uint32_t b_bytes;                 // size of the resident weight region (1024-aligned blocks)
  int      nout;                    // output pieces of the group
  int      out_c0[kMaxOutChunks];   // first channel of the piece inside the group
  int      out_cc[kMaxOutChunks];   // channels in the piece (64/32/16)
  uint32_t out_off[kMaxOutChunks];  // byte offset of the piece inside one staging slice
  uint32_t out_buf_bytes;           // bytes of one staging slice (one warp, one row)
  int      out_nbuf;                // 1 or 2 staging slices per epilogue warp (8 warps; 16 with four streams)
  int      direct_store;            // 1: the epilogue stores its pixels straight from registers to global
                                    // memory (32-byte stores), no staging slices and no TMA store: the
                                    // shared memory goes to A stages instead
  void*    out_ptr;                 // destination tensor [Hd][Wd][CoutPad] fp16 (direct_store)
  int      out_W;                   // Wd
  int      relu, post_op;
  const float* bias;                // fp32 [CoutAlloc]
  // The same values inside the kernel parameters (constant bank) when the op knows them on the host
  // (oidnb200_conv_pack_bias) and CoutAlloc <= kMaxBiasParams: the epilogue then reads its bias through the
  // constant cache instead of shared memory, whose data pipe the tensor core's operand reads already fill on
  // these layers (ncu: l1tex__data_pipe_{tc,lsu}_wavefronts_mem_shared).
  int      bias_in_params;
  float    bias_c[kMaxBiasParams];
  unsigned long long* trace;        // [12 warps][16 tags] wait-cycle counters (OIDN_B200_TRACE builds), else null
  unsigned long long* stamps;       // non-null: {min over CTAs of "past griddepcontrol.wait", max over CTAs of the exit} in %globaltimer
                                    // nanoseconds (two atomics per CTA): when the grid really ran inside a frame whose
                                    // launches overlap through programmatic dependent launch (bench.py's conv time)
  FusedOutput fo;                   // fo.enabled: the epilogue writes the output image instead of the tensor
};

// ---- two chained convolutions in one kernel (conv_pair_tc.cu) ------------------------------------------------------
constexpr int kPairStrip     = 126;  // output pixels of conv B per work-item row (conv A computes 128: one halo pixel each side)
constexpr int kPairThreads   = 704;  // 2 x (TMA + A-MMA warp) + 2 x 4 A-epilogue warps + 2 x 4 B-epilogue warps + 2 B-MMA warps
constexpr int kPairMaxStages = 16;   // input ring stages per stream
constexpr int kPairMaxMid    = 8;    // mid ring stages per stream

struct PairKernelParams
{
  CUtensorMap amap;                 // conv A's source (3D, box {ccA, 130, 1})
  CUtensorMap wmapA, wmapB;         // packed weights of A and B (4D, box {cc, Cout, 3, 1})
  int      H, W;
  int      ccA;                     // A's input channels (padded; one K chunk: 16, 32 or 64)
  int      CA;                      // A's output = B's input channels (32 or 64)
  int      CB;                      // B's padded output channels (16 .. 64)
  int      poolB, reluA, reluB;
  int      nstreams;                // 1 or 2
  int      RA, RB;                  // TMEM accumulator ring slots of A / B per stream
  int      NA, NM;                  // input ring / mid ring stages per stream
  int      RC, nstrips, nrowchunks; // rows per work item, strips of kPairStrip pixels, row chunks
  uint32_t a_stage_bytes, mid_stage_bytes;
  uint32_t wA_blk, wB_blk;          // bytes of one kw block of A's / B's resident weights
  uint32_t wB_off;                  // B's weights inside the weight region
  uint32_t w_bytes;                 // bytes TMA-loaded per CTA (both convs)
  uint32_t w_bytes_smem;            // size of the weight region (1024-aligned)
  uint32_t hiA, hiB;                // high words of the UMMA descriptors (rows of ccA*2 / CA*2 bytes)
  const float* biasA;
  const float* biasB;
  int      bias_in_params;          // as ConvKernelParams: biases of A and B in the constant bank
  float    biasA_c[64], biasB_c[64];
  void*    out_ptr;                 // B's destination tensor [Hd][Wd][CoutPadB] fp16
  int      out_W, CoutPadB;
  unsigned long long* stamps;       // as ConvKernelParams::stamps
  unsigned long long* trace;        // [22 warps][16 tags] wait-cycle counters (OIDN_B200_TRACE builds), else null
  FusedOutput fo;                   // fo.enabled: B's epilogue writes the output image instead of the tensor
  int tapB;                         // tap-packed B (3 output channels, fused output): wmapB = [kh][kw * 3 + c][CA], one unshifted
                                    // view per K step instead of three; the epilogue adds the horizontal taps across lanes
};

} // namespace oidnb200
