/* oidn_b200 kernel-level C ABI.
 *
 * This is the boundary a device module in the reference tree binds to: every entry point below is
 * what one `Engine::new<Op>()` product (core/engine.h:67-75) needs in order to run on a B200, with
 * plain pointers and sizes only (no C++ or torch types). Each function cites the reference
 * interface it stands in for. All functions return 0 on success or a negative oidnb200 error
 * code / positive cudaError_t; oidnb200_last_error() returns a message for the calling thread.
 *
 * Tensors: NHWC fp16, N=1, channels padded to a multiple of 16 ("hwc", tensorBlockC=16 in the
 * reference's vocabulary, core/tensor_layout.h:11-33). Images: the user's strided 1-3 channel
 * fp32/fp16 pixel buffers exactly as passed to oidnSetSharedFilterImage (core/image.h:14-120).
 */
#ifndef OIDN_B200_KERNELS_H
#define OIDN_B200_KERNELS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OIDNB200_API __attribute__((visibility("default")))

typedef void* oidnb200_stream; /* cudaStream_t */

enum
{
  OIDNB200_OK = 0,
  OIDNB200_ERR_INVALID = -1,     /* bad descriptor / argument */
  OIDNB200_ERR_UNSUPPORTED = -2, /* shape cannot be mapped (e.g. weights do not fit in smem) */
  OIDNB200_ERR_DRIVER = -3       /* driver entry point missing / tensor-map encode failed */
};

OIDNB200_API const char* oidnb200_last_error(void);

/* Number of CUDA devices with compute capability 10.x; 0 when no driver/GPU is present.
 * Stands in for CUDADevice::getPhysicalDevices (devices/cuda/cuda_device.cpp:60-101). */
OIDNB200_API int oidnb200_device_count(void);

/* ------------------------------------------------------------------------------------------
 * Conv / ConcatConv (+ fused Pool, + upsampled source)   -- core/conv.h:26-61,
 * core/concat_conv.h, core/pool.h, core/upsample.h; replaces devices/cuda/cutlass_conv.h
 * ------------------------------------------------------------------------------------------ */
typedef struct oidnb200_conv_desc
{
  int H, W;          /* resolution the convolution runs at (its un-pooled output resolution) */
  int C1;            /* padded channels of src1 (multiple of 16) */
  int C2;            /* padded channels of src2, 0 if the op is a plain Conv (ConcatConv otherwise) */
  int Cout;          /* padded output channels (multiple of 16) */
  int relu;          /* Activation::ReLU (core/conv.h:12-16) */
  int post_op;       /* 0 none, 1 PostOp::Pool (dst is H/2 x W/2), 2 PostOp::Upsample (dst is 2H x 2W) */
  int src1_upsampled;/* 1: src1 is stored at H/2 x W/2 and nearest-upsampled by the loader (fused Upsample) */
  int shift_mode;    /* 0 (product). 1,2 are hardware-probe variants used only by tools/probe_conv */
} oidnb200_conv_desc;

typedef struct oidnb200_conv oidnb200_conv;

OIDNB200_API int oidnb200_conv_create(const oidnb200_conv_desc* desc, oidnb200_conv** out);
OIDNB200_API void oidnb200_conv_destroy(oidnb200_conv* conv);

/* Bytes of the packed (device-layout) weight and bias buffers. The weight buffer holds the regular layout
 * [kw][kh][Cout][Cin] and, for the convs that use them, derived copies behind it: the vertically pre-summed rows of an
 * upsampled source (row folding) and the tap-packed copy [kh][kw * 3 + c][Cin] of a 3-channel last conv (fused pairs). */
OIDNB200_API size_t oidnb200_conv_weight_bytes(const oidnb200_conv* conv);
OIDNB200_API size_t oidnb200_conv_bias_bytes(const oidnb200_conv* conv);

/* Host-side weight/bias reorder: TZA "oihw" fp16 [O][I][3][3] + "x" fp16 bias [O]
 * -> packed device layout (zero padded). I = I1 + I2 logical input channels (ConcatConv splits the
 * I axis at I1: core/graph.cpp:205-208). Stands in for reorderWeight/reorderBias
 * (core/tensor_reorder.cpp:8-98). dst buffers are host memory of the sizes returned above. */
OIDNB200_API int oidnb200_conv_pack_weights(const oidnb200_conv* conv, const uint16_t* w_oihw,
                                            int O, int I1, int I2, void* dst_weights);
OIDNB200_API int oidnb200_conv_pack_bias(const oidnb200_conv* conv, const uint16_t* b_x, int O,
                                         void* dst_bias);

/* Bind device pointers (encodes the TMA tensor maps). src2 may be NULL when C2 == 0. */
OIDNB200_API int oidnb200_conv_bind(oidnb200_conv* conv, const void* src1, const void* src2,
                                    const void* weights, const void* bias, void* dst);

/* Conv::submitKernels (core/op.h:44-51). */
OIDNB200_API int oidnb200_conv_launch(const oidnb200_conv* conv, oidnb200_stream stream);

/* Debug/reference SIMT implementation of the same op on the same buffers (tests only). */
OIDNB200_API int oidnb200_conv_launch_simt(const oidnb200_conv* conv, void* scratch_fp16,
                                           oidnb200_stream stream);

/* Introspection for tests / DESIGN.md numbers. */
typedef struct oidnb200_conv_info
{
  int grid, smem_bytes, ngroups, cout_group, nchunks, nstages, ring_slots, rows_per_item;
  int nstrips, nrowchunks, nstreams, out_nbuf;
} oidnb200_conv_info;
OIDNB200_API int oidnb200_conv_get_info(const oidnb200_conv* conv, oidnb200_conv_info* info);
/* Profiling builds (-DOIDN_B200_TRACE) only: device array of 12 x 16 uint64 that receives, per warp
 * role, the cycles spent blocked at each barrier (tools/probe_conv --trace). No effect otherwise. */
OIDNB200_API int oidnb200_conv_set_trace(oidnb200_conv* conv, void* trace_counters);

/* Two chained convolutions B(A(x)) as ONE launch: the tensor between them stays in shared memory (no HBM write /
 * re-read). For the UNet's full-resolution pairs enc_conv0 -> enc_conv1 (+pool) and dec_conv1b -> dec_conv0 (+output
 * process) (core/unet_filter.cpp:468-531). `a` and `b` are ordinary conv ops (created, weights packed, bound as
 * usual; they must outlive the pair): the pair takes A's source / weights / bias and B's weights / bias / destination /
 * fused output process from them. create returns OIDNB200_ERR_UNSUPPORTED when the shapes are not covered (plain
 * 3x3 convs, A: <= 64 input and 32 or 64 output channels, B: <= 64 output channels, optional pool) -- the caller then
 * simply launches the two convs. The result is bit-identical to the two launches. */
typedef struct oidnb200_conv_pair oidnb200_conv_pair;
OIDNB200_API int oidnb200_conv_pair_create(const oidnb200_conv* a, const oidnb200_conv* b, oidnb200_conv_pair** out);
OIDNB200_API void oidnb200_conv_pair_destroy(oidnb200_conv_pair* pair);
OIDNB200_API int oidnb200_conv_pair_bind(oidnb200_conv_pair* pair);   /* after (re)binding either conv */
OIDNB200_API int oidnb200_conv_pair_launch(oidnb200_conv_pair* pair, oidnb200_stream stream);
OIDNB200_API int oidnb200_conv_pair_get_info(const oidnb200_conv_pair* pair, oidnb200_conv_info* info);

/* In-frame timing: `stamps` = device array of 2 x uint64 {UINT64_MAX, 0} (or NULL to switch off). Every launch
 * then records, in %globaltimer nanoseconds, the earliest moment one of its CTAs got past the wait for the
 * previous grid (griddepcontrol.wait: its first activation load) and the latest CTA exit -- when the grid really ran
 * inside a frame whose launches overlap (programmatic dependent launch), which CUDA events around a launch serialise. */
OIDNB200_API int oidnb200_conv_set_stamps(oidnb200_conv* conv, void* stamps);

/* ------------------------------------------------------------------------------------------
 * Images, tiles, transfer functions  -- core/image.h:14-120, core/tile.h, core/color.h:10-166
 * ------------------------------------------------------------------------------------------ */
/* Pixel formats use the public API's numbering (include/OpenImageDenoise/oidn.h:239-254). */
enum
{
  OIDNB200_FORMAT_UNDEFINED = 0,
  OIDNB200_FORMAT_FLOAT = 1, OIDNB200_FORMAT_FLOAT2 = 2, OIDNB200_FORMAT_FLOAT3 = 3,
  OIDNB200_FORMAT_HALF = 257, OIDNB200_FORMAT_HALF2 = 258, OIDNB200_FORMAT_HALF3 = 259
};

/* A strided 1-3 channel image exactly as given to oidnSetSharedFilterImage: device, managed or
 * pinned-host memory (anything the GPU can dereference). ptr == NULL means "image not set". */
typedef struct oidnb200_image
{
  void* ptr;
  int format;          /* OIDNB200_FORMAT_* */
  int W, H;
  size_t pixel_stride; /* bytes */
  size_t row_stride;   /* bytes */
} oidnb200_image;

/* core/tile.h: source/destination origin and size of the rectangle an op works on. */
typedef struct oidnb200_tile
{
  int hSrcBegin, wSrcBegin, hDstBegin, wDstBegin, H, W;
} oidnb200_tile;

enum { OIDNB200_TF_LINEAR = 0, OIDNB200_TF_SRGB = 1, OIDNB200_TF_PU = 2, OIDNB200_TF_LOG = 3 };

/* TransferFunction (core/color.h:10-166): type + input scale, either a value or a device pointer
 * to the autoexposure result (core/color.h:110-123). The PU/Log normalisation 1/forward(65504)
 * (core/color.cpp:9-16) is derived inside. */
typedef struct oidnb200_transfer
{
  int type;                     /* OIDNB200_TF_* */
  float input_scale;            /* used when input_scale_ptr == NULL */
  const float* input_scale_ptr; /* device pointer or NULL */
} oidnb200_transfer;

/* InputProcess::submitKernels (core/input_process.h, devices/gpu/gpu_input_process.h:82-176):
 * fills the whole [TH][TW][C] fp16 tile buffer (C = 16: 3/6/9 used channels, rest zero; zero
 * outside the tile footprint). `color` is the main input (color, or albedo/normal when filtering
 * an auxiliary image alone); albedo/normal may have ptr == NULL. */
OIDNB200_API int oidnb200_input_process_launch(const oidnb200_image* color, const oidnb200_image* albedo,
                                               const oidnb200_image* normal, const oidnb200_tile* tile,
                                               const oidnb200_transfer* tf, int hdr, int snorm,
                                               void* dst, int TH, int TW, int C, oidnb200_stream stream);

/* OutputProcess::submitKernels (core/output_process.h, devices/gpu/gpu_output_process.h:35-73):
 * reads channels 0..2 of the [TH][TW][C] fp16 tensor at tile.{h,w}SrcBegin and writes the
 * tile.H x tile.W rectangle at tile.{h,w}DstBegin of the output image. */
OIDNB200_API int oidnb200_output_process_launch(const void* src, int TH, int TW, int C,
                                                const oidnb200_tile* tile, const oidnb200_transfer* tf,
                                                int hdr, int snorm, const oidnb200_image* dst,
                                                oidnb200_stream stream);

/* OutputProcess fused into the network's last convolution (the reference runs the two ops back to
 * back: core/unet_filter.cpp:495-496 + :228-231). After this call oidnb200_conv_launch applies the
 * output process (devices/gpu/gpu_output_process.h:35-73; same tile / transfer / hdr / snorm meaning as
 * oidnb200_output_process_launch above) to channels 0..2 of its result and writes the tile rectangle
 * straight into `dst`; the conv's tensor `dst` is then NOT written. The result is bit-identical to
 * conv + oidnb200_output_process_launch. Supported for a conv with 16 padded output channels and no
 * post-op, and a FLOAT3 image with pixel stride 12: otherwise OIDNB200_ERR_UNSUPPORTED is returned and
 * the conv stays unfused (the caller then launches the separate pass). dst == NULL removes the
 * fusion. Cheap: call per tile per frame like OutputProcess::setTile/setDst (no tensor-map encode). */
OIDNB200_API int oidnb200_conv_set_output_process(oidnb200_conv* conv, const oidnb200_tile* tile,
                                                  const oidnb200_transfer* tf, int hdr, int snorm,
                                                  const oidnb200_image* dst);

/* Autoexposure (core/autoexposure.h:14-53, devices/gpu/gpu_autoexposure.h:13-164; the reference
 * uses three launches): result = 0.18 / exp2(mean log2 of the bin luminances > 1e-8) (1 if none)
 * stored to *dst (device float). Bins are <= 16x16 pixels, bin i covers rows [i*H/nbh, (i+1)*H/nbh).
 *   oidnb200_autoexposure_launch: the whole image (bins + reduce); scratch = device memory of
 *     oidnb200_autoexposure_scratch_bytes(H, W) (the bin array).
 *   oidnb200_autoexposure_bins_launch: log2(mean luminance) of the bins [bin_h0,bin_h1) x
 *     [bin_w0,bin_w1) of src's bin grid into bins[nbh*nbw] (-inf = bin not counted); only the pixels
 *     of those bins are read, so a GPU that holds one tile of the frame computes the bins of its tile.
 *   oidnb200_autoexposure_reduce_launch: folds a complete bin array in a fixed order -> *dst. The
 *     result depends on the array only (multi-GPU: sum the per-GPU arrays, zeros elsewhere, first). */
OIDNB200_API void oidnb200_autoexposure_bin_grid(int H, int W, int* num_bins_h, int* num_bins_w);
OIDNB200_API size_t oidnb200_autoexposure_scratch_bytes(int H, int W);
OIDNB200_API int oidnb200_autoexposure_launch(const oidnb200_image* src, void* scratch, float* dst,
                                              oidnb200_stream stream);
OIDNB200_API int oidnb200_autoexposure_bins_launch(const oidnb200_image* src, int bin_h0, int bin_h1, int bin_w0,
                                                   int bin_w1, float* bins, oidnb200_stream stream);
OIDNB200_API int oidnb200_autoexposure_reduce_launch(const float* bins, int num_bins, float* dst,
                                                     oidnb200_stream stream);

/* Peer flags: frame-level hand-shake between the GPUs of a tile-sharded frame without a collective (the role
 * Device::submitBarrier plays between the engines of one process, core/unet_filter.cpp:178,243, when the engines are
 * in different processes). A flag array holds one uint32 per rank, in device memory every rank can reach (CUDA IPC /
 * peer access). signal: stores `value` (a growing frame sequence number) into the n given slots -- this rank's slot
 * of every rank's array -- after everything enqueued before it on the stream has completed. wait: returns (in stream
 * order) once all n slots of the LOCAL array hold at least `value`; traps after timeout_s seconds (0 = 10 s) instead
 * of hanging. Both are one 32-thread block without shared memory: they run next to a persistent conv CTA. */
OIDNB200_API int oidnb200_flag_signal_launch(void* const* slots, int n, unsigned int value, oidnb200_stream stream);
OIDNB200_API int oidnb200_flag_wait_launch(const void* flags, int n, unsigned int value, double timeout_s,
                                           oidnb200_stream stream);

/* ImageCopy (core/image_copy.h, devices/gpu/gpu_image_copy.h:15-27): same format and size. */
OIDNB200_API int oidnb200_image_copy_launch(const oidnb200_image* src, const oidnb200_image* dst,
                                            oidnb200_stream stream);

/* Stand-alone Pool / Upsample (core/pool.h, core/upsample.h; devices/gpu/gpu_pool.h:33-52,
 * gpu_upsample.h:33-52) on [H][W][C] fp16 tensors, C a multiple of 8. The UNet graph never
 * launches these (both are fused into the convolutions); they complete the Engine op surface. */
OIDNB200_API int oidnb200_pool_launch(const void* src, int H, int W, int C, void* dst, oidnb200_stream stream);
OIDNB200_API int oidnb200_upsample_launch(const void* src, int H, int W, int C, void* dst, oidnb200_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* OIDN_B200_KERNELS_H */
