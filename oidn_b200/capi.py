"""ctypes bindings of liboidn_b200.so (both C ABIs). Fails loudly when the library is missing:
there is no fallback implementation."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboidn_b200.so")

FORMAT_UNDEFINED, FORMAT_FLOAT, FORMAT_FLOAT2, FORMAT_FLOAT3 = 0, 1, 2, 3
FORMAT_HALF, FORMAT_HALF2, FORMAT_HALF3 = 257, 258, 259
QUALITY_DEFAULT, QUALITY_FAST, QUALITY_BALANCED, QUALITY_HIGH = 0, 4, 5, 6
STORAGE_UNDEFINED, STORAGE_HOST, STORAGE_DEVICE, STORAGE_MANAGED = 0, 1, 2, 3
(ERROR_NONE, ERROR_UNKNOWN, ERROR_INVALID_ARGUMENT, ERROR_INVALID_OPERATION, ERROR_OUT_OF_MEMORY,
 ERROR_UNSUPPORTED_HARDWARE, ERROR_CANCELLED) = range(7)
EXTERNAL_MEMORY_OPAQUE_FD, EXTERNAL_MEMORY_DMA_BUF = 1, 2
TF_LINEAR, TF_SRGB, TF_PU, TF_LOG = 0, 1, 2, 3


class ConvDesc(C.Structure):
  _fields_ = [(n, C.c_int) for n in ("H", "W", "C1", "C2", "Cout", "relu", "post_op", "src1_upsampled", "shift_mode")]


class ConvInfo(C.Structure):
  _fields_ = [(n, C.c_int) for n in ("grid", "smem_bytes", "ngroups", "cout_group", "nchunks", "nstages", "ring_slots",
                                     "rows_per_item", "nstrips", "nrowchunks", "nstreams", "out_nbuf")]


class Image(C.Structure):
  _fields_ = [("ptr", C.c_void_p), ("format", C.c_int), ("W", C.c_int), ("H", C.c_int),
              ("pixel_stride", C.c_size_t), ("row_stride", C.c_size_t)]


class Tile(C.Structure):
  _fields_ = [(n, C.c_int) for n in ("hSrcBegin", "wSrcBegin", "hDstBegin", "wDstBegin", "H", "W")]


class Transfer(C.Structure):
  _fields_ = [("type", C.c_int), ("input_scale", C.c_float), ("input_scale_ptr", C.c_void_p)]


class FilterInfo(C.Structure):
  _fields_ = [(n, C.c_int) for n in ("tileH", "tileW", "tileCountH", "tileCountW", "tileOverlap", "tileAlignment",
                                     "largeModel", "numOps", "staged")] + [("memoryBytes", C.c_size_t)]


class OpTime(C.Structure):
  _fields_ = [("name", C.c_char * 32), ("kind", C.c_int), ("launches", C.c_int), ("ms", C.c_double)]


class TilePlan(C.Structure):
  _fields_ = [(n, C.c_int) for n in ("H", "W", "tileH", "tileW", "tilePadH", "tilePadW", "tileCountH", "tileCountW",
                                     "tileAlignment", "tileOverlap")]


PROGRESS_FUNC = C.CFUNCTYPE(C.c_bool, C.c_void_p, C.c_double)

# name -> (restype, argtypes); every symbol include/*.h declares
KERNEL_ABI = {
  "oidnb200_last_error": (C.c_char_p, []),
  "oidnb200_device_count": (C.c_int, []),
  "oidnb200_conv_create": (C.c_int, [C.POINTER(ConvDesc), C.POINTER(C.c_void_p)]),
  "oidnb200_conv_destroy": (None, [C.c_void_p]),
  "oidnb200_conv_weight_bytes": (C.c_size_t, [C.c_void_p]),
  "oidnb200_conv_bias_bytes": (C.c_size_t, [C.c_void_p]),
  "oidnb200_conv_pack_weights": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
  "oidnb200_conv_pack_bias": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
  "oidnb200_conv_bind": (C.c_int, [C.c_void_p] * 6),
  "oidnb200_conv_launch": (C.c_int, [C.c_void_p, C.c_void_p]),
  "oidnb200_conv_set_output_process": (C.c_int, [C.c_void_p, C.POINTER(Tile), C.POINTER(Transfer), C.c_int, C.c_int,
                                                 C.POINTER(Image)]),
  "oidnb200_conv_launch_simt": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
  "oidnb200_conv_set_trace": (C.c_int, [C.c_void_p, C.c_void_p]),
  "oidnb200_conv_pair_create": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
  "oidnb200_conv_pair_destroy": (None, [C.c_void_p]),
  "oidnb200_conv_pair_bind": (C.c_int, [C.c_void_p]),
  "oidnb200_conv_pair_launch": (C.c_int, [C.c_void_p, C.c_void_p]),
  "oidnb200_conv_pair_get_info": (C.c_int, [C.c_void_p, C.POINTER(ConvInfo)]),
  "oidnb200_conv_set_stamps": (C.c_int, [C.c_void_p, C.c_void_p]),
  "oidnb200_conv_get_info": (C.c_int, [C.c_void_p, C.POINTER(ConvInfo)]),
  "oidnb200_input_process_launch": (C.c_int, [C.POINTER(Image), C.POINTER(Image), C.POINTER(Image), C.POINTER(Tile),
                                              C.POINTER(Transfer), C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                              C.c_int, C.c_void_p]),
  "oidnb200_output_process_launch": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(Tile),
                                               C.POINTER(Transfer), C.c_int, C.c_int, C.POINTER(Image), C.c_void_p]),
  "oidnb200_autoexposure_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int]),
  "oidnb200_autoexposure_launch": (C.c_int, [C.POINTER(Image), C.c_void_p, C.c_void_p, C.c_void_p]),
  "oidnb200_autoexposure_bin_grid": (None, [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
  "oidnb200_autoexposure_bins_launch": (C.c_int, [C.POINTER(Image), C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
  "oidnb200_autoexposure_reduce_launch": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
  "oidnb200_flag_signal_launch": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_uint, C.c_void_p]),
  "oidnb200_flag_wait_launch": (C.c_int, [C.c_void_p, C.c_int, C.c_uint, C.c_double, C.c_void_p]),
  "oidnb200_image_copy_launch": (C.c_int, [C.POINTER(Image), C.POINTER(Image), C.c_void_p]),
  "oidnb200_pool_launch": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
  "oidnb200_upsample_launch": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
}

FILTER_ABI = {
  "oidnb200GetNumPhysicalDevices": (C.c_int, []),
  "oidnb200NewCUDADevice": (C.c_void_p, [C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.c_int]),
  "oidnb200NewDevice": (C.c_void_p, []),
  "oidnb200RetainDevice": (None, [C.c_void_p]),
  "oidnb200ReleaseDevice": (None, [C.c_void_p]),
  "oidnb200SetDeviceInt": (None, [C.c_void_p, C.c_char_p, C.c_int]),
  "oidnb200GetDeviceInt": (C.c_int, [C.c_void_p, C.c_char_p]),
  "oidnb200SetDeviceString": (None, [C.c_void_p, C.c_char_p, C.c_char_p]),
  "oidnb200CommitDevice": (None, [C.c_void_p]),
  "oidnb200SyncDevice": (None, [C.c_void_p]),
  "oidnb200GetDeviceError": (C.c_int, [C.c_void_p, C.POINTER(C.c_char_p)]),
  "oidnb200NewBufferWithStorage": (C.c_void_p, [C.c_void_p, C.c_size_t, C.c_int]),
  "oidnb200NewBuffer": (C.c_void_p, [C.c_void_p, C.c_size_t]),
  "oidnb200GetBufferData": (C.c_void_p, [C.c_void_p]),
  "oidnb200GetBufferSize": (C.c_size_t, [C.c_void_p]),
  "oidnb200ReadBuffer": (None, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]),
  "oidnb200WriteBuffer": (None, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]),
  "oidnb200ReadBufferAsync": (None, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]),
  "oidnb200WriteBufferAsync": (None, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]),
  "oidnb200ReleaseBuffer": (None, [C.c_void_p]),
  "oidnb200CopyRectAsync": (None, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t]),
  "oidnb200GetBufferIpcHandle": (None, [C.c_void_p, C.c_void_p]),
  "oidnb200NewSharedBufferFromIpcHandle": (C.c_void_p, [C.c_void_p, C.c_void_p, C.c_size_t]),
  "oidnb200NewSharedBufferFromFD": (C.c_void_p, [C.c_void_p, C.c_int, C.c_int, C.c_size_t]),
  "oidnb200NewExportableBuffer": (C.c_void_p, [C.c_void_p, C.c_size_t]),
  "oidnb200GetBufferFD": (C.c_int, [C.c_void_p]),
  "oidnb200NewFilter": (C.c_void_p, [C.c_void_p, C.c_char_p]),
  "oidnb200RetainFilter": (None, [C.c_void_p]),
  "oidnb200ReleaseFilter": (None, [C.c_void_p]),
  "oidnb200SetFilterImage": (None, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int] + [C.c_size_t] * 5),
  "oidnb200SetSharedFilterImage": (None, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int] + [C.c_size_t] * 5),
  "oidnb200UnsetFilterImage": (None, [C.c_void_p, C.c_char_p]),
  "oidnb200SetSharedFilterData": (None, [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]),
  "oidnb200UpdateFilterData": (None, [C.c_void_p, C.c_char_p]),
  "oidnb200UnsetFilterData": (None, [C.c_void_p, C.c_char_p]),
  "oidnb200SetFilterBool": (None, [C.c_void_p, C.c_char_p, C.c_bool]),
  "oidnb200GetFilterBool": (C.c_bool, [C.c_void_p, C.c_char_p]),
  "oidnb200SetFilterInt": (None, [C.c_void_p, C.c_char_p, C.c_int]),
  "oidnb200GetFilterInt": (C.c_int, [C.c_void_p, C.c_char_p]),
  "oidnb200SetFilterFloat": (None, [C.c_void_p, C.c_char_p, C.c_float]),
  "oidnb200GetFilterFloat": (C.c_float, [C.c_void_p, C.c_char_p]),
  "oidnb200SetFilterProgressMonitorFunction": (None, [C.c_void_p, PROGRESS_FUNC, C.c_void_p]),
  "oidnb200CommitFilter": (None, [C.c_void_p]),
  "oidnb200ExecuteFilter": (None, [C.c_void_p]),
  "oidnb200ExecuteFilterAsync": (None, [C.c_void_p]),
  "oidnb200GetFilterInfo": (None, [C.c_void_p, C.POINTER(FilterInfo)]),
  "oidnb200GetFilterProfile": (C.c_int, [C.c_void_p, C.POINTER(OpTime), C.c_int]),
  "oidnb200ResetFilterProfile": (None, [C.c_void_p]),
  "oidnb200PlanTiles": (None, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_long, C.POINTER(TilePlan)]),
  "oidnb200PlanTilesMinOverlap": (None, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_long, C.POINTER(TilePlan)]),
  "oidnb200PlanTilesStripAware": (None, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_long, C.POINTER(TilePlan)]),
  "oidnb200EnumerateTiles": (C.c_int, [C.POINTER(TilePlan), C.POINTER(C.c_int), C.c_int]),
  "oidnb200ParseTZA": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_char_p)]),
  "oidnb200PlanArena": (C.c_size_t, [C.c_int, C.POINTER(C.c_size_t), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                     C.POINTER(C.c_size_t)]),
}

_lib = None


def lib():
  """The loaded library with typed prototypes. Raises if liboidn_b200.so has not been built
  (python -c 'import __graft_entry__ as g; g.build()' or make -C oidn_b200/csrc)."""
  global _lib
  if _lib is None:
    if not os.path.exists(LIB_PATH):
      raise RuntimeError("oidn_b200: %s is missing; build it with `make -C oidn_b200/csrc` "
                         "(there is no CPU or PyTorch fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    for table in (KERNEL_ABI, FILTER_ABI):
      for name, (res, args) in table.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
  return _lib
