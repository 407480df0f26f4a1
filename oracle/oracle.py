"""ctypes loader for the CPU oracle (oracle/oidn_oracle.c). TEST INFRASTRUCTURE ONLY:
importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
Nothing under oidn_b200/ may import this module."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboidn_oracle.so")

TF_LINEAR, TF_SRGB, TF_PU, TF_LOG = 0, 1, 2, 3


def build(force=False):
  if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "oidn_oracle.c")):
    subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liboidn_oracle.so"])
  return _LIB


class Image(C.Structure):
  _fields_ = [("ptr", C.c_void_p), ("is_half", C.c_int), ("C", C.c_int), ("H", C.c_int), ("W", C.c_int),
              ("pixel_stride", C.c_size_t), ("row_stride", C.c_size_t)]


class Tile(C.Structure):
  _fields_ = [(n, C.c_int) for n in ("hSrcBegin", "wSrcBegin", "hDstBegin", "wDstBegin", "H", "W")]


class Tiling(C.Structure):
  _fields_ = [(n, C.c_int) for n in ("H", "W", "tileH", "tileW", "tilePadH", "tilePadW", "tileCountH",
                                     "tileCountW", "tileOverlap", "tileAlignment")]


class Params(C.Structure):
  _fields_ = [("filter", C.c_int), ("hdr", C.c_int), ("srgb", C.c_int), ("directional", C.c_int),
              ("input_scale", C.c_float), ("max_tile_pixels", C.c_long), ("num_subdevices", C.c_int)]


class Stats(C.Structure):
  _fields_ = [("tileH", C.c_int), ("tileW", C.c_int), ("tileCountH", C.c_int), ("tileCountW", C.c_int),
              ("tileOverlap", C.c_int), ("large", C.c_int), ("input_scale", C.c_float)]


_lib = None


def lib():
  global _lib
  if _lib is None:
    build()
    L = C.CDLL(_LIB)
    L.oro_set_num_threads.restype = C.c_int; L.oro_set_num_threads.argtypes = [C.c_int]
    L.oro_tf_forward.restype = C.c_float; L.oro_tf_forward.argtypes = [C.c_int, C.c_float]
    L.oro_tf_inverse.restype = C.c_float; L.oro_tf_inverse.argtypes = [C.c_int, C.c_float]
    L.oro_tf_norm_scale.restype = C.c_float; L.oro_tf_norm_scale.argtypes = [C.c_int]
    L.oro_half_to_float.restype = C.c_float; L.oro_half_to_float.argtypes = [C.c_uint16]
    L.oro_float_to_half.restype = C.c_uint16; L.oro_float_to_half.argtypes = [C.c_float]
    L.oro_autoexposure.restype = C.c_float; L.oro_autoexposure.argtypes = [C.POINTER(Image)]
    L.oro_unet_forward.restype = C.c_int
    L.oro_unet_forward.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    L.oro_plan_tiles.restype = None
    L.oro_plan_tiles.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_long, C.POINTER(Tiling)]
    L.oro_filter_execute.restype = C.c_int
    L.oro_filter_execute.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(Image), C.POINTER(Image), C.POINTER(Image),
                                     C.POINTER(Image), C.POINTER(Params), C.POINTER(Stats)]
    L.oro_conv3x3.restype = None
    L.oro_conv3x3.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.oro_input_process.restype = None
    L.oro_input_process.argtypes = [C.POINTER(Image), C.POINTER(Image), C.POINTER(Image), C.POINTER(Tile), C.c_int,
                                    C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.oro_output_process.restype = None
    L.oro_output_process.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(Tile), C.c_int, C.c_int, C.c_int,
                                     C.c_float, C.POINTER(Image)]
    _lib = L
  return _lib


def image_of(arr):
  """Wrap an HxWxC (or HxW) float32/float16 numpy array (any strides) as an oracle image."""
  if arr is None:
    return Image(None, 0, 0, 0, 0, 0, 0)
  a = arr if arr.ndim == 3 else arr[:, :, None]
  assert a.dtype in (np.float32, np.float16) and a.strides[2] == a.itemsize
  return Image(a.ctypes.data, int(a.dtype == np.float16), a.shape[2], a.shape[0], a.shape[1], a.strides[1], a.strides[0])


def tf_forward(t, y):
  L = lib(); return np.array([L.oro_tf_forward(t, float(v)) for v in np.asarray(y, np.float32).ravel()], np.float32).reshape(np.shape(y))


def tf_inverse(t, x):
  L = lib(); return np.array([L.oro_tf_inverse(t, float(v)) for v in np.asarray(x, np.float32).ravel()], np.float32).reshape(np.shape(x))


def autoexposure(color):
  im = image_of(color); return float(lib().oro_autoexposure(C.byref(im)))


def unet_forward(tza, x_hwc):
  """Raw network on an HWC float32 array whose H, W are multiples of 16."""
  x = np.ascontiguousarray(x_hwc, np.float32); H, W, Cc = x.shape
  out = np.empty((H, W, 3), np.float32)
  buf = (C.c_char * len(tza)).from_buffer_copy(tza)
  rc = lib().oro_unet_forward(buf, len(tza), x.ctypes.data, H, W, Cc, out.ctypes.data, out.size)
  if rc < 0:
    raise RuntimeError("oro_unet_forward failed: %d" % rc)
  return out


def plan_tiles(H, W, large=False, dev_alignment=1, num_subdevices=1, max_tile_pixels=2160 * 2160):
  t = Tiling(); lib().oro_plan_tiles(H, W, int(large), dev_alignment, num_subdevices, max_tile_pixels, C.byref(t))
  return {n: getattr(t, n) for n, _ in Tiling._fields_}


def filter_execute(tza, color=None, albedo=None, normal=None, output=None, filter="RT", hdr=False, srgb=False,
                   directional=False, input_scale=float("nan"), max_tile_pixels=0, num_subdevices=1):
  """The whole RT / RTLightmap filter on the CPU. `output` is written in place; returns stats."""
  buf = (C.c_char * len(tza)).from_buffer_copy(tza)
  ims = [image_of(a) for a in (color, albedo, normal, output)]
  prm = Params(0 if filter == "RT" else 1, int(hdr), int(srgb), int(directional), input_scale, max_tile_pixels, num_subdevices)
  st = Stats()
  rc = lib().oro_filter_execute(buf, len(tza), C.byref(ims[0]), C.byref(ims[1]), C.byref(ims[2]), C.byref(ims[3]),
                                C.byref(prm), C.byref(st))
  if rc != 0:
    raise RuntimeError("oro_filter_execute failed: %d" % rc)
  return {n: getattr(st, n) for n, _ in Stats._fields_}
