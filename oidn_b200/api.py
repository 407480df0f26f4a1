"""Python host over the filter-level C ABI, with the vocabulary of the reference's C++ wrapper
(include/OpenImageDenoise/oidn.hpp: DeviceRef / BufferRef / FilterRef). Errors stored by the
library are raised as `Error` after each call (the C API itself never throws, api/api.cpp:17-31).

    dev = Device(); dev.commit()
    f = dev.new_filter("RT")
    f.set_image("color", color_tensor); f.set_image("output", out_tensor)   # HxWx3 CUDA tensors
    f.set("hdr", True); f.set_data("weights", tza_bytes)
    f.commit(); f.execute()
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import (FORMAT_FLOAT, FORMAT_HALF, QUALITY_BALANCED, QUALITY_DEFAULT, QUALITY_FAST,  # noqa: F401
                   QUALITY_HIGH, STORAGE_DEVICE, STORAGE_HOST, STORAGE_MANAGED)

ERROR_NAMES = ["None", "Unknown", "InvalidArgument", "InvalidOperation", "OutOfMemory", "UnsupportedHardware",
               "Cancelled"]


class Error(RuntimeError):
  def __init__(self, code, message):
    super().__init__("%s: %s" % (ERROR_NAMES[code] if 0 <= code < len(ERROR_NAMES) else code, message))
    self.code = code
    self.message = message


def _check(handle):
  msg = C.c_char_p()
  code = capi.lib().oidnb200GetDeviceError(handle, C.byref(msg))
  if code != capi.ERROR_NONE:
    raise Error(code, (msg.value or b"").decode())


def num_physical_devices():
  return capi.lib().oidnb200GetNumPhysicalDevices()


class Device:
  """oidnNewCUDADevice(ids, streams, n) + oidnCommitDevice. One engine per (GPU, stream) pair."""

  def __init__(self, device_ids=(0,), streams=None):
    ids = (C.c_int * len(device_ids))(*device_ids)
    st = None
    if streams is not None:
      st = (C.c_void_p * len(device_ids))(*[int(s) if s else None for s in streams])
    self.device_ids = tuple(device_ids)
    self.streams = tuple(int(s) if s else None for s in streams) if streams is not None else None
    self._h = capi.lib().oidnb200NewCUDADevice(ids, st, len(device_ids))
    if not self._h:
      _check(None)
      raise Error(capi.ERROR_UNKNOWN, "device creation failed")

  def commit(self):
    capi.lib().oidnb200CommitDevice(self._h); _check(self._h)
    return self

  def sync(self):
    capi.lib().oidnb200SyncDevice(self._h); _check(self._h)

  def set(self, name, value):
    if isinstance(value, str):
      capi.lib().oidnb200SetDeviceString(self._h, name.encode(), value.encode())
    else:
      capi.lib().oidnb200SetDeviceInt(self._h, name.encode(), int(value))
    _check(self._h)

  def get(self, name):
    v = capi.lib().oidnb200GetDeviceInt(self._h, name.encode()); _check(self._h)
    return v

  def get_error(self):
    msg = C.c_char_p()
    code = capi.lib().oidnb200GetDeviceError(self._h, C.byref(msg))
    return code, (msg.value or b"").decode()

  def new_filter(self, type):
    h = capi.lib().oidnb200NewFilter(self._h, type.encode()); _check(self._h)
    return Filter(self, h)

  def new_buffer(self, byte_size, storage=STORAGE_DEVICE):
    h = capi.lib().oidnb200NewBufferWithStorage(self._h, byte_size, storage); _check(self._h)
    return Buffer(self, h)

  def copy_rect_async(self, dst_ptr, dst_pitch, src_ptr, src_pitch, width_bytes, height):
    """Copy-engine 2D transfer in stream order (local / peer GPU / pinned host memory)."""
    capi.lib().oidnb200CopyRectAsync(self._h, dst_ptr, dst_pitch, src_ptr, src_pitch, width_bytes, height); _check(self._h)

  def import_buffer(self, ipc_handle, byte_size):
    """Opens a peer process's exported device buffer (CUDA IPC, same node)."""
    raw = (C.c_char * 64).from_buffer_copy(bytes(ipc_handle))
    h = capi.lib().oidnb200NewSharedBufferFromIpcHandle(self._h, raw, byte_size); _check(self._h)
    return Buffer(self, h)

  def new_exportable_buffer(self, byte_size):
    """Device buffer whose memory can be exported as an opaque fd (Buffer.fd())."""
    h = capi.lib().oidnb200NewExportableBuffer(self._h, byte_size); _check(self._h)
    return Buffer(self, h)

  def import_fd(self, fd, byte_size, fd_type=capi.EXTERNAL_MEMORY_OPAQUE_FD):
    """oidnNewSharedBufferFromFD: the buffer owns the fd on success."""
    h = capi.lib().oidnb200NewSharedBufferFromFD(self._h, fd_type, fd, byte_size); _check(self._h)
    return Buffer(self, h)

  def release(self):
    if self._h:
      capi.lib().oidnb200ReleaseDevice(self._h)
      self._h = None

  def __del__(self):
    try:
      self.release()
    except Exception:
      pass


class Buffer:
  def __init__(self, device, handle):
    self.device, self._h = device, handle

  @property
  def data(self):
    return capi.lib().oidnb200GetBufferData(self._h)

  @property
  def size(self):
    return capi.lib().oidnb200GetBufferSize(self._h)

  def ipc_handle(self):
    raw = (C.c_char * 64)()
    capi.lib().oidnb200GetBufferIpcHandle(self._h, raw); _check(self.device._h)
    return bytes(raw)

  def fd(self):
    fd = capi.lib().oidnb200GetBufferFD(self._h); _check(self.device._h)
    return fd

  def write(self, array, byte_offset=0, sync=True):
    a = np.ascontiguousarray(array)
    fn = capi.lib().oidnb200WriteBuffer if sync else capi.lib().oidnb200WriteBufferAsync
    fn(self._h, byte_offset, a.nbytes, a.ctypes.data); _check(self.device._h)

  def read(self, array, byte_offset=0, sync=True):
    assert array.flags["C_CONTIGUOUS"]
    fn = capi.lib().oidnb200ReadBuffer if sync else capi.lib().oidnb200ReadBufferAsync
    fn(self._h, byte_offset, array.nbytes, array.ctypes.data); _check(self.device._h)

  def release(self):
    if self._h:
      capi.lib().oidnb200ReleaseBuffer(self._h)
      self._h = None

  def __del__(self):
    try:
      self.release()
    except Exception:
      pass


def _format_of(dtype_name, channels):
  base = {"float32": FORMAT_FLOAT, "float16": FORMAT_HALF}[dtype_name]
  return base + channels - 1


class Filter:
  def __init__(self, device, handle):
    self.device, self._h = device, handle
    self._keep = {}

  def _ck(self):
    _check(self.device._h)

  def set_image(self, name, image, format=None, width=None, height=None, byte_offset=0, pixel_stride=0,
                row_stride=0):
    """image: a Buffer (+format,width,height), a raw device/pinned pointer (int, +format,width,
    height), or an HxWxC torch tensor / pinned numpy array whose strides are used as given."""
    L = capi.lib()
    if isinstance(image, Buffer):
      L.oidnb200SetFilterImage(self._h, name.encode(), image._h, format, width, height, byte_offset, pixel_stride,
                               row_stride)
    elif isinstance(image, int) or image is None:
      L.oidnb200SetSharedFilterImage(self._h, name.encode(), image, format or 0, width or 0, height or 0, byte_offset,
                                     pixel_stride, row_stride)
    else:
      t = image
      if t.ndim == 2:
        t = t[:, :, None]
      if hasattr(t, "data_ptr"):   # torch
        dt = str(t.dtype).replace("torch.", "")
        es = t.element_size()
        ptr, strides = t.data_ptr(), [s * es for s in t.stride()]
      else:                        # numpy (must be GPU-accessible memory, e.g. pinned)
        dt, ptr, strides = t.dtype.name, t.ctypes.data, list(t.strides)
      assert strides[2] == (2 if dt == "float16" else 4), "channels must be contiguous"
      H, W, Cc = t.shape
      L.oidnb200SetSharedFilterImage(self._h, name.encode(), ptr, _format_of(dt, Cc), W, H, 0, strides[1], strides[0])
      self._keep[name] = image
    self._ck()

  def unset_image(self, name):
    capi.lib().oidnb200UnsetFilterImage(self._h, name.encode()); self._ck()
    self._keep.pop(name, None)

  def set_data(self, name, data):
    """Borrowed host memory (oidnSetSharedFilterData): kept alive by this object."""
    if data is None:
      capi.lib().oidnb200SetSharedFilterData(self._h, name.encode(), None, 0)
    else:
      buf = (C.c_char * len(data)).from_buffer_copy(bytes(data))
      self._keep["data:" + name] = buf
      capi.lib().oidnb200SetSharedFilterData(self._h, name.encode(), C.addressof(buf), len(data))
    self._ck()

  def set_input_scale_ptr(self, dev_ptr):
    """Backend extension for sharded execution: the input scale is read from this device float."""
    capi.lib().oidnb200SetSharedFilterData(self._h, b"inputScalePtr", dev_ptr, 4 if dev_ptr else 0); self._ck()

  def update_data(self, name):
    capi.lib().oidnb200UpdateFilterData(self._h, name.encode()); self._ck()

  def unset_data(self, name):
    capi.lib().oidnb200UnsetFilterData(self._h, name.encode()); self._ck()

  def set(self, name, value):
    L = capi.lib()
    if isinstance(value, bool):
      L.oidnb200SetFilterBool(self._h, name.encode(), value)
    elif isinstance(value, int):
      L.oidnb200SetFilterInt(self._h, name.encode(), value)
    else:
      L.oidnb200SetFilterFloat(self._h, name.encode(), float(value))
    self._ck()

  def get_int(self, name):
    v = capi.lib().oidnb200GetFilterInt(self._h, name.encode()); self._ck()
    return v

  def get_float(self, name):
    v = capi.lib().oidnb200GetFilterFloat(self._h, name.encode()); self._ck()
    return v

  def set_progress_monitor(self, func):
    cb = capi.PROGRESS_FUNC(lambda _u, n: bool(func(n))) if func else capi.PROGRESS_FUNC()
    self._keep["progress"] = cb
    capi.lib().oidnb200SetFilterProgressMonitorFunction(self._h, cb, None); self._ck()

  def commit(self):
    capi.lib().oidnb200CommitFilter(self._h); self._ck()
    return self

  def execute(self):
    capi.lib().oidnb200ExecuteFilter(self._h); self._ck()

  def execute_async(self):
    capi.lib().oidnb200ExecuteFilterAsync(self._h); self._ck()

  def info(self):
    i = capi.FilterInfo()
    capi.lib().oidnb200GetFilterInfo(self._h, C.byref(i)); self._ck()
    return {n: getattr(i, n) for n, _ in capi.FilterInfo._fields_}

  def profile(self, reset=True):
    """[(name, kind, launches, ms)] per op since the last reset (device param "profile" = 1)."""
    arr = (capi.OpTime * 64)()
    n = capi.lib().oidnb200GetFilterProfile(self._h, arr, 64); self._ck()
    out = [(arr[i].name.decode(), arr[i].kind, arr[i].launches, arr[i].ms) for i in range(min(n, 64))]
    if reset:
      capi.lib().oidnb200ResetFilterProfile(self._h)
    return out

  def release(self):
    if self._h:
      capi.lib().oidnb200ReleaseFilter(self._h)
      self._h = None

  def __del__(self):
    try:
      self.release()
    except Exception:
      pass


def plan_tiles(H, W, large=False, dev_alignment=1, num_engines=1, max_tile_pixels=2160 * 2160, policy=0):
  """Tile planner alone (no GPU): returns (plan dict, [tile rect dicts]). policy 0 = the reference's
  search (core/unet_filter.cpp:254-335), 1 = this backend's default (fewest recomputed pixels),
  2 = fewest 128-pixel conv strips (opt-in)."""
  p = capi.TilePlan()
  L = capi.lib()
  fn = (L.oidnb200PlanTiles, L.oidnb200PlanTilesMinOverlap, L.oidnb200PlanTilesStripAware)[int(policy)]
  fn(H, W, int(large), dev_alignment, num_engines, max_tile_pixels, C.byref(p))
  n = p.tileCountH * p.tileCountW
  out = (C.c_int * (12 * n))()
  capi.lib().oidnb200EnumerateTiles(C.byref(p), out, n)
  names = ("hSrc", "wSrc", "hBuf", "wBuf", "H1", "W1", "hOutBuf", "wOutBuf", "hDst", "wDst", "H2", "W2")
  tiles = [dict(zip(names, out[12 * i:12 * i + 12])) for i in range(n)]
  return {k: getattr(p, k) for k, _ in capi.TilePlan._fields_}, tiles
