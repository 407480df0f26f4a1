"""Helpers for the -m gpu tests: device memory through torch, calls through the C ABI (ctypes)."""
import ctypes as C

import numpy as np

from oidn_b200 import capi


def check(rc):
  if rc != 0:
    raise RuntimeError("oidn_b200 ABI error %d: %s" % (rc, capi.lib().oidnb200_last_error().decode()))


def image_of(t):
  """capi.Image for an HxWxC torch tensor (any strides with contiguous channels)."""
  if t is None:
    return capi.Image(None, 0, 0, 0, 0, 0)
  if t.ndim == 2:
    t = t[:, :, None]
  H, W, Cc = t.shape
  es = t.element_size()
  fmt = (capi.FORMAT_HALF if es == 2 else capi.FORMAT_FLOAT) + Cc - 1
  return capi.Image(t.data_ptr(), fmt, W, H, t.stride(1) * es, t.stride(0) * es)


def metrics(got, ref):
  got = np.asarray(got, np.float64); ref = np.asarray(ref, np.float64)
  peak = max(np.abs(ref).max(), 1e-30)
  maxerr = np.abs(got - ref).max() / peak
  mse = np.mean((got - ref) ** 2)
  psnr = 200.0 if mse == 0 else 20 * np.log10(peak / np.sqrt(mse))
  return maxerr, psnr


class ConvOp:
  """One conv through the kernel-level C ABI on torch-allocated fp16 NHWC tensors."""

  def __init__(self, H, W, C1, C2, Cout, relu=1, post_op=0, up=0):
    L = capi.lib()
    self.d = capi.ConvDesc(H, W, C1, C2, Cout, relu, post_op, up, 0)
    self.h = C.c_void_p()
    check(L.oidnb200_conv_create(C.byref(self.d), C.byref(self.h)))

  def info(self):
    i = capi.ConvInfo(); capi.lib().oidnb200_conv_get_info(self.h, C.byref(i))
    return {n: getattr(i, n) for n, _ in capi.ConvInfo._fields_}

  def run(self, src1, src2, w_oihw, bias, I1, I2, simt=False, fused=None):
    """src1/src2: torch fp16 [H][W][Cpad] cuda; w_oihw: np.float16 [O][I1+I2][3][3]; bias np.float16 [O].
    fused = (capi.Tile, capi.Transfer, hdr, snorm, capi.Image): output process inside the epilogue."""
    import torch
    L = capi.lib()
    O = w_oihw.shape[0]
    wb = np.zeros(L.oidnb200_conv_weight_bytes(self.h), np.uint8)
    bb = np.zeros(L.oidnb200_conv_bias_bytes(self.h), np.uint8)
    w = np.ascontiguousarray(w_oihw.astype(np.float16)); b = np.ascontiguousarray(bias.astype(np.float16))
    check(L.oidnb200_conv_pack_weights(self.h, w.ctypes.data, O, I1, I2, wb.ctypes.data))
    check(L.oidnb200_conv_pack_bias(self.h, b.ctypes.data, O, bb.ctypes.data))
    dw = torch.from_numpy(wb).cuda(); db = torch.from_numpy(bb).cuda()
    d = self.d
    Ho, Wo = (d.H // 2, d.W // 2) if d.post_op == 1 else ((d.H * 2, d.W * 2) if d.post_op == 2 else (d.H, d.W))
    out = torch.full((Ho, Wo, d.Cout), float("nan"), dtype=torch.float16, device="cuda")
    check(L.oidnb200_conv_bind(self.h, src1.data_ptr(), src2.data_ptr() if src2 is not None else None,
                               dw.data_ptr(), db.data_ptr(), out.data_ptr()))
    st = torch.cuda.current_stream().cuda_stream
    if fused is not None:
      tile, tf, hdr, snorm, img = fused
      check(L.oidnb200_conv_set_output_process(self.h, C.byref(tile), C.byref(tf), hdr, snorm, C.byref(img)))
    if simt:
      scratch = torch.empty((d.H, d.W, d.Cout), dtype=torch.float16, device="cuda")
      check(L.oidnb200_conv_launch_simt(self.h, scratch.data_ptr(), st))
    else:
      check(L.oidnb200_conv_launch(self.h, st))
    torch.cuda.synchronize()
    self._keep = (dw, db)
    return out

  def __del__(self):
    try:
      capi.lib().oidnb200_conv_destroy(self.h)
    except Exception:
      pass
