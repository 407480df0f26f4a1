"""CPU tier: the C-ABI library loads without a GPU and exports every symbol include/*.h declares."""
import ctypes as C
import os
import re

from oidn_b200 import api, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header):
  text = open(os.path.join(ROOT, "include", header)).read()
  text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
  return re.findall(r"OIDNB200_API\s+[\w\s\*]+?\b(oidnb200\w+)\s*\(", text)


def test_every_declared_symbol_is_exported_and_bound():
  L = capi.lib()
  k = declared_symbols("oidn_b200_kernels.h"); f = declared_symbols("oidn_b200.h")
  assert len(k) >= 19 and len(f) >= 40
  for name in k + f:
    assert hasattr(L, name), "library does not export " + name
  assert set(k) == set(capi.KERNEL_ABI), set(k) ^ set(capi.KERNEL_ABI)
  assert set(f) == set(capi.FILTER_ABI), set(f) ^ set(capi.FILTER_ABI)


def test_no_gpu_is_reported_not_emulated():
  """Without a device the product path fails loudly (UnsupportedHardware), it never falls back."""
  import torch
  if torch.cuda.is_available():
    return
  assert api.num_physical_devices() == 0
  d = api.Device((0,))
  try:
    d.commit()
    raise AssertionError("commit succeeded without a GPU")
  except api.Error as e:
    assert e.code == capi.ERROR_UNSUPPORTED_HARDWARE


def test_product_does_not_reference_the_oracle():
  """oracle/ is test infrastructure: nothing under oidn_b200/ may import, link or load it."""
  bad = []
  for dirpath, _, files in os.walk(os.path.join(ROOT, "oidn_b200")):
    for fn in files:
      if fn.endswith((".py", ".cpp", ".cu", ".h", ".hpp", ".cuh", "Makefile")):
        txt = open(os.path.join(dirpath, fn), errors="ignore").read()
        if re.search(r"oidn_oracle|liboidn_oracle|import oracle|from oracle|oracle/", txt):
          bad.append(os.path.join(dirpath, fn))
  assert not bad, bad
  out = os.popen("ldd %s" % capi.LIB_PATH).read()
  assert "oracle" not in out


def test_conv_planner_runs_without_gpu():
  """Work decomposition is host logic: grids are sized against 148 SMs, smem within 227 KB."""
  L = capi.lib()
  for (H, W, C1, C2, Co, post, up) in [(2160, 3840, 64, 16, 64, 0, 1), (2160, 3840, 32, 0, 32, 1, 0), (1080, 1920, 96, 32, 64, 0, 1),
                                       (135, 240, 80, 0, 96, 0, 0), (270, 480, 256, 128, 192, 0, 1)]:
    d = capi.ConvDesc(H, W, C1, C2, Co, 1, post, up, 0)
    h = C.c_void_p()
    assert L.oidnb200_conv_create(C.byref(d), C.byref(h)) == 0, L.oidnb200_last_error()
    i = capi.ConvInfo(); L.oidnb200_conv_get_info(h, C.byref(i))
    assert 1 <= i.grid <= 148 and i.smem_bytes <= 232448 and i.nstages >= 2
    assert i.ngroups * i.cout_group >= Co and i.nstreams * i.ring_slots * i.cout_group <= 512
    assert i.nstrips == -(-W // 128) and i.nrowchunks * i.rows_per_item >= H
    L.oidnb200_conv_destroy(h)
