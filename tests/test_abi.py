"""CPU tier: the C-ABI library loads without a GPU and exports every symbol include/*.h declares."""
import ctypes as C
import os
import re

import numpy as np

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


def test_conv_planner_stream_and_store_choices(monkeypatch):
  """Every UNet layer shape (base, small, large) plans inside the hardware limits under the default rules and
  under the probing overrides; the concat layers of the base net get their second stream from the
  direct-store epilogue (out_nbuf == 0), narrow layers keep staged stores."""
  L = capi.lib()
  from oidn_b200 import weights
  shapes = set()
  for kind, ic in (("base", 9), ("base", 3), ("small", 3), ("large", 9)):
    prev_out = 0
    for name, cin, cout in weights.unet_layers(kind, ic):
      n = int(name[8])                                  # enc_convN.. / dec_convN..
      level = max(n - 1, 0)
      concat = name.startswith("dec") and name.endswith("a")
      in1 = prev_out if concat else cin
      in2 = cin - in1 if concat else 0
      pool = name.startswith("enc") and 1 <= n <= 4 and (kind != "large" or name.endswith("b"))
      pad = lambda c: -(-c // 16) * 16
      shapes.add((name if (kind, ic) == ("base", 9) else "", 2160 >> level, 3840 >> level, pad(in1), pad(in2) if in2 else 0,
                  pad(cout), int(pool), int(concat)))
      prev_out = cout
  assert len(shapes) > 30

  def plan(shape):
    _, H, W, c1, c2, co, post, up = shape
    d = capi.ConvDesc(H, W, c1, c2, co, 1, post, up, 0); h = C.c_void_p()
    assert L.oidnb200_conv_create(C.byref(d), C.byref(h)) == 0, (shape, L.oidnb200_last_error())
    i = capi.ConvInfo(); L.oidnb200_conv_get_info(h, C.byref(i)); L.oidnb200_conv_destroy(h)
    assert 1 <= i.grid <= 148 and i.smem_bytes <= 232448, shape
    assert i.nstreams in (1, 2, 4) and 2 <= i.nstages <= 24 and 0 <= i.out_nbuf <= 2, shape
    assert i.nstreams * i.ring_slots * i.cout_group <= 512 and i.ring_slots >= 4, shape
    assert i.nstreams == 1 or i.cout_group <= (32 if i.nstreams == 4 else 64), shape
    return i

  base = {s[0]: plan(s) for s in shapes if s[0]}
  for s in shapes:
    plan(s)
  assert all(i.nstreams <= 2 for i in base.values())                       # four streams are opt-in
  assert base["dec_conv2a"].nstreams == 2 and base["dec_conv2a"].out_nbuf == 0
  assert base["dec_conv3a"].nstreams == 2 and base["dec_conv3a"].out_nbuf == 0
  assert base["enc_conv0"].out_nbuf == 2 and base["dec_conv1a"].out_nbuf >= 1 and base["dec_conv1a"].nstreams == 2
  for env in ({"OIDN_B200_STREAMS": "4"}, {"OIDN_B200_STREAMS": "1"}, {"OIDN_B200_DIRECT_STORE": "1"},
              {"OIDN_B200_DIRECT_STORE": "0", "OIDN_B200_STREAMS": "2"}):
    for k, v in env.items():
      monkeypatch.setenv(k, v)
    infos = [plan(s) for s in shapes]
    if env.get("OIDN_B200_STREAMS") == "4":
      assert any(i.nstreams == 4 for i in infos)
    if env.get("OIDN_B200_STREAMS") == "1":
      assert all(i.nstreams == 1 for i in infos)
    if env.get("OIDN_B200_DIRECT_STORE") == "1":
      assert all(i.out_nbuf == 0 for i in infos)
    for k in env:
      monkeypatch.delenv(k)


def test_fused_output_process_argument_checks_without_gpu():
  """oidnb200_conv_set_output_process validates on the host: only the 16-channel last conv and packed fp32 RGB
  images are accepted, the tile must lie inside tensor and image, NULL removes the fusion."""
  L = capi.lib()
  mk = lambda co: (lambda d, h: (L.oidnb200_conv_create(C.byref(d), C.byref(h)), h)[1])(capi.ConvDesc(64, 256, 32, 0, co, 1, 0, 0, 0), C.c_void_p())
  last, wide = mk(16), mk(32)
  tf = capi.Transfer(capi.TF_PU, 1.0, None)
  img = lambda fmt, ps, rs, W=256, H=64: capi.Image(0x1000, fmt, W, H, ps, rs)
  ok_tile, bad_tile = capi.Tile(0, 0, 0, 0, 64, 256), capi.Tile(0, 0, 0, 0, 65, 256)
  f3 = capi.FORMAT_FLOAT + 2
  assert L.oidnb200_conv_set_output_process(last, C.byref(ok_tile), C.byref(tf), 1, 0, C.byref(img(f3, 12, 256 * 12))) == 0
  assert L.oidnb200_conv_set_output_process(last, None, None, 0, 0, None) == 0
  assert L.oidnb200_conv_set_output_process(wide, C.byref(ok_tile), C.byref(tf), 1, 0, C.byref(img(f3, 12, 256 * 12))) == -2
  assert L.oidnb200_conv_set_output_process(last, C.byref(ok_tile), C.byref(tf), 1, 0, C.byref(img(capi.FORMAT_HALF + 2, 6, 256 * 6))) == -2
  assert L.oidnb200_conv_set_output_process(last, C.byref(ok_tile), C.byref(tf), 1, 0, C.byref(img(f3, 16, 256 * 16))) == -2
  assert L.oidnb200_conv_set_output_process(last, C.byref(bad_tile), C.byref(tf), 1, 0, C.byref(img(f3, 12, 256 * 12))) == -1
  bad_tf = capi.Transfer(99, 1.0, None)
  assert L.oidnb200_conv_set_output_process(last, C.byref(ok_tile), C.byref(bad_tf), 1, 0, C.byref(img(f3, 12, 256 * 12))) == -1
  assert L.oidnb200_last_error()
  for h in (last, wide):
    L.oidnb200_conv_destroy(h)


def test_conv_pair_planner_without_gpu():
  """kernels/conv_pair_tc.cu's planner (host side): which conv -> conv pairs fuse, and with what rings. The TMEM
  columns of the two accumulator rings, and the shared memory of weights + input ring + mid ring, must fit."""
  L = capi.lib()

  def conv(H, W, C1, C2, Cout, post=0):
    d = capi.ConvDesc(H, W, C1, C2, Cout, 1, post, 0, 0)
    h = C.c_void_p()
    assert L.oidnb200_conv_create(C.byref(d), C.byref(h)) == 0
    return h

  def plan(a, b):
    p = C.c_void_p()
    rc = L.oidnb200_conv_pair_create(a, b, C.byref(p))
    if rc != 0:
      return rc, None
    i = capi.ConvInfo()
    assert L.oidnb200_conv_pair_get_info(p, C.byref(i)) == 0
    L.oidnb200_conv_pair_destroy(p)
    return 0, {n: getattr(i, n) for n, _ in capi.ConvInfo._fields_}

  H, W = 2160, 3840
  cases = {  # (C1 of A, Cout of A, Cout of B, pool) -> streams
    (16, 32, 32, 1): 2,    # base / small: enc_conv0 -> enc_conv1 + pool
    (64, 32, 16, 0): 2,    # base: dec_conv1b -> dec_conv0
    (32, 32, 16, 0): 2,    # small: dec_conv1b -> dec_conv0
    (16, 64, 64, 1): 1,    # large: enc_conv1a -> enc_conv1b + pool (accumulators need all of TMEM)
    (64, 64, 16, 0): 1,    # large: dec_conv1b -> dec_conv1c (92 KB of resident weights leave room for one stream)
  }
  for (c1, ca, cb, pool), streams in cases.items():
    rc, i = plan(conv(H, W, c1, 0, ca), conv(H, W, ca, 0, cb, pool))
    assert rc == 0, (c1, ca, cb, pool, L.oidnb200_last_error())
    # get_info of a pair: ngroups = mid stages, nchunks = A's accumulator slots, ring_slots = B's, nstages = input stages
    assert i["nstreams"] == streams and i["smem_bytes"] <= 232448 and i["grid"] == 148
    assert i["nstreams"] * (i["nchunks"] * ca + i["ring_slots"] * cb) <= 512
    assert i["nchunks"] >= 3 and i["ring_slots"] >= 4 and i["ngroups"] >= 3 and i["nstages"] >= 3
    assert i["nstrips"] == -(-W // 126) and i["rows_per_item"] % (2 if pool else 1) == 0
    assert i["nrowchunks"] * i["rows_per_item"] >= H
  # not fused: concat source, upsampled source, > 64 channels, A with a post-op, different resolution
  assert plan(conv(H, W, 16, 0, 32), conv(H, W, 32, 16, 32))[0] == -2
  assert plan(conv(H, W, 96, 0, 96), conv(H, W, 96, 0, 96))[0] == -2
  assert plan(conv(H, W, 16, 0, 32, 1), conv(H // 2, W // 2, 32, 0, 32))[0] == -2
  assert plan(conv(H, W, 16, 0, 32), conv(H // 2, W // 2, 32, 0, 32))[0] == -2


def test_row_folding_plan_and_weights_without_gpu(monkeypatch):
  """ConvKernelParams::up_fold (host side): the planner folds an upsampled source only when it keeps the channel group
  and both streams (dec_conv1a), and oidnb200_conv_pack_weights appends the vertically pre-summed taps
  [kw][w0+w1, w1+w2, w2, w0][CoutAlloc][C1] behind the regular [kw][kh][CoutAlloc][CinTot] block."""
  L = capi.lib()

  def make(H, W, C1, C2, Co, up):
    d = capi.ConvDesc(H, W, C1, C2, Co, 1, 0, up, 0)
    h = C.c_void_p()
    assert L.oidnb200_conv_create(C.byref(d), C.byref(h)) == 0, L.oidnb200_last_error()
    return h

  regular = lambda C1, C2, Co: 9 * Co * (C1 + C2) * 2
  # dec_conv1a (64 upsampled + 16 skip -> 64): folded; dec_conv2a (96 + 32 -> 64): the folded weights do not fit -> regular
  a = make(2160, 3840, 64, 16, 64, 1)
  assert L.oidnb200_conv_weight_bytes(a) == regular(64, 16, 64) + 12 * 64 * 64 * 2
  i = capi.ConvInfo(); L.oidnb200_conv_get_info(a, C.byref(i))
  assert i.nstreams == 2 and i.cout_group == 64 and i.ring_slots == 4 and i.smem_bytes <= 232448
  assert L.oidnb200_conv_weight_bytes(make(1080, 1920, 96, 32, 64, 1)) == regular(96, 32, 64)
  assert L.oidnb200_conv_weight_bytes(make(2160, 3840, 64, 16, 64, 0)) == regular(64, 16, 64)      # not upsampled
  monkeypatch.setenv("OIDN_B200_NO_UPFOLD", "1")
  assert L.oidnb200_conv_weight_bytes(make(2160, 3840, 64, 16, 64, 1)) == regular(64, 16, 64)
  monkeypatch.delenv("OIDN_B200_NO_UPFOLD")

  O, I1, I2, C1, C2, Co = 61, 64, 9, 64, 16, 64
  rng = np.random.default_rng(5)
  w = (rng.standard_normal((O, I1 + I2, 3, 3)) * 0.1).astype(np.float16)
  buf = np.zeros(L.oidnb200_conv_weight_bytes(a) // 2, np.float16)
  assert L.oidnb200_conv_pack_weights(a, w.ctypes.data, O, I1, I2, buf.ctypes.data) == 0
  reg = buf[:9 * Co * (C1 + C2)].reshape(3, 3, Co, C1 + C2)          # [kw][kh][o][ci]
  fol = buf[9 * Co * (C1 + C2):].reshape(3, 4, Co, C1)               # [kw][E0,E1,E2,O][o][i]
  np.testing.assert_array_equal(reg[:, :, :O, :I1], w[:, :I1].transpose(3, 2, 0, 1))
  np.testing.assert_array_equal(reg[:, :, :O, C1:C1 + I2], w[:, I1:].transpose(3, 2, 0, 1))
  wf = w[:, :I1].astype(np.float32)                                   # [o][i][kh][kw]
  stacks = [wf[:, :, 0] + wf[:, :, 1], wf[:, :, 1] + wf[:, :, 2], wf[:, :, 2], wf[:, :, 0]]
  for j, st in enumerate(stacks):
    np.testing.assert_array_equal(fol[:, j, :O, :I1], st.astype(np.float16).transpose(2, 0, 1))
  assert not fol[:, :, O:].any() and not fol[:, :, :, I1:].any()      # padding stays zero


def test_tap_packed_last_conv_weights_without_gpu():
  """The 3-channel last conv of a fused pair (dec_conv0 / dec_conv1c): oidnb200_conv_pack_weights appends a copy with the
  horizontal taps as columns, [kh][kw * 3 + c (16 columns)][C1], behind the regular [kw][kh][16][C1] block; other
  convs do not carry it."""
  L = capi.lib()

  def make(H, W, C1, C2, Co, post=0):
    d = capi.ConvDesc(H, W, C1, C2, Co, 1, post, 0, 0)
    h = C.c_void_p()
    assert L.oidnb200_conv_create(C.byref(d), C.byref(h)) == 0, L.oidnb200_last_error()
    return h

  for C1 in (32, 64):
    last = make(2160, 3840, C1, 0, 16)
    assert L.oidnb200_conv_weight_bytes(last) == (9 * 16 * C1 + 48 * C1) * 2
    rng = np.random.default_rng(C1)
    w = (rng.standard_normal((3, C1, 3, 3)) * 0.1).astype(np.float16)
    buf = np.full(L.oidnb200_conv_weight_bytes(last) // 2, np.nan, np.float16)
    assert L.oidnb200_conv_pack_weights(last, w.ctypes.data, 3, C1, 0, buf.ctypes.data) == 0
    reg = buf[:9 * 16 * C1].reshape(3, 3, 16, C1)                     # [kw][kh][o][ci]
    tap = buf[9 * 16 * C1:].reshape(3, 16, C1)                        # [kh][kw * 3 + o][ci]
    np.testing.assert_array_equal(reg[:, :, :3], w.transpose(3, 2, 0, 1))
    for kh in range(3):
      for kw in range(3):
        np.testing.assert_array_equal(tap[kh, kw * 3:kw * 3 + 3], w[:, :, kh, kw])
    assert not tap[:, 9:].any() and not reg[:, :, 3:].any()
    # a 4-channel output does not fit 3 taps x C <= 16 columns per row: the copy stays zero and is not used
    w4 = (rng.standard_normal((4, C1, 3, 3)) * 0.1).astype(np.float16)
    assert L.oidnb200_conv_pack_weights(last, w4.ctypes.data, 4, C1, 0, buf.ctypes.data) == 0
    assert not buf[9 * 16 * C1:].any()
  assert L.oidnb200_conv_weight_bytes(make(2160, 3840, 32, 0, 32)) == 9 * 32 * 32 * 2     # not a 16-column conv
  assert L.oidnb200_conv_weight_bytes(make(2160, 3840, 48, 0, 16)) == 9 * 16 * 48 * 2     # source width not covered
