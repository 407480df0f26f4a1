"""-m gpu: every op of the hot path through the kernel-level C ABI against the CPU oracle."""
import ctypes as C

import numpy as np
import pytest

from oidn_b200 import capi, synth

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from gpu_util import ConvOp, check, image_of, metrics  # noqa: E402


def _rand_half(rng, shape, lo=0.0, hi=1.0):
  return (rng.random(shape, dtype=np.float32) * (hi - lo) + lo).astype(np.float16)


def _oracle_conv(oracle, src, w, b, relu):
  H, W, Cin = src.shape
  O = w.shape[0]
  out = np.empty((H, W, O), np.float32)
  s = np.ascontiguousarray(src, np.float32); wf = np.ascontiguousarray(w, np.float32); bf = np.ascontiguousarray(b, np.float32)
  oracle.lib().oro_conv3x3(s.ctypes.data, H, W, Cin, wf.ctypes.data, bf.ctypes.data, O, relu, out.ctypes.data)
  return out


# (H, W, I1, I2, O, post_op, up)  logical channels; padded to 16 inside
CONV_CASES = [
  (16, 128, 9, 0, 32, 0, 0),     # enc_conv0
  (24, 300, 32, 0, 32, 1, 0),    # enc_conv1 + pool, ragged width
  (38, 200, 32, 0, 48, 1, 0),
  (16, 160, 48, 0, 64, 1, 0),
  (12, 96, 64, 0, 80, 1, 0),
  (8, 48, 80, 0, 96, 0, 0),
  (8, 48, 96, 0, 96, 0, 0),
  (16, 96, 96, 64, 112, 0, 1),   # dec_conv4a: upsampled src1 + skip
  (16, 96, 112, 0, 112, 0, 0),
  (32, 136, 112, 48, 96, 0, 1),
  (32, 260, 96, 32, 64, 0, 1),
  (32, 260, 64, 9, 64, 0, 1),    # dec_conv1a
  (20, 130, 64, 0, 32, 0, 0),
  (20, 130, 32, 0, 3, 0, 0),     # dec_conv0
  (18, 70, 256, 128, 192, 0, 1), # large dec_conv4a
  (18, 70, 192, 0, 256, 0, 0),
  (2, 2, 9, 0, 32, 1, 0),        # smallest poolable tile
  (1, 1, 32, 0, 32, 0, 0),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=["%dx%d_%d+%d_%d_p%d_u%d" % c for c in CONV_CASES])
def test_conv_matches_oracle(case, oracle):
  H, W, I1, I2, O, post, up = case
  rng = np.random.default_rng(hash(case) & 0xFFFF)
  C1, C2, Co = -(-I1 // 16) * 16, (-(-I2 // 16) * 16 if I2 else 0), -(-O // 16) * 16
  H1, W1 = (H // 2, W // 2) if up else (H, W)
  s1 = np.zeros((H1, W1, C1), np.float16); s1[:, :, :I1] = _rand_half(rng, (H1, W1, I1))
  s2 = None
  if I2:
    s2 = np.zeros((H, W, C2), np.float16); s2[:, :, :I2] = _rand_half(rng, (H, W, I2))
  w = (rng.standard_normal((O, I1 + I2, 3, 3)).astype(np.float32) * np.sqrt(2.0 / (9 * (I1 + I2)))).astype(np.float16)
  b = _rand_half(rng, (O,), 0, 0.1)

  op = ConvOp(H, W, C1, C2, Co, relu=1, post_op=post, up=up)
  got = op.run(torch.from_numpy(s1).cuda(), torch.from_numpy(s2).cuda() if I2 else None, w, b, I1, I2).cpu().numpy()

  # oracle: fp32 math on the same fp16-rounded inputs, the reference's op order (upsample, concat, conv, pool)
  x = s1[:, :, :I1].astype(np.float32)
  if up:
    x = np.repeat(np.repeat(x, 2, axis=0), 2, axis=1)
  if I2:
    x = np.concatenate([x, s2[:, :, :I2].astype(np.float32)], axis=2)
  ref = _oracle_conv(oracle, x, w, b, 1)
  if post == 1:
    ref = ref.reshape(H // 2, 2, W // 2, 2, O).max(axis=(1, 3))
  assert not np.isnan(got.astype(np.float32)).any(), "output not fully written"
  # fp16 storage of the result: half an ulp of the value + accumulation-order noise
  np.testing.assert_allclose(got[:, :, :O].astype(np.float32), ref, rtol=2e-3, atol=1e-3)
  assert np.all(got[:, :, O:] == 0), "padded output channels must be zero"


def test_conv_tensor_core_equals_simt_witness():
  rng = np.random.default_rng(5)
  H, W, C1, C2, Co = 40, 300, 64, 16, 64
  s1 = torch.from_numpy(_rand_half(rng, (H // 2, W // 2, C1))).cuda()
  s2 = torch.from_numpy(_rand_half(rng, (H, W, C2))).cuda()
  w = (rng.standard_normal((Co, C1 + C2, 3, 3)) * 0.05).astype(np.float16); b = _rand_half(rng, (Co,), 0, 0.1)
  op = ConvOp(H, W, C1, C2, Co, post_op=0, up=1)
  a = op.run(s1, s2, w, b, C1, C2).float().cpu().numpy()
  r = op.run(s1, s2, w, b, C1, C2, simt=True).float().cpu().numpy()
  np.testing.assert_allclose(a, r, rtol=2e-3, atol=1e-3)


def test_conv_rejects_bad_shapes():
  L = capi.lib()
  h = C.c_void_p()
  for d in (capi.ConvDesc(16, 16, 9, 0, 32, 1, 0, 0, 0), capi.ConvDesc(15, 16, 16, 0, 32, 1, 1, 0, 0),
            capi.ConvDesc(0, 16, 16, 0, 32, 1, 0, 0, 0)):
    assert L.oidnb200_conv_create(C.byref(d), C.byref(h)) == -1
    assert L.oidnb200_last_error()


TF_CASES = [(capi.TF_PU, 1, 0), (capi.TF_SRGB, 0, 0), (capi.TF_LOG, 1, 0), (capi.TF_LINEAR, 0, 1), (capi.TF_LINEAR, 0, 0)]


@pytest.mark.parametrize("tf,hdr,snorm", TF_CASES)
@pytest.mark.parametrize("dtype", ["float32", "float16"])
def test_input_process(tf, hdr, snorm, dtype, oracle):
  W, H, TW, TH = 100, 37, 112, 48
  imgs = synth.benchmark_images(W, H, hdr=bool(hdr), seed=7)
  if snorm:
    imgs["color"] = imgs["color"] * 2 - 1
  imgs["color"][3, 5] = [np.nan, np.inf, -np.inf]; imgs["albedo"][4, 6] = [np.nan, 2.0, -1.0]; imgs["normal"][5, 7] = [np.nan, 3.0, -3.0]
  imgs = {k: v.astype(dtype) for k, v in imgs.items()}
  tile = dict(hSrcBegin=2, wSrcBegin=4, hDstBegin=8, wDstBegin=12, H=30, W=88)
  scale = 0.37 if hdr else 1.0

  ref = np.zeros((TH, TW, 9), np.float32)
  oi = [oracle.image_of(imgs[k]) for k in ("color", "albedo", "normal")]
  ot = oracle.Tile(*[tile[n] for n, _ in oracle.Tile._fields_])
  oracle.lib().oro_input_process(C.byref(oi[0]), C.byref(oi[1]), C.byref(oi[2]), C.byref(ot), tf, hdr, snorm, scale,
                                 ref.ctypes.data, TH, TW, 9)

  t = {k: torch.from_numpy(v).cuda() for k, v in imgs.items()}
  gi = [image_of(t[k]) for k in ("color", "albedo", "normal")]
  gt = capi.Tile(*[tile[n] for n, _ in capi.Tile._fields_])
  gtf = capi.Transfer(tf, scale, None)
  dst = torch.full((TH, TW, 16), float("nan"), dtype=torch.float16, device="cuda")
  check(capi.lib().oidnb200_input_process_launch(C.byref(gi[0]), C.byref(gi[1]), C.byref(gi[2]), C.byref(gt), C.byref(gtf),
                                                 hdr, snorm, dst.data_ptr(), TH, TW, 16, torch.cuda.current_stream().cuda_stream))
  got = dst.float().cpu().numpy()
  assert np.all(got[:, :, 9:] == 0)
  with np.errstate(over="ignore"):
    ref16 = ref.astype(np.float16).astype(np.float32)   # fp16 storage: inf for hdr overflow, like the reference's half tensors
  np.testing.assert_allclose(got[:, :, :9], ref16, rtol=1.5e-3, atol=1e-6)


def test_input_process_vectorised_path_equals_generic(oracle):
  """Packed fp32 RGB + 4-aligned tile -> float4 path; a 4-byte shifted view of the same data -> generic path."""
  W, H = 256, 24
  imgs = synth.benchmark_images(W, H, hdr=True, seed=9)
  tile = capi.Tile(0, 16, 0, 32, 24, 224)
  tf = capi.Transfer(capi.TF_PU, 0.01, None)
  outs = []
  for shift in (0, 1):
    ts = []
    for k in ("color", "albedo", "normal"):
      buf = torch.zeros(H * W * 3 + 4, dtype=torch.float32, device="cuda")
      buf[shift:shift + H * W * 3] = torch.from_numpy(imgs[k].ravel()).cuda()
      ts.append(buf[shift:shift + H * W * 3].view(H, W, 3))
    gi = [image_of(x) for x in ts]
    dst = torch.full((H, 256, 16), float("nan"), dtype=torch.float16, device="cuda")
    check(capi.lib().oidnb200_input_process_launch(C.byref(gi[0]), C.byref(gi[1]), C.byref(gi[2]), C.byref(tile), C.byref(tf),
                                                   1, 0, dst.data_ptr(), H, 256, 16, torch.cuda.current_stream().cuda_stream))
    outs.append(dst.cpu().numpy().view(np.uint16))
  np.testing.assert_array_equal(outs[0], outs[1])


@pytest.mark.parametrize("tf,hdr,snorm", TF_CASES)
@pytest.mark.parametrize("fmt", [("float32", 3), ("float16", 3), ("float32", 1), ("float16", 2)])
def test_output_process(tf, hdr, snorm, fmt, oracle):
  dtype, Cout = fmt
  TH, TW, CT = 48, 128, 16
  rng = np.random.default_rng(3)
  src = np.zeros((TH, TW, CT), np.float16)
  src[:, :, :3] = _rand_half(rng, (TH, TW, 3), -0.1, 1.0)
  src[2, 3, 0] = np.nan; src[2, 4, 1] = np.inf
  tile = dict(hSrcBegin=16, wSrcBegin=32, hDstBegin=5, wDstBegin=8, H=30, W=92)
  H, W = 40, 104
  scale = 0.25 if hdr else 1.0
  ref = np.full((H, W, Cout), -7.0, dtype)
  got_t = torch.full((H, W, Cout), -7.0, dtype=getattr(torch, dtype), device="cuda")
  oi = oracle.image_of(ref)
  ot = oracle.Tile(*[tile[n] for n, _ in oracle.Tile._fields_])
  s32 = src.astype(np.float32)
  oracle.lib().oro_output_process(s32.ctypes.data, TH, TW, CT, C.byref(ot), tf, hdr, snorm, scale, C.byref(oi))
  gi = image_of(got_t)
  gt = capi.Tile(*[tile[n] for n, _ in capi.Tile._fields_])
  gtf = capi.Transfer(tf, scale, None)
  s = torch.from_numpy(src).cuda()
  check(capi.lib().oidnb200_output_process_launch(s.data_ptr(), TH, TW, CT, C.byref(gt), C.byref(gtf), hdr, snorm, C.byref(gi),
                                                  torch.cuda.current_stream().cuda_stream))
  got = got_t.cpu().numpy().astype(np.float32)
  r = ref.astype(np.float32)
  finite = np.isfinite(r)
  assert np.array_equal(finite, np.isfinite(got))
  np.testing.assert_allclose(got[finite], r[finite], rtol=(2e-3 if dtype == "float16" else 2e-5), atol=1e-6)


@pytest.mark.parametrize("tf,hdr,snorm", TF_CASES)
@pytest.mark.parametrize("geom", [(70, 300, dict(hSrcBegin=16, wSrcBegin=32, hDstBegin=5, wDstBegin=9, H=50, W=201)),
                                  (33, 128, dict(hSrcBegin=0, wSrcBegin=0, hDstBegin=0, wDstBegin=0, H=33, W=128)),
                                  (16, 40, dict(hSrcBegin=15, wSrcBegin=39, hDstBegin=2, wDstBegin=1, H=1, W=1))])
def test_conv_fused_output_process_equals_separate_pass(tf, hdr, snorm, geom):
  """The last conv with the output process in its epilogue (oidnb200_conv_set_output_process) writes
  bit-identical pixels to conv + oidnb200_output_process_launch, touches nothing outside the tile's
  rectangle and does not write its tensor."""
  H, W, tile = geom
  rng = np.random.default_rng(11)
  I, O = 32, 3
  src = torch.from_numpy(_rand_half(rng, (H, W, I))).cuda()
  w = (rng.standard_normal((O, I, 3, 3)) * 0.05).astype(np.float16)
  b = np.array([0.1, -0.2, 0.3], np.float16)   # one channel mostly negative before the ReLU
  IH, IW = 64, 256
  scale_t = torch.tensor([0.25 if hdr else 1.0], dtype=torch.float32, device="cuda")
  gt = capi.Tile(*[tile[n] for n, _ in capi.Tile._fields_])
  st = torch.cuda.current_stream().cuda_stream
  for scale_ptr in (None, scale_t.data_ptr()):
    gtf = capi.Transfer(tf, 0.25 if hdr else 1.0, scale_ptr)
    # separate pass
    op = ConvOp(H, W, 32, 0, 16, relu=1)
    tensor = op.run(src, None, w, b, I, 0)
    img_a = torch.full((IH, IW + 3, 3), -7.0, dtype=torch.float32, device="cuda")[:, :IW]   # padded rows
    check(capi.lib().oidnb200_output_process_launch(tensor.data_ptr(), H, W, 16, C.byref(gt), C.byref(gtf), hdr, snorm,
                                                    C.byref(image_of(img_a)), st))
    # fused
    img_b = torch.full((IH, IW + 3, 3), -7.0, dtype=torch.float32, device="cuda")[:, :IW]
    op2 = ConvOp(H, W, 32, 0, 16, relu=1)
    t2 = op2.run(src, None, w, b, I, 0, fused=(gt, gtf, hdr, snorm, image_of(img_b)))
    torch.cuda.synchronize()
    a = img_a.cpu().numpy(); bb = img_b.cpu().numpy()
    np.testing.assert_array_equal(a.view(np.uint32), bb.view(np.uint32))
    inside = np.zeros((IH, IW), bool)
    inside[tile["hDstBegin"]:tile["hDstBegin"] + tile["H"], tile["wDstBegin"]:tile["wDstBegin"] + tile["W"]] = True
    assert np.all(bb[~inside] == -7.0) and np.all(bb[inside] != -7.0)
    assert bool(torch.isnan(t2).all()), "the fused conv must not store its tensor"
    # removing the fusion restores the plain conv
    check(capi.lib().oidnb200_conv_set_output_process(op2.h, None, None, 0, 0, None))
    check(capi.lib().oidnb200_conv_launch(op2.h, st)); torch.cuda.synchronize()
    assert torch.equal(t2, tensor)


def test_conv_fused_output_process_rejects_what_it_cannot_write():
  gt = capi.Tile(0, 0, 0, 0, 8, 8)
  gtf = capi.Transfer(capi.TF_LINEAR, 1.0, None)
  half_img = torch.zeros((8, 8, 3), dtype=torch.float16, device="cuda")
  mono_img = torch.zeros((8, 8, 1), dtype=torch.float32, device="cuda")
  rgb_img = torch.zeros((8, 8, 3), dtype=torch.float32, device="cuda")
  last = ConvOp(8, 8, 32, 0, 16)
  for img in (half_img, mono_img):
    assert capi.lib().oidnb200_conv_set_output_process(last.h, C.byref(gt), C.byref(gtf), 0, 0, C.byref(image_of(img))) == -2
  wide = ConvOp(8, 8, 32, 0, 32)
  assert capi.lib().oidnb200_conv_set_output_process(wide.h, C.byref(gt), C.byref(gtf), 0, 0, C.byref(image_of(rgb_img))) == -2
  big = capi.Tile(0, 0, 0, 0, 9, 8)
  assert capi.lib().oidnb200_conv_set_output_process(last.h, C.byref(big), C.byref(gtf), 0, 0, C.byref(image_of(rgb_img))) == -1


@pytest.mark.parametrize("W,H", [(37, 21), (64, 48), (130, 70), (1920, 1080), (5, 3)])
@pytest.mark.parametrize("dtype", ["float32", "float16"])
def test_autoexposure(W, H, dtype, oracle):
  img = synth.benchmark_images(W, H, hdr=True, albedo=False, normal=False, seed=11)["color"]
  img[0, 0] = [np.nan, -5, np.inf] if dtype == "float32" else [np.nan, -5, 1.0]
  img = img.astype(dtype)
  ref = oracle.autoexposure(img)
  L = capi.lib()
  t = torch.from_numpy(img).cuda()
  scratch = torch.empty(L.oidnb200_autoexposure_scratch_bytes(H, W), dtype=torch.uint8, device="cuda")
  dst = torch.zeros(1, dtype=torch.float32, device="cuda")
  gi = image_of(t)
  for _ in range(2):  # relaunch on the same scratch
    check(L.oidnb200_autoexposure_launch(C.byref(gi), scratch.data_ptr(), dst.data_ptr(), torch.cuda.current_stream().cuda_stream))
  got = float(dst.cpu()[0])
  assert np.isfinite(got) and abs(got - ref) <= 2e-5 * abs(ref), (got, ref)


def test_autoexposure_black_image_gives_one():
  t = torch.zeros((33, 47, 3), dtype=torch.float32, device="cuda")
  L = capi.lib()
  scratch = torch.empty(L.oidnb200_autoexposure_scratch_bytes(33, 47), dtype=torch.uint8, device="cuda")
  dst = torch.zeros(1, dtype=torch.float32, device="cuda")
  gi = image_of(t)
  check(L.oidnb200_autoexposure_launch(C.byref(gi), scratch.data_ptr(), dst.data_ptr(), torch.cuda.current_stream().cuda_stream))
  assert float(dst.cpu()[0]) == 1.0


@pytest.mark.parametrize("W,H", [(130, 70), (1000, 612)])
def test_autoexposure_bins_by_rectangle_equal_whole_image(W, H):
  """The multi-GPU decomposition: bins computed rectangle by rectangle into separate zero-filled arrays
  and summed give bit for bit the array (and the scale) of the single launch. Pixels outside a
  rectangle's bins are poisoned with NaN-free garbage to show they are not read."""
  img = synth.benchmark_images(W, H, hdr=True, albedo=False, normal=False, seed=12)["color"]
  img[H // 2 - 24:H // 2 + 24, : W // 2] = 0.0       # >= 1 whole bin below the 1e-8 threshold (-inf entries)
  L = capi.lib()
  st = torch.cuda.current_stream().cuda_stream
  nbh, nbw = C.c_int(), C.c_int()
  L.oidnb200_autoexposure_bin_grid(H, W, C.byref(nbh), C.byref(nbw))
  nbh, nbw = nbh.value, nbw.value
  assert (nbh, nbw) == ((H + 15) // 16, (W + 15) // 16)
  t = torch.from_numpy(img).cuda()
  whole = torch.zeros(nbh * nbw, dtype=torch.float32, device="cuda")
  ref_scale = torch.zeros(1, device="cuda")
  check(L.oidnb200_autoexposure_bins_launch(C.byref(image_of(t)), 0, nbh, 0, nbw, whole.data_ptr(), st))
  check(L.oidnb200_autoexposure_reduce_launch(whole.data_ptr(), nbh * nbw, ref_scale.data_ptr(), st))
  one = torch.zeros(1, device="cuda")
  scratch = torch.empty(L.oidnb200_autoexposure_scratch_bytes(H, W), dtype=torch.uint8, device="cuda")
  check(L.oidnb200_autoexposure_launch(C.byref(image_of(t)), scratch.data_ptr(), one.data_ptr(), st))
  assert one.cpu().numpy().view(np.uint32)[0] == ref_scale.cpu().numpy().view(np.uint32)[0]
  total = torch.zeros_like(whole)
  hs, ws = [0, nbh // 3, nbh], [0, nbw // 2, nbw]
  for i in range(2):
    for j in range(2):
      # this "rank" holds only the pixels of its bins; everything else is garbage
      h0, h1 = hs[i] * H // nbh, hs[i + 1] * H // nbh
      w0, w1 = ws[j] * W // nbw, ws[j + 1] * W // nbw
      local = torch.full_like(t, 1e30)
      local[h0:h1, w0:w1] = t[h0:h1, w0:w1]
      part = torch.zeros_like(whole)
      check(L.oidnb200_autoexposure_bins_launch(C.byref(image_of(local)), hs[i], hs[i + 1], ws[j], ws[j + 1], part.data_ptr(), st))
      total += part
  assert np.array_equal(total.cpu().numpy().view(np.uint32), whole.cpu().numpy().view(np.uint32))
  assert np.isneginf(whole.cpu().numpy()).any()
  got = torch.zeros(1, device="cuda")
  check(L.oidnb200_autoexposure_reduce_launch(total.data_ptr(), nbh * nbw, got.data_ptr(), st))
  assert got.cpu().numpy().view(np.uint32)[0] == ref_scale.cpu().numpy().view(np.uint32)[0]
  with pytest.raises(RuntimeError):
    check(L.oidnb200_autoexposure_bins_launch(C.byref(image_of(t)), 0, nbh + 1, 0, nbw, whole.data_ptr(), st))


def test_pool_upsample_image_copy(oracle):
  rng = np.random.default_rng(1)
  H, W, Cc = 20, 36, 48
  x = _rand_half(rng, (H, W, Cc), -1, 1)
  t = torch.from_numpy(x).cuda()
  L = capi.lib(); st = torch.cuda.current_stream().cuda_stream
  p = torch.empty((H // 2, W // 2, Cc), dtype=torch.float16, device="cuda")
  check(L.oidnb200_pool_launch(t.data_ptr(), H, W, Cc, p.data_ptr(), st))
  np.testing.assert_array_equal(p.cpu().numpy(), x.reshape(H // 2, 2, W // 2, 2, Cc).max(axis=(1, 3)))
  u = torch.empty((H * 2, W * 2, Cc), dtype=torch.float16, device="cuda")
  check(L.oidnb200_upsample_launch(t.data_ptr(), H, W, Cc, u.data_ptr(), st))
  np.testing.assert_array_equal(u.cpu().numpy(), np.repeat(np.repeat(x, 2, 0), 2, 1))
  for dtype in (torch.float32, torch.float16):
    src = torch.rand((17, 29, 3), device="cuda").to(dtype)
    big = torch.zeros((17, 40, 4), dtype=dtype, device="cuda")
    dstv = big[:, 3:32, :3]
    a, b = image_of(src), image_of(dstv)
    check(L.oidnb200_image_copy_launch(C.byref(a), C.byref(b), st))
    assert torch.equal(dstv, src) and float(big[:, :3].abs().sum()) == 0 and float(big[:, :, 3].abs().sum()) == 0


# ---- conv pairs: B(A(x)) as one launch (kernels/conv_pair_tc.cu) ----------------------------------------------------
# (H, W, I_A, C_A(out of A = in of B), O_B, poolB)
PAIR_CASES = [
  (16, 128, 9, 32, 32, 1),      # enc_conv0 -> enc_conv1 + pool
  (38, 301 * 2, 9, 32, 32, 1),  # several strips of 126, ragged last strip
  (2, 2, 9, 32, 32, 1),         # smallest poolable tile
  (1, 1, 64, 32, 3, 0),
  (33, 127, 64, 32, 3, 0),      # dec_conv1b -> dec_conv0: one pixel more than a strip
  (40, 253, 32, 32, 3, 0),      # small net: dec_conv1b (32) -> dec_conv0
  (24, 130, 9, 64, 64, 1),      # large net: enc_conv1a -> enc_conv1b + pool (one stream)
  (21, 260, 64, 64, 3, 0),      # large net: dec_conv1b -> dec_conv1c
  (300, 140, 64, 32, 3, 0),     # many row chunks
]


@pytest.mark.parametrize("case", PAIR_CASES, ids=["%dx%d_%d_%d_%d_p%d" % c for c in PAIR_CASES])
def test_conv_pair_bit_identical_to_two_launches(case):
  """The fused pair keeps conv A's rows in shared memory (rounded to fp16 exactly as the stored tensor would have
  been, zero outside the image: conv B's padding) -- every output bit must equal the two-launch path, which
  test_conv_matches_oracle pins to the oracle."""
  H, W, IA, CA, OB, pool = case
  L = capi.lib()
  rng = np.random.default_rng(hash(case) & 0xFFFF)
  C1, Cb = -(-IA // 16) * 16, -(-OB // 16) * 16
  s = np.zeros((H, W, C1), np.float16); s[:, :, :IA] = _rand_half(rng, (H, W, IA))
  wa = (rng.standard_normal((CA, IA, 3, 3)) * np.sqrt(2.0 / (9 * IA))).astype(np.float16)
  ba = (rng.random(CA) * 0.2 - 0.1).astype(np.float16)        # some channels go negative before the ReLU
  wb = (rng.standard_normal((OB, CA, 3, 3)) * np.sqrt(2.0 / (9 * CA))).astype(np.float16)
  bb = (rng.random(OB) * 0.2 - 0.1).astype(np.float16)
  src = torch.from_numpy(s).cuda()
  a = ConvOp(H, W, C1, 0, CA, relu=1)
  mid = a.run(src, None, wa, ba, IA, 0)
  b = ConvOp(H, W, CA, 0, Cb, relu=1, post_op=pool)
  ref = b.run(mid, None, wb, bb, CA, 0)
  # rebind B to a fresh destination and poison the tensor between the convs: the pair must not touch either
  out = torch.full_like(ref, float("nan"))
  dwb, dbb = b._keep
  check(L.oidnb200_conv_bind(b.h, mid.data_ptr(), None, dwb.data_ptr(), dbb.data_ptr(), out.data_ptr()))
  mid.fill_(float("nan"))
  pair = C.c_void_p()
  check(L.oidnb200_conv_pair_create(a.h, b.h, C.byref(pair)))
  check(L.oidnb200_conv_pair_bind(pair))
  check(L.oidnb200_conv_pair_launch(pair, torch.cuda.current_stream().cuda_stream))
  torch.cuda.synchronize()
  L.oidnb200_conv_pair_destroy(pair)
  assert bool(torch.isnan(mid).all())
  assert torch.equal(out.view(torch.int16), ref.view(torch.int16))


TAP_CASES = [(33, 127, 64, 32), (40, 253, 32, 32), (21, 260, 64, 64), (300, 140, 64, 32), (1, 1, 64, 32), (3, 126, 64, 32),
             (7, 126 * 3 + 1, 64, 32), (2, 64, 32, 32)]


@pytest.mark.parametrize("case", TAP_CASES, ids=["%dx%d_%d_%d" % c for c in TAP_CASES])
@pytest.mark.parametrize("tf,hdr", [(capi.TF_LINEAR, 0), (capi.TF_PU, 1)])
def test_conv_pair_tap_packed_last_conv(case, tf, hdr, monkeypatch):
  """The 3-channel last conv of a pair with the output process fused takes its horizontal taps as accumulator COLUMNS
  (one MMA view instead of three) and its epilogue adds the three partial sums of neighbouring pixels, across lanes and
  across the warps of a strip. The fp32 sums associate differently from the single accumulator of the two-launch
  path, so the fp16-rounded network output may differ by one fp16 step in a few pixels -- nothing more; with
  OIDN_B200_NO_TAP_PACK the pair is bit-identical to two launches. Nothing outside the tile's rectangle is written."""
  H, W, IA, CA = case
  L = capi.lib()
  rng = np.random.default_rng(hash(case) & 0xFFFF)
  src = torch.from_numpy(_rand_half(rng, (H, W, IA))).cuda()
  wa = (rng.standard_normal((CA, IA, 3, 3)) * np.sqrt(2.0 / (9 * IA))).astype(np.float16)
  ba = (rng.random(CA) * 0.2 - 0.1).astype(np.float16)
  wb = (rng.standard_normal((3, CA, 3, 3)) * np.sqrt(1.0 / (9 * CA))).astype(np.float16)
  bb = np.array([0.4, 0.5, 0.05], np.float16)
  tile = dict(hSrcBegin=0, wSrcBegin=0, hDstBegin=2, wDstBegin=3, H=H, W=W)
  if H > 20:
    tile = dict(hSrcBegin=3, wSrcBegin=5, hDstBegin=2, wDstBegin=3, H=H - 7, W=W - 9)
  gt = capi.Tile(*[tile[n] for n, _ in capi.Tile._fields_])
  gtf = capi.Transfer(tf, 0.25 if hdr else 1.0, None)
  IH, IW = H + 4, W + 6
  st = torch.cuda.current_stream().cuda_stream
  a = ConvOp(H, W, IA, 0, CA, relu=1)
  mid = a.run(src, None, wa, ba, IA, 0)
  img_ref = torch.full((IH, IW, 3), -7.0, dtype=torch.float32, device="cuda")
  b = ConvOp(H, W, CA, 0, 16, relu=0)
  b.run(mid, None, wb, bb, CA, 0, fused=(gt, gtf, hdr, 0, image_of(img_ref)))
  imgs = {}
  for mode in ("tap", "regular"):
    if mode == "regular":
      monkeypatch.setenv("OIDN_B200_NO_TAP_PACK", "1")
    img = torch.full((IH, IW, 3), -7.0, dtype=torch.float32, device="cuda")
    check(L.oidnb200_conv_set_output_process(b.h, C.byref(gt), C.byref(gtf), hdr, 0, C.byref(image_of(img))))
    pair = C.c_void_p()
    check(L.oidnb200_conv_pair_create(a.h, b.h, C.byref(pair)))
    check(L.oidnb200_conv_pair_bind(pair))
    check(L.oidnb200_conv_pair_launch(pair, st))
    torch.cuda.synchronize()
    L.oidnb200_conv_pair_destroy(pair)
    imgs[mode] = img.cpu().numpy()
  ref = img_ref.cpu().numpy()
  np.testing.assert_array_equal(imgs["regular"].view(np.uint32), ref.view(np.uint32))
  inside = np.zeros((IH, IW), bool)
  inside[tile["hDstBegin"]:tile["hDstBegin"] + tile["H"], tile["wDstBegin"]:tile["wDstBegin"] + tile["W"]] = True
  got = imgs["tap"]
  assert np.all(got[~inside] == -7.0) and np.all(got[inside] != -7.0)
  if tf == capi.TF_LINEAR:
    # the image is the fp16 network output itself (clamped to [0, 1]): at most one fp16 step apart, or, where the taps
    # cancel to almost nothing, the fp32 rounding of the partial sums (~0.5 each)
    step = np.maximum(np.abs(ref[inside]), 2.0 ** -14) * 2.0 ** -10
    assert np.all(np.abs(got[inside] - ref[inside]) <= step * 1.001 + 1e-6)
  else:
    np.testing.assert_allclose(got[inside], ref[inside], rtol=5e-2, atol=1e-4)   # PU inverse of a one-step difference (3 % near y = 1)
  assert np.mean(got[inside] != ref[inside]) < 0.05


def test_conv_pair_rejects_what_it_does_not_cover():
  L = capi.lib()
  def pair_rc(da, db):
    a, b = ConvOp(*da), ConvOp(*db)
    p = C.c_void_p()
    rc = L.oidnb200_conv_pair_create(a.h, b.h, C.byref(p))
    if rc == 0:
      L.oidnb200_conv_pair_destroy(p)
    return rc
  assert pair_rc((32, 128, 16, 0, 32), (32, 128, 32, 0, 32)) == 0
  assert pair_rc((32, 128, 16, 0, 32), (32, 128, 32, 16, 32)) == -2          # concat B
  assert pair_rc((32, 128, 96, 0, 96), (32, 128, 96, 0, 96)) == -2           # more than one K chunk / 96 channels
  assert pair_rc((32, 128, 16, 0, 32), (16, 64, 32, 0, 32)) == -2            # different resolution
  assert pair_rc((32, 128, 16, 0, 48), (32, 128, 48, 0, 32)) == -2           # A's output must be 32 or 64 channels
