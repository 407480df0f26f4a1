"""-m gpu: the RT / RTLightmap filters through the filter-level C ABI against the CPU oracle and the
committed golden vectors (reference PyTorch), plus the reference's own behavioural tests
(apps/oidnTest.cpp): tiled == untiled bit-exactly, in-place, sanitisation, formats, errors.

Tolerance (BASELINE.json north_star): max|out - ref| <= 1e-2 * max|ref| and PSNR >= 50 dB against
the fp32 CPU result."""
import numpy as np
import pytest

from oidn_b200 import api, capi, synth, weights

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from gpu_util import metrics  # noqa: E402
from test_oracle_golden import CASES, case_inputs  # noqa: E402

MAX_ERR, MIN_PSNR = 1e-2, 50.0


@pytest.fixture(scope="module")
def device():
  d = api.Device((0,)).commit()
  yield d
  d.release()


def run_filter(device, tza, color=None, albedo=None, normal=None, filt="RT", out_dtype=None, out_channels=None, **params):
  """Runs the CUDA filter on numpy HxWxC inputs; returns (numpy output, filter info)."""
  f = device.new_filter(filt)
  keep = []
  main = color if color is not None else (albedo if albedo is not None else normal)
  for name, img in (("color", color), ("albedo", albedo), ("normal", normal)):
    if img is not None:
      t = torch.from_numpy(np.ascontiguousarray(img)).cuda(); keep.append(t)
      f.set_image(name, t)
  H, W = main.shape[:2]
  Cc = out_channels or (main.shape[2] if main.ndim == 3 else 1)
  out = torch.full((H, W, Cc), -123.0, dtype=getattr(torch, out_dtype or main.dtype.name), device="cuda")
  f.set_image("output", out)
  max_tile = params.pop("maxTilePixels", None)
  if max_tile:
    device.set("maxTilePixels", max_tile)
  for k, v in params.items():
    f.set(k, v)
  f.set_data("weights", tza)
  try:
    f.commit()
    f.execute()
    info = f.info()
  finally:
    if max_tile:
      device.set("maxTilePixels", 3840 * 2176)
  res = out.cpu().numpy()
  f.release()
  return res, info


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_matches_reference_golden_and_oracle(case, golden, oracle, device):
  name, kind, ic, filt, mode, W, H = case
  tza, color, albedo, normal = case_inputs(kind, ic, mode, W, H)
  params = {}
  if filt == "RT":
    params = dict(hdr=(mode == "hdr"), srgb=(mode == "srgb"))
  elif mode == "dir":
    params = dict(directional=True)
  got, info = run_filter(device, tza, color, albedo, normal, filt, **params)
  assert info["largeModel"] == int(kind == "large")
  ref_torch = golden[name + "/output"]               # reference PyTorch fp32
  ref_oracle = np.zeros((H, W, 3), np.float32)       # C restatement of the reference CPU device
  oracle.filter_execute(tza, color=color, albedo=albedo, normal=normal, output=ref_oracle, filter=filt,
                        hdr=(mode == "hdr"), srgb=(mode == "srgb"), directional=(mode == "dir"))
  for ref in (ref_torch, ref_oracle):
    if mode == "dir":   # signed output: measure on the [0,1] mapped range like the other cases
      e, p = metrics(got * 0.5 + 0.5, ref * 0.5 + 0.5)
    else:
      e, p = metrics(got, ref)
    assert e <= MAX_ERR and p >= MIN_PSNR, (name, e, p)


@pytest.mark.parametrize("W,H", [(1, 1), (2, 2), (17, 5), (257, 89), (640, 368)])
def test_sizes_hdr_alb_nrm(W, H, oracle, device):
  tza = weights.model_tza("base", 9, seed=0)
  imgs = synth.benchmark_images(W, H, hdr=True, seed=2)
  got, _ = run_filter(device, tza, imgs["color"], imgs["albedo"], imgs["normal"], hdr=True)
  ref = np.zeros((H, W, 3), np.float32)
  oracle.filter_execute(tza, color=imgs["color"], albedo=imgs["albedo"], normal=imgs["normal"], output=ref, hdr=True)
  e, p = metrics(got, ref)
  assert e <= MAX_ERR and p >= MIN_PSNR, (e, p)


def test_zero_size_image_is_noop(device):
  """core/unet_filter.cpp:150-151: a 0x0 image commits and executes as a no-op."""
  f = device.new_filter("RT")
  keep = torch.zeros(16, device="cuda")
  f.set_image("color", keep.data_ptr(), capi.FORMAT_FLOAT3, 0, 0)
  f.set_image("output", keep.data_ptr(), capi.FORMAT_FLOAT3, 0, 0)
  f.set_data("weights", weights.model_tza("base", 3, seed=0))
  f.commit(); f.execute()
  f.release()


def test_tiled_equals_untiled_bit_exact_and_inplace(device, oracle):
  """apps/oidnTest.cpp:717-775 ("inplace filter"): forced multi-tile in-place == normal run."""
  W, H = 1500, 900
  tza = weights.model_tza("base", 9, seed=0)
  imgs = synth.benchmark_images(W, H, hdr=True, seed=4)
  whole, i1 = run_filter(device, tza, imgs["color"], imgs["albedo"], imgs["normal"], hdr=True)
  assert i1["tileCountH"] * i1["tileCountW"] == 1
  tiled, i2 = run_filter(device, tza, imgs["color"], imgs["albedo"], imgs["normal"], hdr=True, maxMemoryMB=0)
  assert i2["tileCountH"] * i2["tileCountW"] > 1
  np.testing.assert_array_equal(whole.view(np.uint32), tiled.view(np.uint32))

  # in-place (output aliases color) + forced tiling goes through the temporary image + ImageCopy
  f = device.new_filter("RT")
  c = torch.from_numpy(imgs["color"]).cuda(); a = torch.from_numpy(imgs["albedo"]).cuda(); n = torch.from_numpy(imgs["normal"]).cuda()
  f.set_image("color", c); f.set_image("albedo", a); f.set_image("normal", n); f.set_image("output", c)
  f.set("hdr", True); f.set("maxMemoryMB", 0); f.set_data("weights", tza)
  f.commit(); f.execute()
  np.testing.assert_array_equal(c.cpu().numpy().view(np.uint32), whole.view(np.uint32))
  f.release()


@pytest.mark.parametrize("filt,mode", [("RT", "hdr"), ("RT", "ldr"), ("RTLightmap", "hdr"), ("RTLightmap", "dir")])
def test_fused_output_process_equals_separate_pass(filt, mode, device):
  """Device parameter fuseOutput: the output process inside dec_conv0's epilogue (default) against
  the reference's separate pass, bit for bit, single tile and forced multi-tile."""
  W, H = 1500, 900
  ic = 9 if (filt, mode) == ("RT", "hdr") else 3
  tza = weights.model_tza("base", ic, seed=0)
  imgs = synth.benchmark_images(W, H, hdr=(mode == "hdr"), seed=6)
  color = imgs["color"] * 2 - 1 if mode == "dir" else imgs["color"]
  kw = dict(albedo=imgs["albedo"], normal=imgs["normal"]) if ic == 9 else {}
  params = dict(hdr=(mode == "hdr")) if filt == "RT" else (dict(directional=True) if mode == "dir" else {})
  res = {}
  try:
    device.set("fusePairs", 0)   # inside a fused pair the last conv is tap-packed: test_fused_conv_pairs_equal_separate_launches
    for fuse in (1, 0):
      device.set("fuseOutput", fuse)
      for tiled in (False, True):
        extra = dict(maxMemoryMB=0) if tiled else {}
        res[fuse, tiled], info = run_filter(device, tza, color, filt=filt, **kw, **params, **extra)
        assert (info["tileCountH"] * info["tileCountW"] > 1) == tiled
  finally:
    device.set("fuseOutput", 1)
    device.set("fusePairs", 1)
  assert np.isfinite(res[1, False]).all() and not np.any(res[1, False] == -123.0)
  for k in ((1, True), (0, False), (0, True)):
    np.testing.assert_array_equal(res[1, False].view(np.uint32), res[k].view(np.uint32))


def test_large_model_tiles(device, oracle):
  W, H = 1200, 1000
  tza = weights.model_tza("large", 9, seed=0)
  imgs = synth.benchmark_images(W, H, hdr=True, seed=5)
  whole, i1 = run_filter(device, tza, imgs["color"], imgs["albedo"], imgs["normal"], hdr=True, cleanAux=True)
  tiled, i2 = run_filter(device, tza, imgs["color"], imgs["albedo"], imgs["normal"], hdr=True, cleanAux=True, maxMemoryMB=0)
  assert i1["largeModel"] == 1 and i2["tileOverlap"] == 112 and i2["tileCountH"] * i2["tileCountW"] > 1
  np.testing.assert_array_equal(whole.view(np.uint32), tiled.view(np.uint32))


def test_image_sanitization(device):
  """apps/oidnTest.cpp:990-1034: NaN/Inf/negative inputs give finite outputs in range."""
  W, H = 200, 120
  bad = [np.nan, np.inf, -np.inf, -100.0, 1e30]
  for hdr, kind, ic in ((True, "base", 9), (False, "base", 9)):
    tza = weights.model_tza(kind, ic, seed=0)
    for v in bad:
      color = np.full((H, W, 3), v, np.float32); alb = np.full((H, W, 3), v, np.float32); nrm = np.full((H, W, 3), v, np.float32)
      got, _ = run_filter(device, tza, color, alb, nrm, hdr=hdr)
      assert np.isfinite(got).all() and got.min() >= 0.0
      if not hdr:
        assert got.max() <= 1.0


@pytest.mark.parametrize("in_dtype,out_dtype", [("float16", "float16"), ("float32", "float16"), ("float16", "float32")])
def test_half_images(in_dtype, out_dtype, device, oracle):
  W, H = 300, 200
  tza = weights.model_tza("base", 9, seed=0)
  imgs = synth.benchmark_images(W, H, hdr=True, seed=6)
  imgs["color"] *= np.float32(0.02)   # keeps the denoised values inside the fp16 range
  imgs = {k: v.astype(in_dtype) for k, v in imgs.items()}
  got, _ = run_filter(device, tza, imgs["color"], imgs["albedo"], imgs["normal"], hdr=True, out_dtype=out_dtype)
  ref = np.zeros((H, W, 3), out_dtype)
  oracle.filter_execute(tza, color=imgs["color"], albedo=imgs["albedo"], normal=imgs["normal"], output=ref, hdr=True)
  got = got.astype(np.float32); ref = ref.astype(np.float32)
  # a half output overflows to inf where the denoised HDR value exceeds 65504 -- in both or neither
  fin = np.isfinite(ref) & np.isfinite(got)
  assert fin.mean() > 0.98 and np.mean(np.isfinite(ref) != np.isfinite(got)) < 2e-3
  e, p = metrics(got[fin], ref[fin])
  assert e <= MAX_ERR and p >= MIN_PSNR, (e, p)


def test_strided_and_single_channel_images(device, oracle):
  """Row/pixel strides (core/image.h) and C<3 broadcast (core/image_accessor.h:30-94)."""
  W, H = 180, 100
  tza = weights.model_tza("base", 3, seed=0)
  rng = np.random.default_rng(8)
  big = rng.random((H, W + 7, 4), dtype=np.float32)
  tb = torch.from_numpy(big).cuda()
  view = tb[:, 3:3 + W, :3]                      # pixel stride 16 B, row stride (W+7)*16 B
  outb = torch.zeros((H, W + 5, 4), dtype=torch.float32, device="cuda")
  outv = outb[:, 2:2 + W, :3]
  f = device.new_filter("RT")
  f.set_image("color", view); f.set_image("output", outv); f.set_data("weights", tza)
  f.commit(); f.execute()
  got = outv.cpu().numpy()
  color = np.ascontiguousarray(big[:, 3:3 + W, :3])
  ref = np.zeros((H, W, 3), np.float32)
  oracle.filter_execute(tza, color=color, output=ref)
  e, p = metrics(got, ref)
  assert e <= MAX_ERR and p >= MIN_PSNR, (e, p)
  assert float(outb[:, :2].abs().sum()) == 0 and float(outb[:, :, 3].abs().sum()) == 0, "wrote outside the image"
  f.release()
  # 1-channel in/out
  c1 = color[:, :, :1].copy()
  got1, _ = run_filter(device, tza, c1)
  ref1 = np.zeros((H, W, 1), np.float32)
  oracle.filter_execute(tza, color=c1, output=ref1)
  e, p = metrics(got1, ref1)
  assert e <= MAX_ERR and p >= MIN_PSNR, (e, p)


def test_input_scale_and_filter_updates(device, oracle):
  """apps/oidnTest.cpp:779-869: pointer-only changes do not need a rebuild; param changes do."""
  W, H = 240, 136
  tza = weights.model_tza("base", 9, seed=0)
  imgs = synth.benchmark_images(W, H, hdr=True, seed=3)
  f = device.new_filter("RT")
  t = {k: torch.from_numpy(v).cuda() for k, v in imgs.items()}
  out = torch.zeros((H, W, 3), device="cuda")
  for k, v in t.items():
    f.set_image(k, v)
  f.set_image("output", out); f.set("hdr", True); f.set("inputScale", 0.02); f.set_data("weights", tza)
  f.commit(); f.execute()
  ref = np.zeros((H, W, 3), np.float32)
  oracle.filter_execute(tza, color=imgs["color"], albedo=imgs["albedo"], normal=imgs["normal"], output=ref, hdr=True, input_scale=0.02)
  e, p = metrics(out.cpu().numpy(), ref)
  assert e <= MAX_ERR and p >= MIN_PSNR, (e, p)
  # uncommitted change -> InvalidOperation (core/unet_filter.cpp:147-148)
  out2 = torch.zeros((H, W, 3), device="cuda")
  f.set_image("output", out2)
  with pytest.raises(api.Error) as ei:
    f.execute()
  assert ei.value.code == capi.ERROR_INVALID_OPERATION
  f.commit(); f.execute()
  assert torch.equal(out, out2)
  f.release()


def test_error_behaviour(device):
  tza = weights.model_tza("base", 3, seed=0)
  f = device.new_filter("RT")
  with pytest.raises(api.Error) as ei:
    f.commit()                                   # no images
  assert ei.value.code == capi.ERROR_INVALID_OPERATION
  c = torch.zeros((32, 32, 3), device="cuda"); o = torch.zeros((32, 48, 3), device="cuda")
  f.set_image("color", c); f.set_image("output", o); f.set_data("weights", tza)
  with pytest.raises(api.Error) as ei:
    f.commit()                                   # size mismatch
  assert ei.value.code == capi.ERROR_INVALID_OPERATION
  f.set_image("output", torch.zeros((32, 32, 3), device="cuda"))
  f.set("hdr", True); f.set("srgb", True)
  with pytest.raises(api.Error):
    f.commit()
  f.set("srgb", False)
  f.set_data("weights", b"\x00" * 64)            # corrupted blob (apps/oidnTest.cpp:1169-1217)
  with pytest.raises(api.Error) as ei:
    f.commit()
  assert ei.value.code == capi.ERROR_INVALID_OPERATION
  with pytest.raises(api.Error) as ei:
    f.set("quality", 3)
  assert ei.value.code == capi.ERROR_INVALID_ARGUMENT
  with pytest.raises(api.Error) as ei:
    device.new_filter("Nope")
  assert ei.value.code == capi.ERROR_INVALID_ARGUMENT
  host = np.zeros((32, 32, 3), np.float32)        # pageable host memory is not device accessible
  if not device.get("systemMemorySupported"):
    with pytest.raises(api.Error) as ei:
      f.set_image("color", host.ctypes.data, capi.FORMAT_FLOAT3, 32, 32)
    assert ei.value.code == capi.ERROR_INVALID_ARGUMENT
  f.release()


def test_progress_monitor_and_cancel(device):
  W, H = 400, 300
  tza = weights.model_tza("base", 3, seed=0)
  f = device.new_filter("RT")
  c = torch.rand((H, W, 3), device="cuda"); o = torch.zeros((H, W, 3), device="cuda")
  f.set_image("color", c); f.set_image("output", o); f.set_data("weights", tza)
  seen = []
  f.set_progress_monitor(lambda n: (seen.append(n), True)[1])
  f.commit(); f.execute()
  assert seen[0] == 0.0 and seen[-1] == 1.0 and all(b >= a for a, b in zip(seen, seen[1:]))
  f.set_progress_monitor(lambda n: False)
  with pytest.raises(api.Error) as ei:
    f.execute()
  assert ei.value.code == capi.ERROR_CANCELLED
  f.release()


def test_buffers_and_async(device, oracle):
  """oidnNewBuffer / oidnWriteBuffer / oidnSetFilterImage / oidnExecuteFilterAsync + oidnSyncDevice."""
  W, H = 128, 80
  tza = weights.model_tza("small", 3, seed=0)
  color = synth.benchmark_images(W, H, hdr=False, albedo=False, normal=False, seed=12)["color"]
  nbytes = color.nbytes
  bc = device.new_buffer(nbytes); bo = device.new_buffer(nbytes)
  bc.write(color)
  f = device.new_filter("RT")
  f.set_image("color", bc, capi.FORMAT_FLOAT3, W, H); f.set_image("output", bo, capi.FORMAT_FLOAT3, W, H)
  f.set("quality", api.QUALITY_FAST); f.set_data("weights", tza)
  f.commit()
  for _ in range(3):
    f.execute_async()
  device.sync()
  got = np.zeros_like(color); bo.read(got)
  ref = np.zeros_like(color)
  oracle.filter_execute(tza, color=color, output=ref)
  e, p = metrics(got, ref)
  assert e <= MAX_ERR and p >= MIN_PSNR, (e, p)
  with pytest.raises(api.Error) as ei:
    bc.write(np.zeros(nbytes + 4, np.uint8))
  assert ei.value.code == capi.ERROR_INVALID_ARGUMENT
  f.release(); bc.release(); bo.release()


def test_pinned_host_images_zero_copy(device, oracle):
  """Storage::Host images (oidn.h OIDN_STORAGE_HOST): kernels dereference pinned memory directly."""
  W, H = 160, 96
  tza = weights.model_tza("base", 3, seed=0)
  color = synth.benchmark_images(W, H, hdr=False, albedo=False, normal=False, seed=13)["color"]
  hc = torch.from_numpy(color).pin_memory(); ho = torch.zeros((H, W, 3)).pin_memory()
  f = device.new_filter("RT")
  f.set_image("color", hc); f.set_image("output", ho); f.set_data("weights", tza)
  f.commit(); f.execute()
  ref = np.zeros_like(color)
  oracle.filter_execute(tza, color=color, output=ref)
  e, p = metrics(ho.numpy(), ref)
  assert e <= MAX_ERR and p >= MIN_PSNR, (e, p)
  f.release()


def test_frame_graph_replay_bit_identical():
  """Device parameter "graph": frames replayed as one CUDA graph equal the eagerly launched frames bit
  for bit, new image contents are picked up by the replay, and swapping image pointers (what a
  renderer's double buffering does, core/filter.cpp:52-56: no re-commit) re-captures."""
  W, H = 1700, 820
  tza = weights.model_tza("base", 9, seed=0)
  frames = [synth.benchmark_images(W, H, hdr=True, seed=30 + i) for i in range(2)]
  outs = {}
  for graph in (0, 1):
    dev = api.Device((0,)).commit()
    dev.set("graph", graph)
    dev.set("maxTilePixels", 1000 * 1000)  # two tiles per frame
    t = [{k: torch.from_numpy(v).cuda() for k, v in fr.items()} for fr in frames]
    o = [torch.zeros((H, W, 3), device="cuda") for _ in range(2)]
    f = dev.new_filter("RT")
    f.set("hdr", True); f.set_data("weights", tza)
    res = []
    for it in range(6):                    # A A A B B A: eager, capture, replay, eager(B), capture(B), eager(A) ...
      s = 0 if it in (0, 1, 2, 5) else 1
      for k, v in t[s].items():
        f.set_image(k, v)
      f.set_image("output", o[s])
      f.commit()
      if it == 2:                          # same pointers, new pixel values: the replay must read them
        t[0]["color"].mul_(0.5)
      f.execute_async()
      dev.sync()
      res.append(o[s].cpu().numpy().copy())
    assert f.info()["tileCountH"] * f.info()["tileCountW"] > 1
    outs[graph] = res
    f.release(); dev.release()
  for a, b in zip(outs[0], outs[1]):
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
  assert not np.array_equal(outs[1][1], outs[1][2])


def test_prefilter_pipeline_clean_aux(device, oracle):
  """The cleanAux workflow of the reference README ("Denoising with prefiltering", the aux models of
  core/unet_filter.cpp:417-436): albedo and normal are denoised in place by their own filters, then
  feed the beauty filter with cleanAux=true -- three filters on one device, chained through device
  memory, enqueued asynchronously with a single sync at the end."""
  W, H = 320, 208
  tza_aux = weights.model_tza("base", 3, seed=3)
  tza_beauty = weights.model_tza("large", 9, seed=0)
  imgs = synth.benchmark_images(W, H, hdr=True, seed=21)
  t = {k: torch.from_numpy(v).cuda() for k, v in imgs.items()}
  out = torch.zeros((H, W, 3), device="cuda")
  fa = device.new_filter("RT"); fa.set_image("albedo", t["albedo"]); fa.set_image("output", t["albedo"])
  fa.set_data("weights", tza_aux); fa.commit()
  fn = device.new_filter("RT"); fn.set_image("normal", t["normal"]); fn.set_image("output", t["normal"])
  fn.set_data("weights", tza_aux); fn.commit()
  fb = device.new_filter("RT")
  for k, v in t.items():
    fb.set_image(k, v)
  fb.set_image("output", out); fb.set("hdr", True); fb.set("cleanAux", True); fb.set_data("weights", tza_beauty)
  fb.commit()
  assert fb.info()["largeModel"] == 1
  fa.execute_async(); fn.execute_async(); fb.execute_async()
  device.sync()

  alb = np.zeros_like(imgs["albedo"]); nrm = np.zeros_like(imgs["normal"]); ref = np.zeros_like(imgs["color"])
  oracle.filter_execute(tza_aux, albedo=imgs["albedo"], output=alb)
  oracle.filter_execute(tza_aux, normal=imgs["normal"], output=nrm)
  e, p = metrics(t["albedo"].cpu().numpy(), alb)
  assert e <= MAX_ERR and p >= MIN_PSNR, ("albedo", e, p)
  e, p = metrics(t["normal"].cpu().numpy() * 0.5 + 0.5, nrm * 0.5 + 0.5)
  assert e <= MAX_ERR and p >= MIN_PSNR, ("normal", e, p)
  # the beauty pass against the oracle fed with the GPU's own prefiltered aux (isolates the pass) and
  # against the oracle's full chain
  ref2 = np.zeros_like(ref)
  oracle.filter_execute(tza_beauty, color=imgs["color"], albedo=t["albedo"].cpu().numpy(), normal=t["normal"].cpu().numpy(),
                        output=ref2, hdr=True)
  oracle.filter_execute(tza_beauty, color=imgs["color"], albedo=alb, normal=nrm, output=ref, hdr=True)
  got = out.cpu().numpy()
  for r in (ref2, ref):
    e, p = metrics(got, r)
    assert e <= MAX_ERR and p >= MIN_PSNR, ("beauty", e, p)
  fa.release(); fn.release(); fb.release()


@pytest.mark.parametrize("kind,ic,params", [("base", 9, dict(hdr=True)), ("small", 3, dict(quality=api.QUALITY_FAST)),
                                            ("large", 9, dict(hdr=True, cleanAux=True))], ids=["base", "small", "large"])
def test_fused_conv_pairs_equal_separate_launches(kind, ic, params):
  """Device parameter fusePairs (default on): enc_conv0 -> enc_conv1 and dec_conv1b -> dec_conv0 (and their
  counterparts in the small / large nets) as one launch each == two launches each, bit for bit, single- and
  multi-tile, with the output process as a separate pass; with it in the pair's epilogue, within one fp16 step."""
  W, H = 1500, 900
  tza = weights.model_tza(kind, ic, seed=0)
  imgs = synth.benchmark_images(W, H, hdr=bool(params.get("hdr")), albedo=(ic == 9), normal=(ic == 9), seed=8)
  dev = api.Device((0,)).commit()
  res = {}
  try:
    for pairs in (1, 0):
      dev.set("fusePairs", pairs)
      for fuse_out in (1, 0):
        dev.set("fuseOutput", fuse_out)
        for tiled in (False, True):
          extra = dict(maxMemoryMB=0) if tiled else {}
          res[pairs, fuse_out, tiled], info = run_filter(dev, tza, imgs["color"], imgs.get("albedo"), imgs.get("normal"), **params, **extra)
          assert (info["tileCountH"] * info["tileCountW"] > 1) == tiled
  finally:
    dev.release()
  base = res[0, 0, False]
  assert np.isfinite(base).all() and not np.any(base == -123.0)
  # pairs + fused output: the last conv is tap-packed (its horizontal taps are summed in the epilogue, so a few pixels
  # land one fp16 step of the network output away: tests/test_ops_gpu.py::test_conv_pair_tap_packed_last_conv) --
  # identical between tilings, and to everything else within that step through the inverse transfer function
  tap = res[1, 1, False]
  for k, v in res.items():
    if k[0] == 1 and k[1] == 1:
      np.testing.assert_array_equal(tap.view(np.uint32), v.view(np.uint32), err_msg=str(k))
    else:
      np.testing.assert_array_equal(base.view(np.uint32), v.view(np.uint32), err_msg=str(k))
  assert np.mean(tap != base) < 0.05
  np.testing.assert_allclose(tap, base, rtol=2e-2, atol=1e-4 * float(np.abs(base).max()))
