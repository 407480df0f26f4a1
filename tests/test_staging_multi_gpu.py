"""-m gpu: the paths round 1 left to builder-run scripts, now under the driver's pytest run.

* tile staging (device parameter "staging"): frames in pinned host memory / on another GPU
  go through the copy-in / compute / copy-out pipeline and must equal the in-place run bit for bit;
* the reference-shaped multi-GPU entry point oidnb200NewCUDADevice(ids, streams, n > 1)
  (include/OpenImageDenoise/oidn.h:150-153, tiles dealt round-robin: core/unet_filter.cpp:219) -- with
  two engines on ONE GPU everywhere, with two GPUs where the box has them;
* tile sharding across processes (tools/sharded_check.py under torchrun) where the box has >= 2 GPUs;
* BASELINE.json's configs at full size against the CPU oracle (north-star tolerance);
* external memory by file descriptor (oidnNewSharedBufferFromFD, oidn.h:326-329).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from oidn_b200 import api, capi, synth, weights

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from gpu_util import metrics  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MAX_ERR, MIN_PSNR = 1e-2, 50.0


def ngpus():
  return torch.cuda.device_count() if torch.cuda.is_available() else 0


def denoise(dev, tza, imgs, out, **params):
  """imgs/out: dict name -> HxWx3 torch tensor (cuda or pinned host). Returns (filter info)."""
  f = dev.new_filter("RT")
  for k, v in imgs.items():
    f.set_image(k, v)
  f.set_image("output", out)
  for k, v in params.items():
    f.set(k, v)
  f.set_data("weights", tza)
  f.commit(); f.execute()
  info = f.info()
  f.release()
  return info


def reference_run(tza, frame, max_tile_pixels=None, **params):
  """Single engine, frame resident on GPU 0, images dereferenced in place."""
  dev = api.Device((0,)).commit()
  if max_tile_pixels:
    dev.set("maxTilePixels", max_tile_pixels)
  with torch.cuda.device(0):
    t = {k: torch.from_numpy(v).cuda() for k, v in frame.items()}
    out = torch.zeros_like(t["color"])
  info = denoise(dev, tza, t, out, **params)
  assert info["staged"] == 0
  res = out.cpu().numpy()
  dev.release()
  return res, info


@pytest.mark.parametrize("tiles", ["one tile", "four tiles"])
def test_staged_host_frame_equals_in_place(tiles):
  """Pinned host images: auto staging (copy engines, local tile images, per-tile autoexposure bins written into
  engine 0's bin array) == the in-place run on a device-resident frame, bit for bit; also in place in host memory."""
  W, H = 1500, 900
  tza = weights.model_tza("base", 9, seed=0)
  frame = synth.benchmark_images(W, H, hdr=True, seed=41)
  mtp = 1000 * 600 if tiles == "four tiles" else None
  ref, rinfo = reference_run(tza, frame, mtp, hdr=True)
  dev = api.Device((0,)).commit()
  if mtp:
    dev.set("maxTilePixels", mtp)
  host = {k: torch.from_numpy(v).pin_memory() for k, v in frame.items()}
  hout = torch.zeros((H, W, 3)).pin_memory()
  info = denoise(dev, tza, host, hout, hdr=True)
  assert info["staged"] == 1 and (info["tileCountH"], info["tileCountW"]) == (rinfo["tileCountH"], rinfo["tileCountW"])
  assert (info["tileCountH"] * info["tileCountW"] > 1) == (tiles == "four tiles")
  np.testing.assert_array_equal(hout.numpy().view(np.uint32), ref.view(np.uint32))
  # in place: the output aliases the colour image; no rectangle may land before every tile has been read
  hc = host["color"].clone().pin_memory()
  info = denoise(dev, tza, dict(host, color=hc), hc, hdr=True)
  assert info["staged"] == 1
  np.testing.assert_array_equal(hc.numpy().view(np.uint32), ref.view(np.uint32))
  # staging switched off: kernels dereference the pinned memory (zero copy), same result
  dev.set("staging", 0)
  hout.zero_()
  info = denoise(dev, tza, host, hout, hdr=True)
  assert info["staged"] == 0
  np.testing.assert_array_equal(hout.numpy().view(np.uint32), ref.view(np.uint32))
  dev.release()


def test_staged_frames_pipeline_async():
  """Back-to-back oidnb200ExecuteFilterAsync on a host frame: frames overlap inside the library (two slot sets),
  every frame's result is complete after oidnb200SyncDevice, input changes between frames are picked up, and a
  buffer read enqueued behind staged frames sees them (lazy join)."""
  W, H = 1280, 720
  tza = weights.model_tza("small", 3, seed=0)
  frames = [synth.benchmark_images(W, H, hdr=False, albedo=False, normal=False, seed=50 + i)["color"] for i in range(3)]
  refs = [reference_run(tza, {"color": fr}, quality=api.QUALITY_FAST)[0] for fr in frames]
  dev = api.Device((0,)).commit()
  hin = [torch.from_numpy(fr).pin_memory() for fr in frames]
  houts = [torch.zeros((H, W, 3)).pin_memory() for _ in frames]
  f = dev.new_filter("RT")
  f.set("quality", api.QUALITY_FAST); f.set_data("weights", tza)
  for it in range(6):
    i = it % 3
    f.set_image("color", hin[i]); f.set_image("output", houts[i])
    f.commit()                      # pointer-only change: no rebuild (core/filter.cpp:52-56)
    f.execute_async()
  dev.sync()
  assert f.info()["staged"] == 1
  for i in range(3):
    np.testing.assert_array_equal(houts[i].numpy().view(np.uint32), refs[i].view(np.uint32))
  # device buffer as output of a staged frame (input in host memory), read back asynchronously right behind it
  bo = dev.new_buffer(W * H * 12)
  f.set_image("color", hin[1]); f.set_image("output", bo, capi.FORMAT_FLOAT3, W, H); f.commit()
  dev.set("staging", 1)
  f.execute_async()
  got = np.zeros((H, W, 3), np.float32)
  bo.read(got, sync=False)
  dev.sync()
  np.testing.assert_array_equal(got.view(np.uint32), refs[1].view(np.uint32))
  f.release(); bo.release(); dev.release()


@pytest.mark.parametrize("staging", [-1, 0])
def test_two_engines_on_one_gpu(staging):
  """oidnb200NewCUDADevice((0, 0)): tiles dealt round-robin to two (GPU, stream) pairs, event barrier, autoexposure
  on engine 0 (staging=0) or exchanged through engine 0's bin array (staged) -- bit-identical to one engine."""
  W, H = 2000, 1100
  tza = weights.model_tza("base", 9, seed=0)
  frame = synth.benchmark_images(W, H, hdr=True, seed=43)
  dev = api.Device((0, 0)).commit()
  dev.set("staging", staging)
  assert dev.get("numSubdevices") == 2
  t = {k: torch.from_numpy(v).cuda() for k, v in frame.items()}
  out = torch.zeros((H, W, 3), device="cuda")
  info = denoise(dev, tza, t, out, hdr=True)
  assert (info["tileCountH"] * info["tileCountW"]) % 2 == 0 and info["staged"] == (1 if staging else 0)
  # one engine with the same plan: shrink the tile budget until the grids agree
  ref, rinfo = reference_run(tza, frame, info["tileH"] * info["tileW"], hdr=True)
  assert (rinfo["tileCountH"], rinfo["tileCountW"], rinfo["tileH"], rinfo["tileW"]) == \
         (info["tileCountH"], info["tileCountW"], info["tileH"], info["tileW"])
  np.testing.assert_array_equal(out.cpu().numpy().view(np.uint32), ref.view(np.uint32))
  # progress: one unit per op of every tile, monotonic, from both engines' streams
  f = dev.new_filter("RT")
  for k, v in t.items():
    f.set_image(k, v)
  f.set_image("output", out); f.set("hdr", True); f.set_data("weights", tza)
  seen = []
  f.set_progress_monitor(lambda n: (seen.append(n), True)[1])
  f.commit(); f.execute()
  ntiles = info["tileCountH"] * info["tileCountW"]
  assert seen[0] == 0.0 and seen[-1] == 1.0 and all(b >= a for a, b in zip(seen, seen[1:]))
  assert len(seen) == 1 + ntiles * info["numOps"] + 1, (len(seen), ntiles, info["numOps"])
  f.release(); dev.release()


@pytest.mark.skipif(ngpus() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("where", ["gpu0", "host"])
def test_two_gpus_one_device(where):
  """The reference's multi-pair signature on two B200s: the frame lives in GPU 0's HBM (single-pointer contract of
  oidnSetSharedFilterImage) or in pinned host memory; each GPU stages its own tiles over NVLink / its own PCIe link."""
  n = min(ngpus(), 4)
  W, H = 2600, 1500
  tza = weights.model_tza("base", 9, seed=0)
  frame = synth.benchmark_images(W, H, hdr=True, seed=44)
  dev = api.Device(tuple(range(n))).commit()
  if where == "gpu0":
    with torch.cuda.device(0):
      t = {k: torch.from_numpy(v).cuda() for k, v in frame.items()}
      out = torch.zeros((H, W, 3), device="cuda")
  else:
    hb = {k: dev.new_buffer(v.nbytes, api.STORAGE_HOST) for k, v in frame.items()}   # NUMA-interleaved pinned memory
    ho = dev.new_buffer(W * H * 12, api.STORAGE_HOST)
    for k, v in frame.items():
      hb[k].write(v)
  f = dev.new_filter("RT")
  if where == "gpu0":
    for k, v in t.items():
      f.set_image(k, v)
    f.set_image("output", out)
  else:
    for k, b in hb.items():
      f.set_image(k, b, capi.FORMAT_FLOAT3, W, H)
    f.set_image("output", ho, capi.FORMAT_FLOAT3, W, H)
  f.set("hdr", True); f.set_data("weights", tza); f.commit()
  for _ in range(3):
    f.execute_async()
  dev.sync()
  info = f.info()
  assert info["staged"] == 1 and (info["tileCountH"] * info["tileCountW"]) % n == 0
  if where == "gpu0":
    got = out.cpu().numpy()
  else:
    got = np.zeros((H, W, 3), np.float32); ho.read(got)
  ref, rinfo = reference_run(tza, frame, info["tileH"] * info["tileW"], hdr=True)
  assert (rinfo["tileCountH"], rinfo["tileCountW"]) == (info["tileCountH"], info["tileCountW"])
  np.testing.assert_array_equal(got.view(np.uint32), ref.view(np.uint32))
  f.release()
  if where == "host":
    for b in list(hb.values()) + [ho]:
      b.release()
  dev.release()


@pytest.mark.skipif(ngpus() < 2, reason="needs 2 GPUs")
def test_sharded_processes_bit_identical():
  """One process per GPU (torchrun, NCCL): tools/sharded_check.py -- frame on rank 0 (staged and direct P2P) and
  distributed frame == one GPU with the same tile plan, bit for bit, and within tolerance of the oracle."""
  n = 2 if ngpus() < 4 else 4
  cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
         "--master-port", "29611", os.path.join(ROOT, "tools", "sharded_check.py")]
  r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
  assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
  assert "bit-identical to single GPU: True" in r.stdout


def test_sharded_in_place_copies_back_own_tiles_only():
  """numShards > 1 with in-place multi-tile filtering: a shard writes its own rectangles and leaves the rest of the
  image alone (the temporary it renders into is uninitialised outside its tiles)."""
  W, H = 1500, 900
  tza = weights.model_tza("base", 3, seed=0)
  color = synth.benchmark_images(W, H, hdr=False, albedo=False, normal=False, seed=45)["color"]
  dev = api.Device((0,)).commit()
  dev.set("maxTilePixels", 1000 * 600)
  _, tiles = api.plan_tiles(H, W, False, 1, 2, 1000 * 600, 1)
  _, rinfo = reference_run(tza, {"color": color}, 1000 * 600, numShards=2, shardIndex=0)   # the plan for 2 units
  whole = np.zeros_like(color)
  for shard in (0, 1):
    c = torch.from_numpy(color).cuda()
    f = dev.new_filter("RT")
    f.set_image("color", c); f.set_image("output", c)
    f.set("numShards", 2); f.set("shardIndex", shard); f.set_data("weights", tza)
    f.commit(); f.execute()
    info = f.info()
    assert info["tileCountH"] * info["tileCountW"] == len(tiles) > 1
    got = c.cpu().numpy()
    mine = np.zeros((H, W), bool)
    for i, t in enumerate(tiles):
      if i % 2 == shard:
        mine[t["hDst"]:t["hDst"] + t["H2"], t["wDst"]:t["wDst"] + t["W2"]] = True
    np.testing.assert_array_equal(got[~mine].view(np.uint32), color[~mine].view(np.uint32))   # other shards' pixels untouched
    assert not np.array_equal(got[mine], color[mine])
    whole[mine] = got[mine]
    f.release()
  # the two shards together == the unsharded filter with the same plan
  dev1 = api.Device((0,)).commit(); dev1.set("maxTilePixels", rinfo["tileH"] * rinfo["tileW"])
  c = torch.from_numpy(color).cuda(); o = torch.zeros_like(c)
  i1 = denoise(dev1, tza, {"color": c}, o)
  if (i1["tileCountH"], i1["tileCountW"]) == (rinfo["tileCountH"], rinfo["tileCountW"]):
    np.testing.assert_array_equal(whole.view(np.uint32), o.cpu().numpy().view(np.uint32))
  dev1.release(); dev.release()


def test_released_buffer_stays_alive_while_set():
  """oidnSetFilterImage keeps the buffer (core/image.h: Image holds a Ref<Buffer>): releasing it right after is legal."""
  W, H = 256, 144
  tza = weights.model_tza("small", 3, seed=0)
  color = synth.benchmark_images(W, H, hdr=False, albedo=False, normal=False, seed=46)["color"]
  ref, _ = reference_run(tza, {"color": color}, quality=api.QUALITY_FAST)
  dev = api.Device((0,)).commit()
  bc = dev.new_buffer(color.nbytes); bo = dev.new_buffer(color.nbytes)
  bc.write(color)
  f = dev.new_filter("RT")
  f.set_image("color", bc, capi.FORMAT_FLOAT3, W, H); f.set_image("output", bo, capi.FORMAT_FLOAT3, W, H)
  bc.release()                                       # the filter still holds it
  junk = [torch.full((color.nbytes // 4,), 7.0, device="cuda") for _ in range(4)]   # would reuse the freed block
  f.set("quality", api.QUALITY_FAST); f.set_data("weights", tza); f.commit(); f.execute()
  got = np.zeros_like(color); bo.read(got)
  np.testing.assert_array_equal(got.view(np.uint32), ref.view(np.uint32))
  del junk
  f.release(); bo.release(); dev.release()


def test_shared_buffer_from_fd_round_trip():
  """oidnNewSharedBufferFromFD (oidn.h:326-329, devices/cuda/cuda_external_buffer.cpp): device memory exported as an
  opaque fd, imported as a second buffer aliasing the same memory, used as a filter image."""
  W, H = 320, 200
  tza = weights.model_tza("small", 3, seed=0)
  color = synth.benchmark_images(W, H, hdr=False, albedo=False, normal=False, seed=47)["color"]
  ref, _ = reference_run(tza, {"color": color}, quality=api.QUALITY_FAST)
  dev = api.Device((0,)).commit()
  with pytest.raises(api.Error) as ei:
    dev.import_fd(0, 4096, capi.EXTERNAL_MEMORY_DMA_BUF)         # the CUDA device takes opaque fds only
  assert ei.value.code == capi.ERROR_INVALID_ARGUMENT
  r, w = os.pipe()
  with pytest.raises(api.Error) as ei:
    dev.import_fd(r, 4096)                                       # not a memory object
  assert ei.value.code == capi.ERROR_INVALID_ARGUMENT
  for fd in (r, w):
    try:
      os.close(fd)      # the driver may already have closed the read end while probing it
    except OSError:
      pass
  owner = dev.new_exportable_buffer(color.nbytes)                # the "renderer" side
  with pytest.raises(api.Error):
    dev.new_buffer(64).fd()                                      # plain buffers are not exportable
  fd = owner.fd()
  assert fd >= 0
  shared = dev.import_fd(fd, color.nbytes)                       # owns fd now
  assert shared.size == color.nbytes and shared.data != owner.data
  owner.write(color)                                             # written through one mapping ...
  back = np.zeros_like(color); shared.read(back)                 # ... visible through the other
  np.testing.assert_array_equal(back, color)
  bo = dev.new_buffer(color.nbytes)
  f = dev.new_filter("RT")
  f.set_image("color", shared, capi.FORMAT_FLOAT3, W, H); f.set_image("output", bo, capi.FORMAT_FLOAT3, W, H)
  f.set("quality", api.QUALITY_FAST); f.set_data("weights", tza); f.commit(); f.execute()
  got = np.zeros_like(color); bo.read(got)
  np.testing.assert_array_equal(got.view(np.uint32), ref.view(np.uint32))
  f.release(); shared.release(); owner.release(); bo.release(); dev.release()


# ---- BASELINE.json configs at full size against the oracle ---------------------------------------------------------
FULL = [
  ("config1 RT hdr+alb+nrm 1920x1080", "RT", "base", 9, 1920, 1080, dict(hdr=True)),
  ("config2 RT hdr+alb+nrm 3840x2160 quality=high", "RT", "base", 9, 3840, 2160, dict(hdr=True, quality=api.QUALITY_HIGH)),
  ("config4 RTLightmap hdr 4096x4096", "RTLightmap", "base", 3, 4096, 4096, dict()),
  ("config5 RT ldr 1280x720 quality=fast", "RT", "small", 3, 1280, 720, dict(quality=api.QUALITY_FAST)),
]


@pytest.mark.parametrize("case", FULL, ids=[c[0].split()[0] for c in FULL])
def test_baseline_config_full_size_vs_oracle(case, oracle):
  name, filt, kind, ic, W, H, params = case
  tza = weights.model_tza(kind, ic, seed=0)
  hdr = bool(params.get("hdr")) or filt == "RTLightmap"
  imgs = synth.benchmark_images(W, H, hdr=hdr, albedo=(ic == 9), normal=(ic == 9), seed=1)
  dev = api.Device((0,)).commit()
  t = {k: torch.from_numpy(v).cuda() for k, v in imgs.items()}
  out = torch.zeros((H, W, 3), device="cuda")
  f = dev.new_filter(filt)
  for k, v in t.items():
    f.set_image(k, v)
  f.set_image("output", out)
  for k, v in params.items():
    f.set(k, v)
  f.set_data("weights", tza); f.commit(); f.execute()
  got = out.cpu().numpy()
  f.release(); dev.release()
  ref = np.zeros((H, W, 3), np.float32)
  oracle.filter_execute(tza, output=ref, filter=filt, hdr=hdr, **imgs)
  e, p = metrics(got, ref)
  print("%s: max|err|/peak = %.3e, PSNR = %.1f dB" % (name, e, p))
  assert e <= MAX_ERR and p >= MIN_PSNR, (name, e, p)


@pytest.mark.parametrize("W,H,engines", [(2400, 1700, 1), (2400, 1700, 2)] +
                         ([(7680, 4320, 1)] if os.environ.get("OIDN_B200_FULL_8K") == "1" else []))
def test_large_unet_clean_aux_multi_tile_vs_oracle(W, H, engines, oracle):
  """BASELINE config 3's model (large UNet: cleanAux=true, quality=high; core/unet_filter.cpp:417-436,449-452) on a
  forced multi-tile plan, one engine and two engines, against the ORACLE (not against itself). The full 7680x4320
  frame runs with OIDN_B200_FULL_8K=1 (the oracle needs ~2 minutes of host time; log under profiles/)."""
  tza = weights.model_tza("large", 9, seed=0)
  imgs = synth.benchmark_images(W, H, hdr=True, seed=5)
  dev = api.Device((0,) * engines).commit()
  if W < 7680:
    dev.set("maxTilePixels", 1300 * 1000)
  t = {k: torch.from_numpy(v).cuda() for k, v in imgs.items()}
  out = torch.zeros((H, W, 3), device="cuda")
  info = denoise(dev, tza, t, out, hdr=True, cleanAux=True, quality=api.QUALITY_HIGH)
  assert info["largeModel"] == 1 and info["tileOverlap"] == 112
  assert W >= 7680 or info["tileCountH"] * info["tileCountW"] >= 4
  got = out.cpu().numpy()
  dev.release()
  ref = np.zeros((H, W, 3), np.float32)
  oracle.filter_execute(tza, output=ref, hdr=True, **imgs)
  e, p = metrics(got, ref)
  print("large UNet cleanAux %dx%d, %dx%d tiles, %d engine(s): max|err|/peak = %.3e, PSNR = %.1f dB"
        % (W, H, info["tileCountW"], info["tileCountH"], engines, e, p))
  assert e <= MAX_ERR and p >= MIN_PSNR, (e, p)


def test_staged_half_and_row_strided_host_images():
  """Staging moves row segments with copy engines: half-precision images (6-byte pixels), a row stride wider than the
  image, and a 1-channel image go through it unchanged; strided PIXELS are not staged (copy engines cannot skip
  bytes) and fall back to zero copy. Every variant == the same images resident on the device."""
  W, H = 1700, 420
  tza = weights.model_tza("base", 3, seed=0)
  color = synth.benchmark_images(W, H, hdr=False, albedo=False, normal=False, seed=48)["color"]
  dev = api.Device((0,)).commit()
  dev.set("maxTilePixels", 1000 * 432)     # two tiles

  def run(cimg, oimg):
    f = dev.new_filter("RT")
    f.set_image("color", cimg); f.set_image("output", oimg); f.set_data("weights", tza)
    f.commit(); f.execute()
    info = f.info()
    f.release()
    return info

  for dtype in (torch.float16, torch.float32):
    c = torch.from_numpy(color).to(dtype)
    ref = torch.zeros((H, W, 3), dtype=dtype, device="cuda")
    assert run(c.cuda(), ref)["staged"] == 0
    # packed host images
    hc, ho = c.clone().pin_memory(), torch.zeros((H, W, 3), dtype=dtype).pin_memory()
    info = run(hc, ho)
    assert info["staged"] == 1 and info["tileCountH"] * info["tileCountW"] > 1
    assert torch.equal(ho, ref.cpu())
    # rows padded to W + 9 pixels
    big_c = torch.zeros((H, W + 9, 3), dtype=dtype).pin_memory(); big_c[:, 4:4 + W] = c
    big_o = torch.full((H, W + 9, 3), -5.0, dtype=dtype).pin_memory()
    assert run(big_c[:, 4:4 + W], big_o[:, 2:2 + W])["staged"] == 1
    assert torch.equal(big_o[:, 2:2 + W], ref.cpu())
    assert bool((big_o[:, :2] == -5).all()) and bool((big_o[:, 2 + W:] == -5).all()), "wrote outside the image"
    # strided pixels (4 values per pixel, 3 used): not staged, dereferenced in place
    wide_c = torch.zeros((H, W, 4), dtype=dtype).pin_memory(); wide_c[:, :, :3] = c
    wide_o = torch.zeros((H, W, 4), dtype=dtype).pin_memory()
    assert run(wide_c[:, :, :3], wide_o[:, :, :3])["staged"] == 0
    assert not bool(wide_o[:, :, 3].any())
    if dtype == torch.float16:
      assert torch.equal(wide_o[:, :, :3], ref.cpu())
    else:
      # a 16-byte fp32 pixel is not written by the fused output process, so the last conv runs un-packed here and
      # tap-packed in `ref` (tests/test_ops_gpu.py::test_conv_pair_tap_packed_last_conv): one fp16 step of the network
      # output in a few pixels, through the sRGB inverse (x = y^2.4)
      a, b = wide_o[:, :, :3].numpy(), ref.cpu().numpy()
      assert np.mean(a != b) < 0.05
      np.testing.assert_allclose(a, b, rtol=5e-3, atol=1e-5)
  # one channel
  c1 = torch.from_numpy(color[:, :, :1].copy())
  ref1 = torch.zeros((H, W, 1), device="cuda")
  run(c1.cuda(), ref1)
  h1, o1 = c1.clone().pin_memory(), torch.zeros((H, W, 1)).pin_memory()
  assert run(h1, o1)["staged"] == 1
  assert torch.equal(o1, ref1.cpu())
  dev.release()


def test_staged_progress_and_cancel():
  """Progress monitor on a staged frame: one unit per op of every tile (+1 for the autoexposure), reported from the
  engines' compute streams; returning false cancels with OIDN_ERROR_CANCELLED and leaves the device usable."""
  W, H = 900, 500
  tza = weights.model_tza("base", 9, seed=0)
  frame = synth.benchmark_images(W, H, hdr=True, seed=49)
  dev = api.Device((0, 0)).commit()
  host = {k: torch.from_numpy(v).pin_memory() for k, v in frame.items()}
  hout = torch.zeros((H, W, 3)).pin_memory()
  f = dev.new_filter("RT")
  for k, v in host.items():
    f.set_image(k, v)
  f.set_image("output", hout); f.set("hdr", True); f.set_data("weights", tza)
  seen = []
  f.set_progress_monitor(lambda n: (seen.append(n), True)[1])
  f.commit(); f.execute()
  info = f.info()
  ntiles = info["tileCountH"] * info["tileCountW"]
  assert info["staged"] == 1 and ntiles % 2 == 0
  assert seen[0] == 0.0 and seen[-1] == 1.0 and all(b >= a for a, b in zip(seen, seen[1:]))
  assert len(seen) == 1 + ntiles * info["numOps"] + 1
  good = hout.clone()
  f.set_progress_monitor(lambda n: n < 0.3)
  with pytest.raises(api.Error) as ei:
    f.execute()
  assert ei.value.code == capi.ERROR_CANCELLED
  f.set_progress_monitor(None)
  hout.zero_()
  f.execute()
  assert torch.equal(hout, good)
  f.release(); dev.release()
