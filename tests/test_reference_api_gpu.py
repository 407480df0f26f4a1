"""-m gpu: the drop-in boundary exercised through the REFERENCE's own library and applications.

baseline/_b200_ops (built in the build container by tools/build_integration_module.sh, binaries only, travels with
the snapshot) holds the UNMODIFIED reference core + C API (libOpenImageDenoise.so, libOpenImageDenoise_core.so,
oidnTest, oidnBenchmark) and this backend as the device module the core loads for OIDN_DEVICE_TYPE_CUDA
(integration/b200_module.cpp -DOIDN_B200_OP_LEVEL: the reference's filters, core/graph.cpp, arena planner and tile
loop drive this backend's Engine ops). The tests call oidnNewDevice -> oidnNewFilter("RT") -> oidnSetSharedFilterImage
-> oidnSetSharedFilterData("weights") -> oidnCommitFilter -> oidnExecuteFilter (include/OpenImageDenoise/oidn.h) and
check the result against the CPU oracle with the north-star tolerance, against this backend's own filter-level path,
and run the reference's test application on the module. Skipped when the binaries are absent (a checkout without
/root/reference cannot build them).
"""
import ctypes as C
import filecmp
import os
import shutil
import subprocess

import numpy as np
import pytest

from oidn_b200 import api, capi, synth, weights

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from gpu_util import metrics  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = os.path.join(ROOT, "baseline", "_b200_ops")
MAX_ERR, MIN_PSNR = 1e-2, 50.0

needs_module = pytest.mark.skipif(not os.path.exists(os.path.join(OPS, "lib", "libOpenImageDenoise_device_cuda.so.2.4.1")),
                                  reason="baseline/_b200_ops not built (tools/build_integration_module.sh)")


def _sync_backend_library():
  """The module resolves liboidn_b200.so next to itself: make that the library as built NOW."""
  src = os.path.join(ROOT, "oidn_b200", "liboidn_b200.so")
  dst = os.path.join(OPS, "lib", "liboidn_b200.so")
  if not os.path.exists(dst) or not filecmp.cmp(src, dst, shallow=False):
    shutil.copyfile(src, dst)


@pytest.fixture(scope="module")
def oidn():
  _sync_backend_library()
  lib = os.path.join(OPS, "lib")
  C.CDLL(os.path.join(lib, "libOpenImageDenoise_core.so.2.4.1"), mode=C.RTLD_GLOBAL)
  R = C.CDLL(os.path.join(lib, "libOpenImageDenoise.so.2.4.1"))
  R.oidnNewDevice.restype = C.c_void_p; R.oidnNewDevice.argtypes = [C.c_int]
  R.oidnCommitDevice.argtypes = [C.c_void_p]
  R.oidnSyncDevice.argtypes = [C.c_void_p]
  R.oidnGetDeviceInt.restype = C.c_int; R.oidnGetDeviceInt.argtypes = [C.c_void_p, C.c_char_p]
  R.oidnNewFilter.restype = C.c_void_p; R.oidnNewFilter.argtypes = [C.c_void_p, C.c_char_p]
  R.oidnSetSharedFilterImage.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int] + [C.c_size_t] * 5
  R.oidnSetSharedFilterData.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]
  R.oidnSetFilterBool.argtypes = [C.c_void_p, C.c_char_p, C.c_bool]
  R.oidnSetFilterInt.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
  R.oidnCommitFilter.argtypes = [C.c_void_p]
  R.oidnExecuteFilter.argtypes = [C.c_void_p]
  R.oidnExecuteFilterAsync.argtypes = [C.c_void_p]
  R.oidnGetDeviceError.restype = C.c_int; R.oidnGetDeviceError.argtypes = [C.c_void_p, C.POINTER(C.c_char_p)]
  R.oidnReleaseFilter.argtypes = [C.c_void_p]; R.oidnReleaseDevice.argtypes = [C.c_void_p]
  return R


def _check(R, d):
  msg = C.c_char_p()
  code = R.oidnGetDeviceError(d, C.byref(msg))
  assert code == 0, "reference API error %d: %s" % (code, (msg.value or b"").decode())


def _run_reference_api(R, tza, imgs, W, H, frames=1, **params):
  d = R.oidnNewDevice(3)                                   # OIDN_DEVICE_TYPE_CUDA -> the module of baseline/_b200_ops
  R.oidnCommitDevice(d); _check(R, d)
  assert R.oidnGetDeviceInt(d, b"type") == 3
  f = R.oidnNewFilter(d, b"RT"); _check(R, d)
  t = {k: torch.from_numpy(v).cuda() for k, v in imgs.items()}
  out = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
  for k, v in list(t.items()) + [("output", out)]:
    R.oidnSetSharedFilterImage(f, k.encode(), v.data_ptr(), 3, W, H, 0, 12, 12 * W)   # OIDN_FORMAT_FLOAT3
  blob = C.create_string_buffer(tza, len(tza))
  R.oidnSetSharedFilterData(f, b"weights", blob, len(tza))
  for k, v in params.items():
    if isinstance(v, bool):
      R.oidnSetFilterBool(f, k.encode(), v)
    else:
      R.oidnSetFilterInt(f, k.encode(), v)
  R.oidnCommitFilter(f); _check(R, d)
  for _ in range(frames):
    R.oidnExecuteFilterAsync(f)
  R.oidnSyncDevice(d); _check(R, d)
  got = out.cpu().numpy()
  R.oidnReleaseFilter(f); R.oidnReleaseDevice(d)
  return got


@needs_module
@pytest.mark.parametrize("case", [("base", 640, 368, {}), ("base", 2100, 1300, {"maxMemoryMB": 600}),
                                  ("large", 1300, 800, {"cleanAux": True, "quality": 6})],
                         ids=["base one tile", "base tiled by the reference's planner", "large cleanAux"])
def test_reference_api_runs_on_this_backend(case, oidn, oracle):
  kind, W, H, extra = case
  tza = weights.model_tza(kind, 9, seed=0)
  imgs = synth.benchmark_images(W, H, hdr=True, seed=2)
  got = _run_reference_api(oidn, tza, imgs, W, H, frames=2, hdr=True, **extra)
  ref = np.zeros((H, W, 3), np.float32)
  oracle.filter_execute(tza, output=ref, hdr=True, **imgs)
  e, p = metrics(got, ref)
  print("reference API on the op-level module, %s UNet %dx%d: max|err|/peak = %.3e, PSNR = %.1f dB" % (kind, W, H, e, p))
  assert e <= MAX_ERR and p >= MIN_PSNR, (e, p)
  if not extra.get("maxMemoryMB"):
    # same kernels under the reference's graph and under this backend's own graph: same bits (single tile)
    dev = api.Device((0,)).commit()
    t = {k: torch.from_numpy(v).cuda() for k, v in imgs.items()}
    out = torch.zeros((H, W, 3), device="cuda")
    f = dev.new_filter("RT")
    for k, v in t.items():
      f.set_image(k, v)
    f.set_image("output", out); f.set("hdr", True); f.set_data("weights", tza)
    for k, v in extra.items():
      f.set(k, v)
    f.commit(); f.execute()
    if f.info()["tileCountH"] * f.info()["tileCountW"] == 1:
      np.testing.assert_array_equal(out.cpu().numpy().view(np.uint32), got.view(np.uint32))
    f.release(); dev.release()


@needs_module
def test_reference_test_application_passes_on_the_module():
  """apps/oidnTest.cpp (the reference's Catch2 suite: API errors, buffers, single / multiple filters, in-place,
  filter updates, async, progress monitor + cancellation, sanitisation, weights) with --device cuda = this backend.
  The built-in weights of this reference build are the synthetic passthrough set (oidn_b200.weights.make_weights)."""
  _sync_backend_library()
  env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(OPS, "lib"))
  r = subprocess.run([os.path.join(OPS, "bin", "oidnTest"), "--device", "cuda"], capture_output=True, text=True, timeout=600, env=env)
  tail = r.stdout[-1500:]
  print(tail)
  assert r.returncode == 0, tail
  assert "All tests passed" in r.stdout
