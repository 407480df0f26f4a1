"""Secondary parity witness: the reference's own CUDA device (baseline/_ref, built by
tools/build_reference_cuda.sh; fp16 CUTLASS path, not the fp32 oracle) and this backend on the same images and the
same TZA weights, both through their public C APIs, next to the CPU oracle.
usage (GPU box): python tools/reference_cuda_witness.py [W H]
Status: written at the end of round 1 when the GPU budget ran out (the one attempt failed on the library search path,
fixed since); to be run first thing in round 2 and promoted to a -m gpu test that skips without baseline/_ref."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def metrics(got, ref):
  got = np.asarray(got, np.float64); ref = np.asarray(ref, np.float64)
  peak = max(np.abs(ref).max(), 1e-30)
  mse = np.mean((got - ref) ** 2)
  return np.abs(got - ref).max() / peak, (200.0 if mse == 0 else 20 * np.log10(peak / np.sqrt(mse)))


def main():
  import torch
  from oidn_b200 import api, synth, weights
  import oracle as orc
  W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (640, 368)
  tza = weights.model_tza("base", 9, seed=0)
  imgs = synth.benchmark_images(W, H, hdr=True, seed=2)
  t = {k: torch.from_numpy(v).cuda() for k, v in imgs.items()}

  # this backend
  dev = api.Device((0,)).commit()
  f = dev.new_filter("RT")
  ours = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
  for k, v in t.items():
    f.set_image(k, v)
  f.set_image("output", ours); f.set("hdr", True); f.set_data("weights", tza); f.commit(); f.execute()
  ours_np = ours.cpu().numpy()
  f.release(); dev.release()

  # the reference's CUDA device through oidn.h (include/OpenImageDenoise/oidn.h)
  libdir = os.path.join(ROOT, "baseline", "_ref", "lib")
  C.CDLL(os.path.join(libdir, "libOpenImageDenoise_core.so.2.4.1"), mode=C.RTLD_GLOBAL)   # its rpath is the /tmp build dir
  R = C.CDLL(os.path.join(libdir, "libOpenImageDenoise.so"))
  R.oidnNewDevice.restype = C.c_void_p; R.oidnNewDevice.argtypes = [C.c_int]
  R.oidnCommitDevice.argtypes = [C.c_void_p]
  R.oidnNewFilter.restype = C.c_void_p; R.oidnNewFilter.argtypes = [C.c_void_p, C.c_char_p]
  R.oidnSetSharedFilterImage.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t]
  R.oidnSetSharedFilterData.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]
  R.oidnSetFilterBool.argtypes = [C.c_void_p, C.c_char_p, C.c_bool]
  R.oidnCommitFilter.argtypes = [C.c_void_p]; R.oidnExecuteFilter.argtypes = [C.c_void_p]
  R.oidnGetDeviceError.restype = C.c_int; R.oidnGetDeviceError.argtypes = [C.c_void_p, C.POINTER(C.c_char_p)]
  R.oidnReleaseFilter.argtypes = [C.c_void_p]; R.oidnReleaseDevice.argtypes = [C.c_void_p]

  def check(d):
    msg = C.c_char_p()
    code = R.oidnGetDeviceError(d, C.byref(msg))
    if code:
      raise RuntimeError("reference error %d: %s" % (code, (msg.value or b"").decode()))

  rd = R.oidnNewDevice(3)   # OIDN_DEVICE_TYPE_CUDA
  R.oidnCommitDevice(rd); check(rd)
  rf = R.oidnNewFilter(rd, b"RT"); check(rd)
  ref_out = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
  for k, v in list(t.items()) + [("output", ref_out)]:
    R.oidnSetSharedFilterImage(rf, k.encode(), v.data_ptr(), 3, W, H, 0, 12, 12 * W)   # OIDN_FORMAT_FLOAT3
  blob = C.create_string_buffer(tza, len(tza))
  R.oidnSetSharedFilterData(rf, b"weights", blob, len(tza))
  R.oidnSetFilterBool(rf, b"hdr", True)
  R.oidnCommitFilter(rf); check(rd)
  R.oidnExecuteFilter(rf); check(rd)
  ref_np = ref_out.cpu().numpy()
  R.oidnReleaseFilter(rf); R.oidnReleaseDevice(rd)

  oracle_np = np.zeros((H, W, 3), np.float32)
  orc.filter_execute(tza, color=imgs["color"], albedo=imgs["albedo"], normal=imgs["normal"], output=oracle_np, hdr=True)
  print("RT hdr+alb+nrm %dx%d, base UNet, same TZA weights (max|err|/peak, PSNR dB):" % (W, H))
  print("  oidn_b200            vs CPU oracle (fp32)     : %.3e, %.1f" % metrics(ours_np, oracle_np))
  print("  reference CUDA device vs CPU oracle (fp32)     : %.3e, %.1f" % metrics(ref_np, oracle_np))
  print("  oidn_b200            vs reference CUDA device : %.3e, %.1f" % metrics(ours_np, ref_np))


if __name__ == "__main__":
  main()
