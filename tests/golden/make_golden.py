"""Generates tests/golden/golden_v1.npz by running the REFERENCE's own Python implementation
(/root/reference/training: model.py, color.py, tza.py) in the build container.

Run here only (the reference tree does not exist on the GPU box):
    python tests/golden/make_golden.py
The committed .npz is what tests/test_oracle_golden.py pins oracle/oidn_oracle.c against; inputs and
weights are regenerated in the test from seeds (oidn_b200.synth / oidn_b200.weights) and their
sha256 is stored here so a drift in either generator is caught.

The per-image pipeline restates training/infer.py:67-116 (Infer.__call__) with the reference's
functions: color*exposure -> transfer.forward -> zero-pad to multiples of model.alignment ->
model -> crop -> clamp(min=0) -> transfer.inverse -> /exposure (hdr) or clamp(max=1); auxiliary
features enter as albedo and normal*0.5+0.5 (training/dataset.py preprocessing).
"""
import hashlib
import os
import sys
import tempfile
import types

import numpy as np
import torch
import torch.nn.functional as F

REF = "/root/reference/training"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
sys.modules.setdefault("OpenImageIO", types.ModuleType("OpenImageIO"))  # only needed by image I/O

import color as refcolor  # noqa: E402
import model as refmodel  # noqa: E402
import tza as reftza      # noqa: E402

from oidn_b200 import synth, weights  # noqa: E402

# (name, kind, ic, filter, mode, W, H)
CASES = [
  ("rt_hdr_alb_nrm_base", "base", 9, "RT", "hdr", 72, 40),
  ("rt_ldr_small", "small", 3, "RT", "ldr", 50, 34),
  ("rt_hdr_calb_cnrm_large", "large", 9, "RT", "hdr", 48, 32),
  ("rtlightmap_hdr_base", "base", 3, "RTLightmap", "hdr", 40, 40),
  ("rtlightmap_dir_base", "base", 3, "RTLightmap", "dir", 33, 17),
  ("rt_srgb_base", "base", 3, "RT", "srgb", 32, 32),
]


def sha(a):
  return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ref_model(kind, ic, tza_bytes):
  """Loads our TZA bytes through the reference's Reader into the reference's model class."""
  with tempfile.NamedTemporaryFile(suffix=".tza", delete=False) as f:
    f.write(tza_bytes)
    path = f.name
  reader = reftza.Reader(path)
  m = refmodel.UNetLarge(ic, 3) if kind == "large" else refmodel.UNet(ic, 3, small=(kind == "small"))
  sd = {}
  for name in m.state_dict().keys():
    t, layout = reader[name]
    sd[name] = torch.from_numpy(np.array(t).astype(np.float32))
  m.load_state_dict(sd)
  m.eval()
  os.unlink(path)
  return m


def ref_writer_bytes(tensors):
  """The same tensors through the reference's Writer (byte-compare with ours)."""
  with tempfile.NamedTemporaryFile(suffix=".tza", delete=False) as f:
    path = f.name
  with reftza.Writer(path) as w:
    for name, t in tensors.items():
      w.write(name, t, "oihw" if t.ndim == 4 else "x")
  data = open(path, "rb").read()
  os.unlink(path)
  return data


def run_case(name, kind, ic, filt, mode, W, H):
  torch.set_num_threads(4)
  wts = weights.make_weights(kind, ic, seed=0)
  blob = weights.write_tza(wts)
  assert blob == ref_writer_bytes(wts), "TZA writer differs from training/tza.py"
  m = ref_model(kind, ic, blob)

  hdr = mode == "hdr"
  imgs = synth.benchmark_images(W, H, hdr=hdr, albedo=(ic == 9), normal=(ic == 9), seed=1,
                                color_range=None)
  color = imgs["color"]
  if mode == "dir":
    color = (color * 2.0 - 1.0).astype(np.float32)  # directional lightmaps are signed [-1,1]
  if filt == "RTLightmap":
    tf = refcolor.LogTransferFunction() if hdr else refcolor.LinearTransferFunction()
  else:
    tf = {"hdr": refcolor.PUTransferFunction(), "ldr": refcolor.SRGBTransferFunction(),
          "srgb": refcolor.LinearTransferFunction()}[mode]
  exposure = refcolor.autoexposure(color) if hdr else 1.0
  c = torch.from_numpy(color).permute(2, 0, 1)[None].clone()
  if mode == "dir":
    c = c * 0.5 + 0.5                     # snorm -> [0,1] (cpu_input_process.isph:46-50)
  if hdr:
    c = c * exposure
  c = tf.forward(c)
  planes = [c]
  if ic == 9:
    planes.append(torch.from_numpy(imgs["albedo"]).permute(2, 0, 1)[None])
    planes.append(torch.from_numpy(imgs["normal"]).permute(2, 0, 1)[None] * 0.5 + 0.5)
  x = torch.cat(planes, 1)
  shape = x.shape
  al = m.alignment
  x = F.pad(x, (0, (shape[3] + al - 1) // al * al - shape[3], 0, (shape[2] + al - 1) // al * al - shape[2]))
  with torch.no_grad():
    y = m(x).float()
  y = y[:, :, :shape[2], :shape[3]]
  net = y.clone()
  y = torch.clamp(y, min=0.)
  y = tf.inverse(y)
  if mode == "dir":
    y = torch.clamp(y * 2.0 - 1.0, min=-1.0)   # cpu_output_process.isph:56-61
  if hdr:
    y = y / exposure
  else:
    y = torch.clamp(y, max=1.)
  out = y[0].permute(1, 2, 0).contiguous().numpy()
  netout = net[0].permute(1, 2, 0).contiguous().numpy()
  print("%-26s net out range [%.3f, %.3f] mean %.3f | image out range [%.4g, %.4g] exposure %.6g"
        % (name, netout.min(), netout.max(), netout.mean(), out.min(), out.max(), exposure))
  return {
    name + "/output": out.astype(np.float32),
    name + "/net": netout.astype(np.float32),
    name + "/exposure": np.float32(exposure),
    name + "/sha_color": np.bytes_(sha(color)),
    name + "/sha_tza": np.bytes_(hashlib.sha256(blob).hexdigest()),
  }


def transfer_tables():
  ys = np.concatenate([np.float32([0, 1e-7, 1.57945760e-06, 2e-6, 1e-4, 0.0031308, 0.01, 3.22087631e-02, 0.05]),
                       np.logspace(-5, np.log10(65504.0), 200, dtype=np.float32)]).astype(np.float32)
  xs = np.linspace(0, 1, 257, dtype=np.float32)
  out = {"tf/ys": ys, "tf/xs": xs}
  for nm, tf in (("pu", refcolor.PUTransferFunction()), ("log", refcolor.LogTransferFunction()),
                 ("srgb", refcolor.SRGBTransferFunction())):
    yy = ys if nm != "srgb" else np.clip(ys, 0, 1)
    out["tf/%s_fwd" % nm] = tf.forward(torch.from_numpy(yy)).numpy().astype(np.float32)
    out["tf/%s_inv" % nm] = tf.inverse(torch.from_numpy(xs)).numpy().astype(np.float32)
  out["tf/pu_norm"] = np.float32(refcolor.PU_NORM_SCALE)
  out["tf/log_norm"] = np.float32(refcolor.LOG_NORM_SCALE)
  return out


def autoexposure_cases():
  out = {}
  for i, (W, H) in enumerate([(37, 21), (64, 48), (130, 70)]):
    img = synth.benchmark_images(W, H, hdr=True, albedo=False, normal=False, seed=3 + i)["color"]
    if i == 2:
      img = (img * np.float32(1e-3)).astype(np.float32)
    out["ae/%d/value" % i] = np.float32(refcolor.autoexposure(img))
    out["ae/%d/dims" % i] = np.int32([W, H, 3 + i])
  return out


def main():
  data = {}
  for case in CASES:
    data.update(run_case(*case))
  data.update(transfer_tables())
  data.update(autoexposure_cases())
  path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
  np.savez_compressed(path, **data)
  print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
  main()
