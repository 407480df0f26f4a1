"""Pins oracle/oidn_oracle.c against vectors produced by the REFERENCE's own PyTorch implementation
(training/model.py, color.py, tza.py; generator: tests/golden/make_golden.py, run in the build
container where /root/reference exists). Inputs and weights are regenerated from seeds and their
sha256 must match what the generator saw."""
import hashlib

import numpy as np
import pytest

from oidn_b200 import synth, weights

# (name, kind, ic, filter, mode, W, H) -- same table as tests/golden/make_golden.py
CASES = [
  ("rt_hdr_alb_nrm_base", "base", 9, "RT", "hdr", 72, 40),
  ("rt_ldr_small", "small", 3, "RT", "ldr", 50, 34),
  ("rt_hdr_calb_cnrm_large", "large", 9, "RT", "hdr", 48, 32),
  ("rtlightmap_hdr_base", "base", 3, "RTLightmap", "hdr", 40, 40),
  ("rtlightmap_dir_base", "base", 3, "RTLightmap", "dir", 33, 17),
  ("rt_srgb_base", "base", 3, "RT", "srgb", 32, 32),
]


def case_inputs(kind, ic, mode, W, H):
  hdr = mode == "hdr"
  imgs = synth.benchmark_images(W, H, hdr=hdr, albedo=(ic == 9), normal=(ic == 9), seed=1)
  color = imgs["color"]
  if mode == "dir":
    color = (color * 2.0 - 1.0).astype(np.float32)
  return weights.model_tza(kind, ic, seed=0), color, imgs.get("albedo"), imgs.get("normal")


def sha(a):
  return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_filter_matches_reference_pytorch(case, golden, oracle):
  name, kind, ic, filt, mode, W, H = case
  tza, color, albedo, normal = case_inputs(kind, ic, mode, W, H)
  assert sha(color) == golden[name + "/sha_color"].item().decode(), "synthetic input generator drifted"
  assert hashlib.sha256(tza).hexdigest() == golden[name + "/sha_tza"].item().decode(), "weight generator drifted"
  out = np.zeros((H, W, 3), np.float32)
  st = oracle.filter_execute(tza, color=color, albedo=albedo, normal=normal, output=out, filter=filt,
                             hdr=(mode == "hdr"), srgb=(mode == "srgb"), directional=(mode == "dir"))
  ref = golden[name + "/output"]
  if mode == "hdr":
    assert abs(st["input_scale"] - float(golden[name + "/exposure"])) <= 2e-6 * float(golden[name + "/exposure"])
  # fp32 C vs fp32 PyTorch (different summation order + libm): tight relative agreement
  scale = np.abs(ref).max()
  assert np.abs(out - ref).max() <= 2e-4 * scale, (np.abs(out - ref).max(), scale)
  assert st["large"] == int(kind == "large")


def test_transfer_functions(golden, oracle):
  ys, xs = golden["tf/ys"], golden["tf/xs"]
  for nm, t in (("pu", oracle.TF_PU), ("log", oracle.TF_LOG), ("srgb", oracle.TF_SRGB)):
    yy = ys if nm != "srgb" else np.clip(ys, 0, 1)
    np.testing.assert_allclose(oracle.tf_forward(t, yy), golden["tf/%s_fwd" % nm], rtol=3e-6, atol=1e-7)
    np.testing.assert_allclose(oracle.tf_inverse(t, xs), golden["tf/%s_inv" % nm], rtol=2e-5, atol=1e-7)
  assert abs(oracle.lib().oro_tf_norm_scale(oracle.TF_PU) - float(golden["tf/pu_norm"])) < 1e-7
  assert abs(oracle.lib().oro_tf_norm_scale(oracle.TF_LOG) - float(golden["tf/log_norm"])) < 1e-7


def test_autoexposure(golden, oracle):
  for i in range(3):
    W, H, seed = [int(v) for v in golden["ae/%d/dims" % i]]
    img = synth.benchmark_images(W, H, hdr=True, albedo=False, normal=False, seed=seed)["color"]
    if i == 2:
      img = (img * np.float32(1e-3)).astype(np.float32)
    ref = float(golden["ae/%d/value" % i])
    assert abs(oracle.autoexposure(img) - ref) <= 3e-6 * ref


def test_half_conversion_matches_numpy(oracle):
  rng = np.random.default_rng(0)
  f = np.concatenate([rng.standard_normal(2000).astype(np.float32) * s for s in (1e-8, 1e-4, 1, 1e3, 1e5)])
  f = np.concatenate([f, np.float32([0, -0.0, np.inf, -np.inf, 65504, 65520, 6e-8, 5.96e-8, 3e-8])])
  L = oracle.lib()
  got = np.array([L.oro_float_to_half(float(v)) for v in f], np.uint16)
  with np.errstate(over="ignore"):
    np.testing.assert_array_equal(got, f.astype(np.float16).view(np.uint16))
  h = np.arange(0, 65536, 7, dtype=np.uint16)
  back = np.array([L.oro_half_to_float(int(v)) for v in h], np.float32)
  np.testing.assert_array_equal(back.view(np.uint32)[~np.isnan(back)], h.view(np.float16).astype(np.float32).view(np.uint32)[~np.isnan(back)])
