"""Synthetic noisy inputs, generated the way the reference's own benchmark does.

apps/utils/random.h:12-48 (LCG: state = 1664525*state + 1013904223, float = state * 2^-32) and
apps/oidnBenchmark.cpp:96-145 (one generator seeded with 1; images filled in the order albedo
U[0,1), normal U[-1,1), color U[0,100) for hdr / U[0,1) for ldr; 3 channels, row-major).
"""
import numpy as np

_MUL = np.uint64(1664525)
_INC = np.uint64(1013904223)
_MASK = np.uint64(0xFFFFFFFF)


class Random:
  """Vectorised restatement of the reference LCG (jump-ahead by closed form of the affine map)."""

  def __init__(self, seed=1):
    self.state = np.uint64(seed)

  def floats(self, n):
    # state_k = a^k * s + c*(a^k - 1)/(a - 1)  (mod 2^32); build a^k and the geometric sums by doubling
    n = int(n)
    mul = np.empty(n, dtype=np.uint64)
    add = np.empty(n, dtype=np.uint64)
    if n == 0:
      return np.empty(0, dtype=np.float32)
    mul[0] = _MUL
    add[0] = _INC
    filled = 1
    while filled < n:
      m = min(filled, n - filled)
      # compose: step (filled + i) = step(filled) after step(i)  for i in 1..m
      am, cm = mul[filled - 1], add[filled - 1]
      mul[filled:filled + m] = (mul[:m] * am) & _MASK
      add[filled:filled + m] = (add[:m] * am + cm) & _MASK
      filled += m
    states = (mul * self.state + add) & _MASK
    self.state = states[-1]
    return (states.astype(np.float32) * np.float32(2.3283064365386962890625e-10)).astype(np.float32)


def benchmark_images(width, height, hdr=True, albedo=True, normal=True, seed=1, color_range=None):
  """Returns dict(color, albedo, normal) of float32 HxWx3 arrays, oidnBenchmark order and ranges."""
  rng = Random(seed)
  n = width * height * 3
  out = {}
  if albedo:
    out["albedo"] = rng.floats(n).reshape(height, width, 3)
  if normal:
    out["normal"] = (np.float32(-1.0) + rng.floats(n) * np.float32(2.0)).reshape(height, width, 3)
  hi = np.float32(color_range if color_range is not None else (100.0 if hdr else 1.0))
  out["color"] = (rng.floats(n) * hi).reshape(height, width, 3)
  return out
