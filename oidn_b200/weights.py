"""Synthetic UNet weights in the reference's TZA format.

The reference's bundled weights/*.tza are Git-LFS pointers in this checkout, so every config runs
on random-init weights of the same architecture (BASELINE.json north_star). Layer names and shapes
follow training/model.py:59-103 (UNet / small) and :165-208 (UNetLarge); the container layout follows
training/tza.py:12-108 as parsed by core/tza.cpp:27-103. numpy's PCG64 stream is stable across
versions and machines, so a (kind, ic, seed) triple names the same bytes everywhere.
"""
import io
import struct

import numpy as np

TZA_MAGIC = 0x41D7
TZA_VERSION = (2, 0)


def unet_layers(kind, ic, oc=3):
  """[(name, cin, cout)] in state-dict order (training/model.py)."""
  if kind in ("base", "small"):
    if kind == "small":
      ec1, ec2, ec3, ec4, ec5, dc4, dc3, dc2a, dc2b, dc1a, dc1b = 32, 32, 32, 32, 32, 64, 64, 64, 32, 32, 32
    else:
      ec1, ec2, ec3, ec4, ec5, dc4, dc3, dc2a, dc2b, dc1a, dc1b = 32, 48, 64, 80, 96, 112, 96, 64, 64, 64, 32
    return [("enc_conv0", ic, ec1), ("enc_conv1", ec1, ec1), ("enc_conv2", ec1, ec2),
            ("enc_conv3", ec2, ec3), ("enc_conv4", ec3, ec4), ("enc_conv5a", ec4, ec5),
            ("enc_conv5b", ec5, ec5), ("dec_conv4a", ec5 + ec3, dc4), ("dec_conv4b", dc4, dc4),
            ("dec_conv3a", dc4 + ec2, dc3), ("dec_conv3b", dc3, dc3), ("dec_conv2a", dc3 + ec1, dc2a),
            ("dec_conv2b", dc2a, dc2b), ("dec_conv1a", dc2b + ic, dc1a), ("dec_conv1b", dc1a, dc1b),
            ("dec_conv0", dc1b, oc)]
  if kind == "large":
    ec1, ec2, ec3, ec4, ec5, dc4, dc3, dc2, dc1 = 64, 96, 128, 192, 256, 192, 128, 96, 64
    return [("enc_conv1a", ic, ec1), ("enc_conv1b", ec1, ec1), ("enc_conv2a", ec1, ec2),
            ("enc_conv2b", ec2, ec2), ("enc_conv3a", ec2, ec3), ("enc_conv3b", ec3, ec3),
            ("enc_conv4a", ec3, ec4), ("enc_conv4b", ec4, ec4), ("enc_conv5a", ec4, ec5),
            ("enc_conv5b", ec5, ec5), ("dec_conv4a", ec5 + ec3, dc4), ("dec_conv4b", dc4, dc4),
            ("dec_conv3a", dc4 + ec2, dc3), ("dec_conv3b", dc3, dc3), ("dec_conv2a", dc3 + ec1, dc2),
            ("dec_conv2b", dc2, dc2), ("dec_conv1a", dc2 + ic, dc1), ("dec_conv1b", dc1, dc1),
            ("dec_conv1c", dc1, oc)]
  raise ValueError("unknown UNet kind: %r" % (kind,))


def flops_per_pixel(kind, ic):
  """Algorithmic FLOP per output pixel (unpadded channels), SURVEY.md section 8(d)."""
  res = {"enc_conv0": 1, "enc_conv1": 1, "enc_conv1a": 1, "enc_conv1b": 1, "enc_conv2": 4, "enc_conv2a": 4,
         "enc_conv2b": 4, "enc_conv3": 16, "enc_conv3a": 16, "enc_conv3b": 16, "enc_conv4": 64,
         "enc_conv4a": 64, "enc_conv4b": 64, "enc_conv5a": 256, "enc_conv5b": 256, "dec_conv4a": 64,
         "dec_conv4b": 64, "dec_conv3a": 16, "dec_conv3b": 16, "dec_conv2a": 4, "dec_conv2b": 4,
         "dec_conv1a": 1, "dec_conv1b": 1, "dec_conv1c": 1, "dec_conv0": 1}
  return sum(2.0 * 9 * ci * co / res[n] for n, ci, co in unet_layers(kind, ic))


# Scale/shift applied to the last convolution so the network output lands in roughly (0.05, 0.9) of
# the transfer domain on benchmark-style inputs (measured once with the reference's PyTorch model by
# tests/golden/make_golden.py --calibrate; fixed here so the bytes never depend on the machine).
_LAST_LAYER = {
  ("base", 9): (0.05, 0.30), ("base", 3): (0.05, 0.30), ("small", 3): (0.10, 0.45),
  ("small", 9): (0.10, 0.45), ("large", 9): (0.06, 0.45), ("large", 3): (0.06, 0.45),
}


def make_weights(kind, ic, seed=0, passthrough=False):
  """He-init weights N(0, 2/(9*cin)), biases U(0, 0.1), rounded to fp16 (what both sides consume).

  passthrough=True: the last three convolutions additionally carry the first three input channels (the
  main image in the transfer domain) straight to the three output channels -- centre tap 1 from the skip
  connection of dec_conv1a, then channel i -> i -- with the rest of those rows scaled to a 2 % perturbation.
  Every other layer keeps its random weights (and its full cost). The network then returns its input
  plus a small perturbation, so a denoised constant-0.5 image stays inside the [0.1, 1] window the
  reference's own test application checks with trained weights (apps/oidnTest.cpp:465-474): these are the
  blobs compiled into the reference build used for integration runs (tools/build_reference_cuda.sh)."""
  rng = np.random.Generator(np.random.PCG64(seed))
  out = {}
  layers = unet_layers(kind, ic)
  for idx, (name, cin, cout) in enumerate(layers):
    w = rng.standard_normal((cout, cin, 3, 3), dtype=np.float32) * np.float32(np.sqrt(2.0 / (9 * cin)))
    b = rng.random((cout,), dtype=np.float32) * np.float32(0.1)
    if passthrough and idx >= len(layers) - 3:
      first = idx == len(layers) - 3            # dec_conv1a: concat(upsampled, input) -> the input starts at cin - ic
      for o in range(3):
        w[o] *= np.float32(0.02 / max(1.0, float(np.abs(w[o]).sum())))
        w[o, (cin - ic if first else 0) + o, 1, 1] = np.float32(1.0)
        b[o] = np.float32(0.0)
    elif idx == len(layers) - 1:
      scale, shift = _LAST_LAYER.get((kind, ic), (0.1, 0.35))
      w = w * np.float32(scale)
      b = b * np.float32(scale) + np.float32(shift)
    out[name + ".weight"] = w.astype(np.float16)
    out[name + ".bias"] = b.astype(np.float16)
  return out


def write_tza(tensors):
  """Serialise {name: ndarray} to TZA v2 bytes (same byte stream as training/tza.py's Writer)."""
  f = io.BytesIO()
  f.write(struct.pack("<HBBQ", TZA_MAGIC, TZA_VERSION[0], TZA_VERSION[1], 0))
  table = []

  def pad():
    off = f.tell()
    f.write(b"\0" * ((off + 63) // 64 * 64 - off))

  for name, t in tensors.items():
    t = np.ascontiguousarray(t)
    layout = "oihw" if t.ndim == 4 else "x"
    dtype = {np.dtype(np.float32): "f", np.dtype(np.float16): "h"}[t.dtype]
    pad()
    table.append((name, t.shape, layout, dtype, f.tell()))
    f.write(t.tobytes())
  pad()
  table_offset = f.tell()
  f.write(struct.pack("<I", len(table)))
  for name, shape, layout, dtype, off in table:
    nb = name.encode()
    f.write(struct.pack("<H", len(nb)) + nb)
    f.write(struct.pack("<B", len(shape)))
    for d in shape:
      f.write(struct.pack("<I", d))
    f.write(layout.encode("ascii") + dtype.encode("ascii"))
    f.write(struct.pack("<Q", off))
  f.seek(4)
  f.write(struct.pack("<Q", table_offset))
  return f.getvalue()


def read_tza(blob):
  """Parse TZA bytes into {name: ndarray}; raises ValueError on a corrupted blob."""
  def need(off, n):
    if off + n > len(blob):
      raise ValueError("invalid or corrupted weights blob")
  need(0, 12)
  magic, major, _minor, table = struct.unpack_from("<HBBQ", blob, 0)
  if magic != TZA_MAGIC:
    raise ValueError("invalid or corrupted weights blob")
  if major != 2:
    raise ValueError("unsupported weights blob version")
  off = table
  need(off, 4)
  (n,) = struct.unpack_from("<I", blob, off); off += 4
  out = {}
  for _ in range(n):
    need(off, 2); (ln,) = struct.unpack_from("<H", blob, off); off += 2
    need(off, ln); name = blob[off:off + ln].decode(); off += ln
    need(off, 1); nd = blob[off]; off += 1
    need(off, 4 * nd); shape = struct.unpack_from("<%dI" % nd, blob, off); off += 4 * nd
    need(off, nd + 1); layout = blob[off:off + nd].decode(); off += nd
    dt = chr(blob[off]); off += 1
    if layout not in ("x", "oihw"):
      raise ValueError("invalid tensor layout")
    if dt not in "fh":
      raise ValueError("invalid tensor data type")
    need(off, 8); (doff,) = struct.unpack_from("<Q", blob, off); off += 8
    dtype = np.float32 if dt == "f" else np.float16
    count = int(np.prod(shape))
    need(doff, count * np.dtype(dtype).itemsize)
    out[name] = np.frombuffer(blob, dtype=dtype, count=count, offset=doff).reshape(shape)
  return out


def model_tza(kind, ic, seed=0, passthrough=False):
  return write_tza(make_weights(kind, ic, seed, passthrough))
