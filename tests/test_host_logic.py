"""CPU tier: host-side logic of the path -- tile scheduler, TZA parser, arena planner -- against the
oracle's restatement of the reference and against first principles."""
import ctypes as C
import struct

import numpy as np
import pytest

from oidn_b200 import api, capi, weights

SIZES = [(1, 1), (2, 2), (16, 16), (89, 257), (720, 1280), (1080, 1920), (2160, 3840), (4096, 4096), (4320, 7680),
         (3001, 80), (80, 3001), (5000, 5000), (777, 12345)]


@pytest.mark.parametrize("H,W", SIZES)
@pytest.mark.parametrize("large", [False, True])
@pytest.mark.parametrize("engines", [1, 2, 4, 8])
def test_tile_planner_equals_reference_restatement(H, W, large, engines, oracle):
  """oracle.plan_tiles restates core/unet_filter.cpp:254-335 line by line; ours is an independent
  implementation of the same search."""
  for max_px in (2160 * 2160, 3840 * 2176, 1000 * 1000):
    ref = oracle.plan_tiles(H, W, large, 1, engines, max_px)
    got, tiles = api.plan_tiles(H, W, large, 1, engines, max_px)
    for k in ("tileH", "tileW", "tilePadH", "tilePadW", "tileCountH", "tileCountW", "tileOverlap", "tileAlignment"):
      assert got[k] == ref[k], (k, got, ref)
    check_tiles(H, W, got, tiles)


def check_tiles(H, W, plan, tiles):
  """Output rectangles partition the image; every input rectangle carries >= overlap context on
  interior edges, sits 16-aligned in the tile buffer and stays inside image and buffer."""
  cover = np.zeros((H, W), np.uint8)
  ov = plan["tileOverlap"]
  for t in tiles:
    assert t["H2"] > 0 and t["W2"] > 0
    cover[t["hDst"]:t["hDst"] + t["H2"], t["wDst"]:t["wDst"] + t["W2"]] += 1
    assert 0 <= t["hSrc"] and t["hSrc"] + t["H1"] <= H and 0 <= t["wSrc"] and t["wSrc"] + t["W1"] <= W
    assert t["hBuf"] >= 0 and t["hBuf"] + t["H1"] <= plan["tileH"] and t["wBuf"] >= 0 and t["wBuf"] + t["W1"] <= plan["tileW"]
    assert t["hBuf"] % 16 == 0 and t["wBuf"] % 16 == 0 and t["hSrc"] % 16 == 0 and t["wSrc"] % 16 == 0
    # context: distance from the output rect to the input rect edge, unless that edge is the image border
    assert t["hDst"] - t["hSrc"] >= (ov if t["hSrc"] > 0 else 0)
    assert (t["hSrc"] + t["H1"]) - (t["hDst"] + t["H2"]) >= (ov if t["hSrc"] + t["H1"] < H else 0)
    assert t["wDst"] - t["wSrc"] >= (ov if t["wSrc"] > 0 else 0)
    assert (t["wSrc"] + t["W1"]) - (t["wDst"] + t["W2"]) >= (ov if t["wSrc"] + t["W1"] < W else 0)
    # output position inside the buffer is consistent with the input placement
    assert t["hOutBuf"] - t["hBuf"] == t["hDst"] - t["hSrc"] and t["wOutBuf"] - t["wBuf"] == t["wDst"] - t["wSrc"]
  assert cover.min() == 1 and cover.max() == 1


def test_survey_tilings():
  """SURVEY.md Appendix A (re-simulation of the reference planner)."""
  p, t = api.plan_tiles(2160, 3840)
  assert (p["tileW"], p["tileH"], p["tileCountW"], p["tileCountH"]) == (2016, 2160, 2, 1)
  assert [(x["wSrc"], x["wDst"], x["W2"]) for x in t] == [(0, 0, 1920), (1824, 1920, 1920)]
  p, _ = api.plan_tiles(4320, 7680, large=True)
  assert (p["tileW"], p["tileH"], p["tileCountW"], p["tileCountH"], p["tileOverlap"]) == (2096, 1600, 4, 3, 112)
  p, _ = api.plan_tiles(4320, 7680, large=True, num_engines=8)
  assert (p["tileW"], p["tileH"], p["tileCountW"] * p["tileCountH"]) == (1472, 1248, 24)
  p, _ = api.plan_tiles(4096, 4096)
  assert (p["tileW"], p["tileH"], p["tileCountW"], p["tileCountH"]) == (2144, 2144, 2, 2)
  p, _ = api.plan_tiles(1080, 1920)
  assert (p["tileW"], p["tileH"], p["tileCountW"] * p["tileCountH"]) == (1920, 1088, 1)
  # B200 default (4K frame = one tile)
  p, _ = api.plan_tiles(2160, 3840, max_tile_pixels=3840 * 2176)
  assert p["tileCountW"] * p["tileCountH"] == 1


@pytest.mark.parametrize("H,W", SIZES)
@pytest.mark.parametrize("large", [False, True])
@pytest.mark.parametrize("units", [1, 2, 3, 4, 8])
def test_min_overlap_planner(H, W, large, units):
  """This backend's default search (tilePolicy=1): same geometry invariants as the reference's plan,
  tile count divisible by the units (engines x shards), never more work per unit than the reference's."""
  for max_px in (7680 * 4352, 3840 * 2176, 1000 * 1000):
    ref, _ = api.plan_tiles(H, W, large, 1, units, max_px, policy=0)
    got, tiles = api.plan_tiles(H, W, large, 1, units, max_px, policy=1)
    check_tiles(H, W, got, tiles)
    n_ref, n_got = ref["tileCountH"] * ref["tileCountW"], got["tileCountH"] * got["tileCountW"]
    if n_ref % units == 0 and ref["tileH"] * ref["tileW"] <= max_px:   # the reference search found a valid plan
      assert n_got % units == 0 and got["tileH"] * got["tileW"] <= max_px
      assert (n_got // units) * got["tileH"] * got["tileW"] <= (n_ref // units) * ref["tileH"] * ref["tileW"]
    assert got["tileOverlap"] == ref["tileOverlap"] and got["tileAlignment"] == ref["tileAlignment"]


def test_min_overlap_tilings_8k():
  """8K frame on 1/2/4/8 B200s: one tile per GPU, recomputed pixels +0 / +2.5 / +7.1 / +12.3 %."""
  want = {1: (1, 1, 7680, 4320), 2: (2, 1, 3936, 4320), 4: (2, 2, 3936, 2256), 8: (4, 2, 2064, 2256)}
  for n, (cw, ch, tw, th) in want.items():
    p, _ = api.plan_tiles(4320, 7680, False, 1, n, 7680 * 4352, policy=1)
    assert (p["tileCountW"], p["tileCountH"], p["tileW"], p["tileH"]) == (cw, ch, tw, th), (n, p)


@pytest.mark.parametrize("H,W", SIZES)
@pytest.mark.parametrize("units", [1, 2, 4, 8])
def test_strip_aware_planner(H, W, units):
  """tilePolicy=2 (opt-in): same geometry invariants; tile widths are costed in whole 128-pixel conv strips,
  so it never needs more strip-rows per unit than the default search."""
  def strips(p):   # full-resolution strip-rows per unit
    return (p["tileCountH"] * p["tileCountW"] // units) * p["tileH"] * -(-p["tileW"] // 128)
  for max_px in (7680 * 4352, 1000 * 1000):
    base, _ = api.plan_tiles(H, W, False, 1, units, max_px, policy=1)
    got, tiles = api.plan_tiles(H, W, False, 1, units, max_px, policy=2)
    check_tiles(H, W, got, tiles)
    assert (got["tileCountH"] * got["tileCountW"]) % units == 0 or (base["tileCountH"] * base["tileCountW"]) % units != 0
    assert got["tileH"] * got["tileW"] <= max_px or base["tileH"] * base["tileW"] > max_px
  p, _ = api.plan_tiles(4320, 7680, False, 1, 8, 7680 * 4352, policy=2)
  assert (p["tileCountW"], p["tileCountH"], p["tileW"], p["tileH"]) == (2, 4, 3936, 1232), p
  q, _ = api.plan_tiles(4320, 7680, False, 1, 8, 7680 * 4352, policy=1)
  assert strips(p) < strips(q)


def parse(blob):
  msg = C.c_char_p()
  buf = (C.c_char * max(len(blob), 1)).from_buffer_copy(blob or b"\0")
  n = capi.lib().oidnb200ParseTZA(buf, len(blob), C.byref(msg))
  return n, (msg.value or b"").decode()


def test_tza_parser_accepts_reference_format():
  for kind, ic, ntens in (("base", 9, 32), ("small", 3, 32), ("large", 9, 38)):
    blob = weights.model_tza(kind, ic)
    assert parse(blob) == (ntens, "")
    back = weights.read_tza(blob)
    for name, t in weights.make_weights(kind, ic).items():
      np.testing.assert_array_equal(back[name], t)


def test_tza_parser_rejects_malformed_blobs():
  """Messages and error class of core/tza.cpp:27-103 (all Error::InvalidOperation = 3)."""
  blob = bytearray(weights.model_tza("small", 3))
  assert parse(b"") == (-3, "invalid or corrupted weights blob")
  assert parse(bytes(blob[:11]))[0] == -3
  bad = bytearray(blob); bad[0] ^= 0xFF
  assert parse(bytes(bad)) == (-3, "invalid or corrupted weights blob")
  bad = bytearray(blob); bad[2] = 3
  assert parse(bytes(bad)) == (-3, "unsupported weights blob version")
  bad = bytearray(blob); struct.pack_into("<Q", bad, 4, len(blob) + 1)
  assert parse(bytes(bad))[0] == -3
  assert parse(bytes(blob[:-9]))[0] == -3                    # truncated table
  table = struct.unpack_from("<Q", blob, 4)[0]
  # first table entry: u32 n, u16 len, name, u8 ndims, dims, layout, dtype, u64 offset
  ln = struct.unpack_from("<H", blob, table + 4)[0]
  nd_off = table + 6 + ln
  nd = blob[nd_off]
  layout_off = nd_off + 1 + 4 * nd
  bad = bytearray(blob); bad[layout_off] = ord("q")
  assert parse(bytes(bad)) == (-3, "invalid tensor layout")
  bad = bytearray(blob); bad[layout_off + nd] = ord("d")
  assert parse(bytes(bad)) == (-3, "invalid tensor data type")
  bad = bytearray(blob); struct.pack_into("<Q", bad, layout_off + nd + 1, len(blob) - 2)
  assert parse(bytes(bad))[0] == -3                           # tensor data runs past the blob


def test_tza_parser_rejects_wrapping_dimensions():
  """A crafted table entry whose element count wraps size_t (65536^4 = 2^64) or whose dims exceed INT_MAX must not
  pass the bounds check with a wrapped byte size; ranks other than 1 / 4 are refused before the dims are read."""
  def blob_with(dims, layout, dtype=b"h"):
    name = b"enc_conv0.weight"
    entry = struct.pack("<H", len(name)) + name + struct.pack("<B", len(dims)) + b"".join(struct.pack("<I", d) for d in dims)
    entry += layout + dtype + struct.pack("<Q", 12)
    return struct.pack("<HBBQ", 0x41D7, 2, 0, 12) + struct.pack("<I", 1) + entry
  assert parse(blob_with([1, 1, 1, 1], b"oihw")) == (1, "")        # sanity: a well-formed one-element tensor parses
  assert parse(blob_with([65536] * 4, b"oihw")) == (-3, "invalid or corrupted weights blob")
  assert parse(blob_with([2**31, 2**31, 2, 2], b"oihw")) == (-3, "invalid or corrupted weights blob")
  assert parse(blob_with([2**31 - 1, 2**31 - 1, 2**31 - 1, 8], b"oihw", b"f")) == (-3, "invalid or corrupted weights blob")
  assert parse(blob_with([4, 4], b"xy")) == (-3, "invalid tensor layout")
  assert parse(blob_with([1] * 255, b"x" * 255)) == (-3, "invalid tensor layout")


def plan_arena(sizes, first, last):
  n = len(sizes)
  off = (C.c_size_t * n)()
  total = capi.lib().oidnb200PlanArena(n, (C.c_size_t * n)(*sizes), (C.c_int * n)(*first), (C.c_int * n)(*last), off)
  return total, list(off)


def test_arena_planner_packs_lifetimes():
  # a chain a->b->c->d where only neighbours overlap in time: two slots suffice
  total, off = plan_arena([1000, 1000, 1000, 1000], [0, 1, 2, 3], [1, 2, 3, 4])
  assert total != 2**64 - 1 and total <= 2 * 1024
  assert all(o % 256 == 0 for o in off)
  # random lifetimes: never overlapping in space and time at once (validated inside), never worse than the sum
  rng = np.random.default_rng(0)
  for _ in range(50):
    n = int(rng.integers(2, 40))
    sizes = [int(s) for s in rng.integers(1, 1 << 20, n)]
    first = [int(v) for v in rng.integers(0, 30, n)]
    last = [f + int(v) for f, v in zip(first, rng.integers(0, 10, n))]
    total, off = plan_arena(sizes, first, last)
    assert total != 2**64 - 1, "planner produced overlapping live allocations"
    assert total <= sum((s + 255) // 256 * 256 for s in sizes)
    live_peak = max(sum(s for s, f, l in zip(sizes, first, last) if f <= t <= l) for t in range(41))
    assert total >= live_peak


def test_unet_arena_is_close_to_the_live_peak():
  """Tensors of the fused base UNet at a 1920x1088 tile: the planner's total must
  be within 15 % of the peak live set (in + d2b + d1a ~ 0.40 GB)."""
  px = 1920 * 1088
  def b(c, s): return px // (s * s) * c * 2
  # (bytes, producer op, last consumer op) in graph order: input=0, convs 1..16, output=17
  t = [(b(16, 1), 0, 14), (b(32, 1), 1, 2), (b(32, 2), 2, 12), (b(48, 4), 3, 10), (b(64, 8), 4, 8), (b(80, 16), 5, 6),
       (b(96, 16), 6, 7), (b(96, 16), 7, 8), (b(112, 8), 8, 9), (b(112, 8), 9, 10), (b(96, 4), 10, 11), (b(96, 4), 11, 12),
       (b(64, 2), 12, 13), (b(64, 2), 13, 14), (b(64, 1), 14, 15), (b(32, 1), 15, 16), (b(16, 1), 16, 17)]
  total, _ = plan_arena([x[0] for x in t], [x[1] for x in t], [x[2] for x in t])
  peak = max(sum(s for s, f, l in t if f <= op <= l) for op in range(18))
  assert peak <= total <= 1.15 * peak, (total, peak)


def test_bench_reference_arm_contract():
  """`bench.py --impl reference` (the driver's reference arm): the oracle port on host cores, one JSON line with the
  contract's keys -- `impl`, the metric / unit of the product arm, a `cpu_baseline` describing the run and an `e2e`
  that repeats the value with zero transfer bytes. A tiny frame keeps it to a second; no GPU involved."""
  import json
  import os
  import subprocess
  import sys
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--width", "96", "--height", "64",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=root)
  assert out.returncode == 0, out.stderr[-500:]
  line = json.loads(out.stdout.strip().splitlines()[-1])
  assert line["impl"] == "reference" and line["unit"] == "Mpix/s" and line["higher_is_better"] is True
  assert line["n_gpus"] == 1 and line["steps"] == 1 and line["vs_baseline"] is None and line["value"] > 0
  assert line["config"]["width"] == 96 and line["config"]["height"] == 64 and "workload" in line["config"]
  cb = line["cpu_baseline"]
  assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
  assert line["e2e"] == {"value": line["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_bench_reference_arm_under_torchrun_prints_once():
  """N > 1: the driver launches the reference arm like the product arm (torchrun, one rank per GPU); rank 0 alone runs
  and prints the line, the other ranks exit 0 without work."""
  import json
  import os
  import socket
  import subprocess
  import sys
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
  out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--width", "96", "--height", "64", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300, cwd=root)
  assert out.returncode == 0, out.stderr[-500:]
  lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
  assert len(lines) == 1
  line = json.loads(lines[0])
  assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["value"] > 0
