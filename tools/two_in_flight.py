"""Device-resident frames on ONE GPU: one filter back to back against two device/stream/filter sets alternating
(frame f+1's autoexposure + input process and its first convs overlap the low-resolution tail of frame f).
Run: python tools/two_in_flight.py [W H K]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oidn_b200 import api, synth, weights  # noqa: E402

W, H, K = (int(a) for a in (sys.argv[1:4] + ["3840", "2160", "40"][len(sys.argv) - 1:]))
tza = weights.model_tza("base", 9, seed=0)
imgs = synth.benchmark_images(W, H, hdr=True, seed=1)
nb = W * H * 12


def make(n):
  sets = []
  for _ in range(n):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
      d = api.Device((0,), streams=[s.cuda_stream]).commit()
      bufs = {k: d.new_buffer(nb) for k in ("color", "albedo", "normal", "output")}
      for k, v in imgs.items():
        bufs[k].write(v)
      f = d.new_filter("RT")
      for k, b in bufs.items():
        f.set_image(k, b, api.capi.FORMAT_FLOAT3, W, H)
      f.set("hdr", True); f.set("quality", api.QUALITY_HIGH); f.set_data("weights", tza); f.commit()
    sets.append((s, d, bufs, f))
  return sets


def run(sets, frames):
  n = len(sets)
  for i in range(6):
    with torch.cuda.stream(sets[i % n][0]):
      sets[i % n][3].execute_async()
  torch.cuda.synchronize()
  e0, e1, join = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event()
  e0.record(sets[0][0])
  for s, *_ in sets[1:]:
    s.wait_event(e0)
  for i in range(frames):
    with torch.cuda.stream(sets[i % n][0]):
      sets[i % n][3].execute_async()
  for s, *_ in sets[1:]:
    join.record(s); sets[0][0].wait_event(join)
  e1.record(sets[0][0])
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / frames


outs = []
for n in (1, 2, 3, 1, 2):
  sets = make(n)
  ms = run(sets, K)
  o = np.zeros((H, W, 3), np.float32); sets[-1][2]["output"].read(o); outs.append(o)
  print("%d set(s): %.4f ms/frame = %.1f Mpix/s" % (n, ms, W * H / ms / 1e3), flush=True)
  for s, d, bufs, f in sets:
    f.release()
    for b in bufs.values():
      b.release()
    d.release()
  time.sleep(0.5)
print("outputs identical:", all(np.array_equal(o.view(np.uint32), outs[0].view(np.uint32)) for o in outs))
