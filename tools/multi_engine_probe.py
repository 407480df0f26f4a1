"""One process, one device object over N GPUs (oidnb200NewCUDADevice(ids, streams, N)): frame time of the 8K base-UNet
frame by where the frame lives and how it is reached. usage: python tools/multi_engine_probe.py [N] [W H]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oidn_b200 import api, capi, synth, weights  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (7680, 4320)
K = 10
tza = weights.model_tza("base", 9, seed=0)
imgs = synth.benchmark_images(W, H, hdr=True, seed=1)


def run(label, staging, where, user_streams=False, policy=None):
  streams = None
  if user_streams:
    streams = []
    for g in range(N):
      with torch.cuda.device(g):
        streams.append(torch.cuda.Stream())
  dev = api.Device(tuple(range(N)), streams=[s.cuda_stream for s in streams] if streams else None).commit()
  dev.set("staging", staging)
  if policy is not None:
    dev.set("tilePolicy", policy)
  f = dev.new_filter("RT")
  keep = []
  if where == "gpu0":
    with torch.cuda.device(0):
      t = {k: torch.from_numpy(v).cuda() for k, v in imgs.items()}
      out = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
    for k, v in t.items():
      f.set_image(k, v)
    f.set_image("output", out)
    keep = [t, out]
  else:
    hb = {k: dev.new_buffer(v.nbytes, api.STORAGE_HOST) for k, v in imgs.items()}
    ho = dev.new_buffer(W * H * 12, api.STORAGE_HOST)
    for k, v in imgs.items():
      hb[k].write(v)
      f.set_image(k, hb[k], capi.FORMAT_FLOAT3, W, H)
    f.set_image("output", ho, capi.FORMAT_FLOAT3, W, H)
    keep = [hb, ho]
  f.set("hdr", True); f.set_data("weights", tza); f.commit()
  for _ in range(3):
    f.execute_async()
  dev.sync()
  info = f.info()
  t0 = time.perf_counter()
  for _ in range(K):
    f.execute_async()
  t_enq = time.perf_counter() - t0
  dev.sync()
  dt = (time.perf_counter() - t0) / K * 1e3
  # one frame alone (latency)
  t0 = time.perf_counter(); f.execute(); lat = (time.perf_counter() - t0) * 1e3
  print("%-58s %7.3f ms/frame pipelined, %7.3f ms alone, enqueue %.3f ms/frame, staged=%d, tiles %dx%d of %dx%d" % (
    label, dt, lat, t_enq / K * 1e3, info["staged"], info["tileCountW"], info["tileCountH"], info["tileW"], info["tileH"]), flush=True)
  f.release(); dev.release()
  del keep


print("N=%d GPUs, %dx%d, wall clock over %d frames" % (N, W, H, K), flush=True)
run("frame in GPU 0 HBM, staged, own streams", -1, "gpu0")
run("frame in GPU 0 HBM, staged, caller streams (eager join)", -1, "gpu0", user_streams=True)
run("frame in GPU 0 HBM, in place (P2P loads/stores)", 0, "gpu0")
run("frame in GPU 0 HBM, staged, tilePolicy 2", -1, "gpu0", policy=2)
run("frame in pinned host memory, staged", -1, "host")
run("frame in pinned host memory, zero copy", 0, "host")


def raw_pcie(n_gpus, mb=512, reps=5):
  """What the platform gives: concurrent pinned host -> device (and back) copies on n GPUs, aggregate GB/s."""
  host = [torch.empty(mb << 20, dtype=torch.uint8).pin_memory() for _ in range(n_gpus)]
  devb, streams = [], []
  for g in range(n_gpus):
    with torch.cuda.device(g):
      devb.append(torch.empty(mb << 20, dtype=torch.uint8, device="cuda"))
      streams.append(torch.cuda.Stream())
  res = {}
  for name, fn in (("H2D", lambda g: devb[g].copy_(host[g], non_blocking=True)), ("D2H", lambda g: host[g].copy_(devb[g], non_blocking=True))):
    for _ in range(2):
      for g in range(n_gpus):
        with torch.cuda.device(g), torch.cuda.stream(streams[g]):
          fn(g)
    for g in range(n_gpus):
      torch.cuda.synchronize(g)
    t0 = time.perf_counter()
    for _ in range(reps):
      for g in range(n_gpus):
        with torch.cuda.device(g), torch.cuda.stream(streams[g]):
          fn(g)
    for g in range(n_gpus):
      torch.cuda.synchronize(g)
    res[name] = n_gpus * reps * (mb << 20) / (time.perf_counter() - t0) / 1e9
  return res


for n in sorted({1, 2, min(4, N), N}):
  if n <= N:
    r = raw_pcie(n)
    print("raw pinned-memory copies on %d GPU(s) at once: H2D %.1f GB/s aggregate (%.1f per GPU), D2H %.1f GB/s (%.1f per GPU)"
          % (n, r["H2D"], r["H2D"] / n, r["D2H"], r["D2H"] / n), flush=True)
