#!/usr/bin/env python
"""All five BASELINE.json configs on ONE B200 through the filter C ABI (bench.py measures configs[1]
only; the others are parity-test cases, and this tool records what they cost).

  python tools/bench_configs.py [--steps K] [--parity] [--only 1,5]

Per config: device-resident ms/frame (CUDA events on the device's stream, K frames after warm-up),
Mpix/s, the tile plan, the conv kernel's algorithmic TFLOP/s from per-op events, and with --parity
the max|err|/peak and PSNR against the CPU oracle on the full frame (configs whose oracle run fits
in about a minute). Config 5 is a 64-frame stream: frames are enqueued back to back with
oidnb200ExecuteFilterAsync and one sync at the end (apps/oidnTest.cpp:873-935 usage); per-frame
device latency comes from events between frames, host enqueue cost from the wall clock.
One JSON line per config on stdout.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from oidn_b200 import api, synth, weights  # noqa: E402

CONFIGS = {
  1: dict(name="RT hdr+alb+nrm 1920x1080 (oidnBenchmark RT.hdr_alb_nrm.1920x1080)", type="RT", W=1920, H=1080, model=("base", 9),
          hdr=True, aux=True, quality=api.QUALITY_HIGH, oracle_s=3),
  2: dict(name="RT hdr+alb+nrm 3840x2160 quality=high", type="RT", W=3840, H=2160, model=("base", 9), hdr=True, aux=True,
          quality=api.QUALITY_HIGH, oracle_s=10),
  3: dict(name="RT hdr+calb+cnrm 7680x4320 quality=high (large UNet), one GPU", type="RT", W=7680, H=4320, model=("large", 9),
          hdr=True, aux=True, clean_aux=True, quality=api.QUALITY_HIGH, oracle_s=120),
  4: dict(name="RTLightmap hdr 4096x4096", type="RTLightmap", W=4096, H=4096, model=("base", 3), hdr=True, aux=False,
          quality=api.QUALITY_HIGH, oracle_s=20),
  41: dict(name="RTLightmap directional 4096x4096", type="RTLightmap", W=4096, H=4096, model=("base", 3), hdr=False, aux=False,
           directional=True, quality=api.QUALITY_HIGH, oracle_s=20),
  5: dict(name="RT ldr color-only 1280x720 quality=fast, 64-frame stream", type="RT", W=1280, H=720, model=("small", 3), hdr=False,
          aux=False, quality=api.QUALITY_FAST, stream=64, oracle_s=1),
}


def make_filter(dev, cfg, t, out, tza):
  f = dev.new_filter(cfg["type"])
  for k, v in t.items():
    f.set_image(k, v)
  f.set_image("output", out)
  if cfg.get("directional"):
    f.set("directional", True)
  else:
    f.set("hdr", bool(cfg["hdr"]))
  if cfg.get("clean_aux"):
    f.set("cleanAux", True)
  f.set("quality", cfg["quality"])
  f.set_data("weights", tza)
  f.commit()
  return f


def run_config(cid, cfg, args, torch):
  W, H = cfg["W"], cfg["H"]
  kind, ic = cfg["model"]
  tza = weights.model_tza(kind, ic, seed=0)
  if cfg.get("directional"):
    imgs = {"color": synth.benchmark_images(W, H, hdr=False, albedo=False, normal=False, seed=1)["color"] * 2 - 1}
  else:
    imgs = synth.benchmark_images(W, H, hdr=cfg["hdr"], albedo=cfg["aux"], normal=cfg["aux"], seed=1)
  stream = torch.cuda.Stream()
  res = {"config": cid, "name": cfg["name"], "width": W, "height": H, "model": "%s ic=%d" % (kind, ic)}
  with torch.cuda.stream(stream):
    dev = api.Device((0,), streams=[stream.cuda_stream]).commit()
    if args.graph:
      dev.set("graph", 1)
    t = {k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in imgs.items()}
    out = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
    f = make_filter(dev, cfg, t, out, tza)
    info = f.info()
    res["tiles"] = "%dx%d of %dx%d" % (info["tileCountW"], info["tileCountH"], info["tileW"], info["tileH"])
    res["large_model"] = bool(info["largeModel"])
    K = cfg.get("stream") or args.steps
    for _ in range(5):
      f.execute_async()
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    w0 = time.perf_counter()
    evs[0].record(stream)
    for i in range(K):
      f.execute_async()
      evs[i + 1].record(stream)
    w_enq = time.perf_counter()
    torch.cuda.synchronize()
    w1 = time.perf_counter()
    per = np.array([evs[i].elapsed_time(evs[i + 1]) for i in range(K)])
    ms = evs[0].elapsed_time(evs[K]) / K
    res.update(frames=K, ms_per_frame=round(ms, 4), mpix_s=round(W * H / ms / 1e3, 1),
               frame_ms_p50=round(float(np.percentile(per, 50)), 4), frame_ms_p99=round(float(np.percentile(per, 99)), 4),
               frame_ms_max=round(float(per.max()), 4),
               host_enqueue_ms_per_frame=round((w_enq - w0) * 1e3 / K, 4), wall_ms_per_frame=round((w1 - w0) * 1e3 / K, 4))
    if cfg.get("stream"):
      # latency of ONE frame submitted to an idle device (execute + sync, wall clock)
      lat = []
      for _ in range(20):
        torch.cuda.synchronize()
        a = time.perf_counter(); f.execute(); lat.append((time.perf_counter() - a) * 1e3)
      res["single_frame_latency_ms"] = {"p50": round(float(np.percentile(lat, 50)), 4), "min": round(min(lat), 4), "max": round(max(lat), 4)}
    # per-op device times -> conv TFLOP/s (algorithmic, unpadded channels, image pixels)
    dev.set("profile", 1)
    f.execute(); f.profile()
    for _ in range(min(K, 10)):
      f.execute_async()
    torch.cuda.synchronize()
    prof = f.profile()
    dev.set("profile", 0)
    n = min(K, 10)
    conv_ms = sum(m for _, kind_, _, m in prof if kind_ == 0) / n
    res["conv_ms"] = round(conv_ms, 4)
    res["conv_tflops"] = round(weights.flops_per_pixel(kind, ic) * W * H / (conv_ms * 1e-3) / 1e12, 1)
    res["elementwise_ms"] = round(sum(m for _, kind_, _, m in prof if kind_ != 0) / n, 4)
    # in-frame conv intervals (%globaltimer stamps, frames one at a time): what the frame really spends in convs
    dev.set("profile", 2)
    f.execute(); f.profile()
    for _ in range(n):
      f.execute_async()
    torch.cuda.synchronize()
    prof2 = f.profile()
    dev.set("profile", 0)
    res["conv_union_ms"] = round(sum(m for _, kind_, _, m in prof2 if kind_ == 3) / n, 4)
    res["conv_layers_in_frame_us"] = {nm: round(m / n * 1e3, 1) for nm, kind_, _, m in prof2 if kind_ == 0 and m > 0}
    got = out.cpu().numpy()
    f.release(); dev.release()
  res["finite"] = bool(np.isfinite(got).all())
  if args.parity and cfg["oracle_s"] <= args.parity_budget:
    import oracle as orc
    ref = np.zeros((H, W, 3), np.float32)
    kw = dict(output=ref, hdr=bool(cfg["hdr"]), filter=cfg["type"], directional=bool(cfg.get("directional")))
    kw.update({k: np.ascontiguousarray(v) for k, v in imgs.items()})
    t0 = time.time()
    orc.filter_execute(tza, **kw)
    peak = float(np.abs(ref).max())
    err = float(np.abs(got - ref).max()) / peak
    psnr = float(20 * np.log10(peak / np.sqrt(np.mean((got - ref) ** 2))))
    res["parity"] = {"max_abs_err_over_peak": float("%.3e" % err), "psnr_db": round(psnr, 1), "oracle_s": round(time.time() - t0, 1),
                     "pass": bool(err <= 1e-2 and psnr >= 50.0)}
  return res


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--steps", type=int, default=20)
  ap.add_argument("--parity", action="store_true")
  ap.add_argument("--parity-budget", type=int, default=30, help="skip the oracle for configs expected to take longer (s)")
  ap.add_argument("--only", default="")
  ap.add_argument("--graph", action="store_true", help="device parameter graph=1 (CUDA-graph replay of the frame)")
  args = ap.parse_args()
  import torch
  if not torch.cuda.is_available():
    raise SystemExit("needs a GPU")
  only = [int(x) for x in args.only.split(",") if x]
  for cid, cfg in CONFIGS.items():
    if only and cid not in only:
      continue
    print(json.dumps(run_config(cid, cfg, args, torch)), flush=True)


if __name__ == "__main__":
  main()
