#!/usr/bin/env python
"""Benchmark of the B200 denoising path (BASELINE.json metric: Mpix/s and ms/frame, RT hdr+alb+nrm).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one frame through oidnb200ExecuteFilterAsync (autoexposure + input process + 16 fused
convolutions + output process per tile). N=1: BASELINE.json configs[1] (RT filter, HDR color +
albedo + normal, 3840x2160, quality=high -> base UNet; there is no large variant for this feature
set). N>1 (torchrun, one rank per GPU): the 7680x4320 frame of the same filter,
tile-sharded across the ranks; every rank holds the inputs of its own tiles (tile + overlap), the
autoexposure bin array is completed with one NCCL all-reduce (~0.5 MB), output rectangles are
assembled in rank 0's buffer over NVLink (CUDA IPC peer mappings, copy engines) and a 4-byte
all-reduce joins the frame. `frame_on_rank0` reports the variant where rank 0 holds the whole frame.

Prints ONE JSON line (rank 0). `value` is device-resident throughput (inputs already in HBM), `e2e`
is the same metric through the public API with pinned host buffers (H2D + D2H inside the timed
region). `--impl reference` times the CPU restatement of the reference's CPU device (oracle/) on the
host cores: the reference CPU device itself needs ISPC + oneTBB and cannot be built in this image.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from oidn_b200 import synth, weights  # noqa: E402

METRIC = "Mpix/s, RT hdr+alb+nrm (ms/frame in ms_per_step)"
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def load_peaks():
  p = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(p):
    try:
      d = json.load(open(p))
      out = {k: float(d[k]) for k in ("hbm_gbs", "bf16_tflops") if k in d}
      out["bf16_tflops_sustained"] = float(d.get("bf16_tflops_sustained", out.get("bf16_tflops", 0)))
      if len(out) == 3:
        return out, "measured (MEASURED_PEAKS.json)"
    except Exception:
      pass
  return dict(FALLBACK_PEAKS), "fallback (B200_PROFILING.md)"


class ClockSampler:
  """Samples SM clocks, power and clock-event (throttle) reasons while the timed region runs: NVML polled in
  this process every ~2 ms (the timed region of a 4K bench is tens of ms), nvidia-smi -lms 100 as fallback."""
  Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
       "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
  NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

  def __init__(self, index):
    self.rows, self.proc, self.nvml, self.run = [], None, None, True
    try:
      import pynvml
      pynvml.nvmlInit()
      h = None
      try:
        import torch
        h = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + str(torch.cuda.get_device_properties(index).uuid))
      except Exception:
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
      self.nvml, self.h = pynvml, h
      self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
      self.t = threading.Thread(target=self._poll, daemon=True); self.t.start()
      return
    except Exception:
      self.nvml = None
    try:
      self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                    "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
      self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
    except Exception:
      self.proc = None

  def _poll(self):
    n = self.nvml
    bits = [n.nvmlClocksEventReasonHwSlowdown, n.nvmlClocksEventReasonHwThermalSlowdown,
            n.nvmlClocksEventReasonSwThermalSlowdown, n.nvmlClocksEventReasonSwPowerCap]
    while self.run:
      try:
        sm = n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        try:
          pw = n.nvmlDeviceGetPowerUsage(self.h) / 1000.0
        except Exception:
          pw = 0.0
        self.rows.append((time.time(), [str(sm), str(self.max_sm), "%.1f" % pw] + ["Active" if r & b else "Not Active" for b in bits]))
      except Exception:
        pass
      time.sleep(0.002)

  def _read(self):
    for line in self.proc.stdout:
      self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

  def window(self, t0, t1):
    rows = [r for t, r in list(self.rows) if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
    if not rows:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
    sm = sorted(int(float(r[0])) for r in rows if r[0].replace(".", "").isdigit())
    reasons = [n for i, n in enumerate(self.NAMES) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in rows)]
    pw = [float(r[2]) for r in rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
    return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(float(rows[0][1])) if rows[0][1].replace(".", "").isdigit() else None,
            "power_w": round(max(pw), 1) if pw else None, "samples": len(rows), "reasons": reasons,
            "source": "nvml, 2 ms poll" if self.nvml else "nvidia-smi -lms 100"}

  def stop(self):
    self.run = False
    if self.proc:
      self.proc.terminate()


def cpu_port_throughput(tza, W, H, budget_s=20.0):
  """Times the oracle (CPU restatement of the reference CPU device) on crops of the workload.
  Only place besides tests/ and smoke() that executes oracle/ (as the reported CPU baseline)."""
  sys.path.insert(0, os.path.join(ROOT, "oracle"))
  import oracle as orc
  cores = orc.lib().oro_set_num_threads(os.cpu_count() or 1)   # the threads actually used
  best = None
  for w, h in ((480, 272), (960, 544), (1920, 1080), (W, H)):
    if w > W or h > H:
      break
    imgs = synth.benchmark_images(w, h, hdr=True, seed=1)
    out = np.zeros((h, w, 3), np.float32)
    t0 = time.time()
    orc.filter_execute(tza, color=imgs["color"], albedo=imgs["albedo"], normal=imgs["normal"], output=out, hdr=True)
    dt = time.time() - t0
    best = {"value": round(w * h / dt / 1e6, 4), "unit": "Mpix/s", "cores": cores, "kind": "port",
            "sample": "oracle/oidn_oracle.c (OpenMP fp32 restatement of devices/cpu) on a %dx%d frame of the same workload, %.1f s" % (w, h, dt),
            "seconds": round(dt, 2)}
    if dt * 4.5 > budget_s:   # the next size is 4x the pixels
      break
  return best


def run_reference(args, rank):
  if rank != 0:
    return
  tza = weights.model_tza("base", 9, seed=0)
  W, H = args.width, args.height
  sys.path.insert(0, os.path.join(ROOT, "oracle"))
  import oracle as orc
  # torchrun exports OMP_NUM_THREADS=1 to its workers: ask for all host cores, report what is in effect
  cores = orc.lib().oro_set_num_threads(os.cpu_count() or 1)
  # one step = the WHOLE frame of the workload when steps + warm-up of it fit the time budget (300 s: the default
  # 20 + 5 steps of the 4K frame do on a 16-core host), else the largest crop that does (a bounded sample)
  probe = cpu_port_throughput(tza, 480, 272, budget_s=1.0)
  budget_px = probe["value"] * 1e6 * 300.0 / max(args.steps + args.warmup, 1)
  if budget_px >= W * H:
    w, h = W, H
  else:
    px = max(480 * 272, int(budget_px))
    w = min(W, max(480, int((px * 16 / 9) ** 0.5) // 16 * 16)); h = min(H, max(272, w * 9 // 16 // 16 * 16))
  imgs = synth.benchmark_images(w, h, hdr=True, seed=1)
  out = np.zeros((h, w, 3), np.float32)
  run = lambda: orc.filter_execute(tza, color=imgs["color"], albedo=imgs["albedo"], normal=imgs["normal"], output=out, hdr=True)
  for _ in range(args.warmup):
    run()
  t0 = time.time()
  for _ in range(args.steps):
    run()
  dt = (time.time() - t0) / args.steps
  val = w * h / dt / 1e6
  sample = ("oracle port (C/OpenMP restatement of the reference CPU device; the ISPC+TBB device is unbuildable here), %s per step"
            % ("the whole %dx%d frame" % (w, h) if (w, h) == (W, H) else "a %dx%d crop of the %dx%d frame" % (w, h, W, H)))
  print(json.dumps({
    "impl": "reference", "metric": METRIC, "value": round(val, 4), "unit": "Mpix/s", "n_gpus": args.gpus, "steps": args.steps,
    "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak",
    "vs_baseline": None, "dtype": "f32", "data": "synthetic",
    "config": workload_config(args, args.gpus),
    "cpu_baseline": {"value": round(val, 4), "unit": "Mpix/s", "cores": cores, "kind": "port", "sample": sample},
    "e2e": {"value": round(val, 4), "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def workload_config(args, n):
  W, H = frame_size(args, n)
  return {"workload": "RT filter, HDR color+albedo+normal fp32, %dx%d, quality=high (base UNet, 16 convs), autoexposure on, "
                      "oidnBenchmark LCG inputs, synthetic He-init TZA weights" % (W, H),
          "width": W, "height": H, "l2": "inputs and every intermediate tensor are larger than the 126 MB L2 (no flush needed)",
          "sharding": "single GPU" if n == 1 else "tile-sharded across %d ranks; every rank holds its tiles' inputs (tile + overlap), "
                      "output assembled in rank 0's HBM by NVLink peer copies" % n}


def frame_size(args, n):
  if args.width and args.height and args.explicit_size:
    return args.width, args.height
  return (3840, 2160) if n == 1 else (7680, 4320)


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=20)
  ap.add_argument("--warmup", type=int, default=5)
  ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
  ap.add_argument("--width", type=int, default=0)
  ap.add_argument("--height", type=int, default=0)
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--no-e2e", action="store_true")
  ap.add_argument("--no-rank0", action="store_true", help="N>1: skip the secondary frame-on-rank-0 measurement")
  ap.add_argument("--no-8k", action="store_true", help="N=1: skip the 8K base-UNet and large-UNet lines")
  ap.add_argument("--no-kernel-to-beat", action="store_true", help="N=1: do not run the reference's own CUDA device (baseline/_ref)")
  ap.add_argument("--no-single-process", action="store_true", help="N>1: skip the one-process multi-engine device measurement")
  args = ap.parse_args()
  args.explicit_size = bool(args.width and args.height)
  rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
  local_rank = int(os.environ.get("LOCAL_RANK", "0"))
  args.warmup = max(args.warmup, 3 if args.impl == "ours" else 0)
  if not args.explicit_size:
    args.width, args.height = frame_size(args, max(world, args.gpus))
  if args.impl == "reference":
    args.width, args.height = frame_size(args, max(world, args.gpus)) if not args.explicit_size else (args.width, args.height)
    run_reference(args, rank)
    return

  import torch
  from oidn_b200 import api
  if not torch.cuda.is_available():
    raise SystemExit("bench.py needs a GPU: the product path has no CPU fallback")
  if world > 1:
    from oidn_b200 import sharded
    sharded.bench_main(args, rank, world, local_rank)
    return

  torch.cuda.set_device(local_rank)
  W, H = args.width, args.height
  K, Wm = args.steps, args.warmup
  peaks, peaks_src = load_peaks()
  tza = weights.model_tza("base", 9, seed=0)
  imgs = synth.benchmark_images(W, H, hdr=True, seed=1)

  stream = torch.cuda.Stream()
  sampler = ClockSampler(local_rank)
  with torch.cuda.stream(stream):
    dev = api.Device((local_rank,), streams=[stream.cuda_stream]).commit()
    t = {k: torch.from_numpy(v).cuda() for k, v in imgs.items()}
    out = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
    f = dev.new_filter("RT")
    for k, v in t.items():
      f.set_image(k, v)
    f.set_image("output", out)
    f.set("hdr", True); f.set("quality", api.QUALITY_HIGH); f.set_data("weights", tza)
    f.commit()
    info = f.info()
    ntiles = info["tileCountH"] * info["tileCountW"]

    # ---- device-resident throughput ---------------------------------------------------------
    for _ in range(Wm):
      f.execute_async()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.time()
    e0.record(stream)
    for _ in range(K):
      f.execute_async()
    e1.record(stream)
    torch.cuda.synchronize()
    t1 = time.time()
    ms = e0.elapsed_time(e1) / K
    clocks = sampler.window(t0, t1)

    # ---- where the frame's time goes -------------------------------------------------------------------------
    # (1) in-frame conv intervals: every conv grid stamps %globaltimer when its first CTA gets past the wait for the
    #     previous grid and when its last CTA exits (device parameter profile=2). Launches overlap through programmatic
    #     dependent launch (prologues run under the previous grid's tail), so the frame's conv time is the UNION of
    #     the intervals -- by construction not more than the frame (same K back-to-back frames as the timed region)
    dev.set("profile", 2)
    f.execute(); f.profile()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record(stream)
    for _ in range(K):
      f.execute_async()            # back to back, as in the timed region: the stamps stay on the device until profile()
    s1.record(stream)
    torch.cuda.synchronize()
    stamped_ms = s0.elapsed_time(s1) / K   # this pass's own frame time (the board warms up from pass to pass)
    prof2 = f.profile()
    conv_union_ms = sum(m for _, kind, _, m in prof2 if kind == 3) / K
    conv_layers = {n: round(m / K, 4) for n, kind, _, m in prof2 if kind == 0}
    conv_launches = sum(n for _, kind, n, _ in prof2 if kind == 0) // K
    dev.set("profile", 0)

    # ---- the same frames for >= 3 s (the short run above is a burst at full clocks; this is the regime a stream
    # of frames settles in: power-capped clocks) -----------------------------------------------------------------
    n_long = max(K, int(3000.0 / ms) + 1)
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    tl0 = time.time()
    l0.record(stream)
    for _ in range(n_long):
      f.execute_async()
    l1.record(stream)
    torch.cuda.synchronize()
    tl1 = time.time()
    ms_long = l0.elapsed_time(l1) / n_long
    clocks_long = sampler.window(tl0, tl1)

    # (2) CUDA events around every op (serialises the launches: used for the elementwise passes only)
    dev.set("profile", 1)
    f.execute(); f.profile()
    for _ in range(K):
      f.execute_async()
    torch.cuda.synchronize()
    prof = f.profile()
    dev.set("profile", 0)
    in_ms = sum(m for _, kind, _, m in prof if kind == 1) / K
    out_ms = sum(m for _, kind, _, m in prof if kind == 2) / K
    in_launches = sum(n for _, kind, n, _ in prof if kind == 1) // K
    out_launches = sum(n for _, kind, n, _ in prof if kind == 2) // K
    # per frame: autoexposure bins + fold, then per tile the input process, the conv launches (a fused conv pair is
    # one launch; counted from the kernels' own in-frame stamps) and the output process when it is a pass of its own
    launches = K * (2 + in_launches + conv_launches + out_launches)
    flop = weights.flops_per_pixel("base", 9) * W * H
    # the convs' share of a frame comes from the stamped pass (union / that pass's frame time); applied to the timed
    # region's frame time it gives the conv time there: never more than the step
    share = min(conv_union_ms / stamped_ms, 1.0)
    conv_ms = share * ms
    conv_tf = flop / (conv_ms * 1e-3) / 1e12
    conv_tf_long = flop / (ms_long * share * 1e-3) / 1e12
    traffic = None
    tp = os.path.join(ROOT, "profiles", "conv_traffic.json")
    if os.path.exists(tp):
      try:
        traffic = json.load(open(tp)).get("dram_bytes_per_launch_avg")
      except Exception:
        traffic = None
    roofline = {"bound": "tensor", "kernel": "conv3x3_tc_kernel (%d launches/frame)" % conv_launches,
                "achieved": round(conv_tf, 1), "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": round(conv_tf / peaks["bf16_tflops"], 4), "traffic": traffic,
                "peak_source": peaks_src + ", burst dense bf16: the timed region is %.0f ms at full clocks" % (ms * K),
                "alg_flop_per_launch_avg": flop / max(conv_launches, 1),
                "avg_launch_ms": round(conv_ms / max(conv_launches, 1), 5),
                "conv_ms_per_frame": round(conv_ms, 4), "share_of_step": round(share, 4),
                "stamped_pass": {"ms_per_step": round(stamped_ms, 4), "conv_union_ms": round(conv_union_ms, 4)},
                "how": "union of the conv grids' in-frame intervals (%globaltimer stamps: first CTA past griddepcontrol.wait .. "
                       "last CTA exit) over K back-to-back frames, divided by that pass's frame time = the convs' share of a "
                       "frame; conv_ms_per_frame = share x ms_per_step",
                "sustained": {"seconds": round(tl1 - tl0, 2), "frames": n_long, "ms_per_step": round(ms_long, 4),
                              "achieved": round(conv_tf_long, 1), "peak": peaks["bf16_tflops_sustained"],
                              "frac": round(conv_tf_long / peaks["bf16_tflops_sustained"], 4),
                              "how": "same frames back to back for >= 3 s; conv time = ms_per_step x share_of_step; "
                                     "peak = sustained dense bf16 of MEASURED_PEAKS.json", "clocks": clocks_long}}
    px = W * H
    passes = {
      "input_process": {"ms": round(in_ms, 4), "alg_bytes": px * (36 + 32), "GB/s": round(px * 68 / (in_ms * 1e-3) / 1e9, 1),
                        "frac_hbm": round(px * 68 / (in_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], 4)},
      "output_process": ({"ms": round(out_ms, 4), "alg_bytes": px * (32 + 12), "GB/s": round(px * 44 / (out_ms * 1e-3) / 1e9, 1),
                          "frac_hbm": round(px * 44 / (out_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], 4)} if out_launches else
                         {"ms": 0.0, "fused_into": "dec_conv0 epilogue (no tensor write / re-read, no launch); its time is inside "
                                                   "conv_layers_ms.dec_conv0, which then moves 64 B/px in + 12 B/px out"}),
      "conv_layers_ms": conv_layers,
    }
    f.release()
    del t, out

    # ---- end to end: host frames in, host frames out --------------------------------------------------------
    e2e = e2e_alt = None
    if not args.no_e2e:
      e2e = bench_e2e_staged(api, torch, local_rank, imgs, tza, W, H, K, Wm)
      e2e_alt = bench_e2e(api, torch, local_rank, imgs, tza, W, H, K, Wm)
    dev.release()
    in_flight = None if args.no_e2e else bench_frames_in_flight(api, torch, local_rank, imgs, tza, W, H, K, Wm)

    # ---- the other single-GPU lines of BASELINE.json's metric: the 8K frame (base of the 1->N scaling curve) and
    # config 3's model (large UNet: cleanAux + quality=high) on it ------------------------------------------------
    extra = {}
    if not args.no_8k and not args.explicit_size:
      extra = bench_8k(api, torch, local_rank, max(K // 2, 5), peaks)
  kernel_to_beat = None if args.no_kernel_to_beat else reference_cuda_device(sampler, W, H)
  sampler.stop()

  cpu = None if args.no_cpu_baseline else cpu_port_throughput(tza, W, H)
  line = {
    "metric": METRIC, "value": round(px / (ms * 1e-3) / 1e6, 1), "unit": "Mpix/s", "n_gpus": 1, "steps": K, "warmup": Wm,
    "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
    "dtype": "f16", "accumulate": "f32", "data": "synthetic",
    "config": dict(workload_config(args, 1), tiles="%dx%d of %dx%d" % (info["tileCountW"], info["tileCountH"], info["tileW"], info["tileH"])),
    "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "passes": passes,
    "e2e_caller_double_buffered": e2e_alt, "kernel_to_beat": kernel_to_beat, "three_frames_in_flight": in_flight,
  }
  line.update(extra)
  print(json.dumps(line))


def bench_frames_in_flight(api, torch, gpu, imgs, tza, W, H, K, Wm, nsets=3):
  """Device-resident throughput with several frames in flight (a renderer that owns several frame buffers): nsets
  device/stream/filter sets on the one GPU, frames alternate between them, so the autoexposure + input process and
  the first convs of frame f+1 run under the low-resolution layers of frame f, whose persistent grids leave SMs idle.
  Reported NEXT to the headline, which stays one frame at a time on one filter."""
  nb = W * H * 12
  sets = []
  for _ in range(nsets):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
      d = api.Device((gpu,), streams=[s.cuda_stream]).commit()
      bufs = {k: d.new_buffer(nb) for k in ("color", "albedo", "normal", "output")}
      for k, v in imgs.items():
        bufs[k].write(v)
      f = d.new_filter("RT")
      for k, b in bufs.items():
        f.set_image(k, b, api.capi.FORMAT_FLOAT3, W, H)
      f.set("hdr", True); f.set("quality", api.QUALITY_HIGH); f.set_data("weights", tza); f.commit()
    sets.append((s, d, bufs, f))

  def frame(i):
    with torch.cuda.stream(sets[i % nsets][0]):
      sets[i % nsets][3].execute_async()

  for i in range(max(Wm, nsets) + nsets):
    frame(i)
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record(sets[0][0])
  for s, *_ in sets[1:]:
    s.wait_event(e0)
  for i in range(K):
    frame(i)
  for s, *_ in sets[1:]:
    join = torch.cuda.Event(); join.record(s); sets[0][0].wait_event(join)
  e1.record(sets[0][0])
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1) / K
  for s, d, bufs, f in sets:
    f.release()
    for b in bufs.values():
      b.release()
    d.release()
  return {"sets": nsets, "ms_per_step": round(ms, 4), "value": round(W * H / ms / 1e3, 1), "unit": "Mpix/s",
          "how": "%d device/stream/filter sets on one GPU, frames alternate; CUDA events around K frames" % nsets}


def bench_e2e_staged(api, torch, gpu, imgs, tza, W, H, K, Wm, gpus=None):
  """End to end through the public API with the frame in HOST memory: the images given to
  oidnb200SetSharedFilterImage are pinned host buffers, and ONE filter is executed asynchronously frame after frame.
  The library stages the tiles itself (copy-in / compute / copy-out streams per engine, two slot sets), so the
  copy-in of frame f+1 and the copy-out of frame f-1 overlap the convolutions of frame f -- what `oidnBenchmark
  --buffer hostcopy` leaves to the application (apps/oidnBenchmark.cpp:165-180,343-359). The result alternates between
  two host output images (a consumer reads frame f-2 while f is in flight). gpus: one engine per listed GPU, every
  engine pulls its own tiles over its own PCIe link."""
  gpus = tuple(gpus) if gpus else (gpu,)
  nb = W * H * 12
  d = api.Device(gpus).commit()
  hb = {k: d.new_buffer(nb, api.STORAGE_HOST) for k in ("color", "albedo", "normal")}
  ho = [d.new_buffer(nb, api.STORAGE_HOST) for _ in range(2)]
  for k, b in hb.items():
    b.write(imgs[k])
  f = d.new_filter("RT")
  for k, b in hb.items():
    f.set_image(k, b, api.capi.FORMAT_FLOAT3, W, H)
  f.set_image("output", ho[0], api.capi.FORMAT_FLOAT3, W, H)
  f.set("hdr", True); f.set("quality", api.QUALITY_HIGH); f.set_data("weights", tza)
  f.commit()

  def frame(i):
    f.set_image("output", ho[i % 2], api.capi.FORMAT_FLOAT3, W, H)   # pointer-only change: no rebuild
    f.commit()
    f.execute_async()

  for i in range(max(Wm, 2)):
    frame(i)
  d.sync()
  assert f.info()["staged"] == 1
  info = f.info()
  t0 = time.perf_counter()
  for i in range(K):
    frame(i)
  d.sync()
  dt = (time.perf_counter() - t0) / K
  ntiles = info["tileCountH"] * info["tileCountW"]
  _, tiles = api.plan_tiles(H, W, False, 1, len(gpus), d.get("maxTilePixels"), d.get("tilePolicy"))
  h2d = sum(t["H1"] * t["W1"] for t in tiles) * 36
  f.release()
  for b in list(hb.values()) + ho:
    b.release()
  d.release()
  return {"value": round(W * H / dt / 1e6, 1), "unit": "Mpix/s", "ms_per_step": round(dt * 1e3, 4),
          "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": nb, "engines": len(gpus),
          "tiles": "%dx%d of %dx%d" % (info["tileCountW"], info["tileCountH"], info["tileW"], info["tileH"]) if ntiles else "",
          "how": "pinned host fp32 images set with oidnb200SetFilterImage (host-storage buffers), one filter, "
                 "oidnb200ExecuteFilterAsync per frame and one oidnb200SyncDevice at the end; the library stages tiles with "
                 "copy engines on its own copy-in / compute / copy-out streams (device parameter staging, auto); wall clock"}


def bench_8k(api, torch, gpu, K, peaks):
  """7680x4320 on ONE GPU: the base UNet (same filter as the headline; the N=1 point of the 8K scaling curve) and the
  large UNet (BASELINE config 3: cleanAux=true + quality=high)."""
  W, H = 7680, 4320
  imgs = synth.benchmark_images(W, H, hdr=True, seed=1)
  out = {}
  stream = torch.cuda.Stream()
  with torch.cuda.stream(stream):
    dev = api.Device((gpu,), streams=[stream.cuda_stream]).commit()
    t = {k: torch.from_numpy(v).cuda() for k, v in imgs.items()}
    o = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
    for key, kind, clean in (("frame_8k", "base", False), ("large_unet_8k", "large", True)):
      f = dev.new_filter("RT")
      for k, v in t.items():
        f.set_image(k, v)
      f.set_image("output", o)
      f.set("hdr", True); f.set("quality", api.QUALITY_HIGH); f.set("cleanAux", clean)
      f.set_data("weights", weights.model_tza(kind, 9, seed=0))
      f.commit()
      info = f.info()
      for _ in range(3):
        f.execute_async()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      torch.cuda.synchronize()
      e0.record(stream)
      for _ in range(K):
        f.execute_async()
      e1.record(stream)
      torch.cuda.synchronize()
      ms = e0.elapsed_time(e1) / K
      dev.set("profile", 2)
      f.execute(); f.profile()
      for _ in range(3):
        f.execute_async()
      torch.cuda.synchronize()
      prof = f.profile()
      dev.set("profile", 0)
      union = min(sum(m for _, kd, _, m in prof if kd == 3) / 3, ms)   # the stamped frames run a little later (warmer board)
      tf = weights.flops_per_pixel(kind, 9) * W * H / (union * 1e-3) / 1e12
      out[key] = {"workload": "RT hdr+alb+nrm 7680x4320, %s UNet%s, one GPU" % (kind, " (cleanAux, quality=high)" if clean else ""),
                  "ms_per_step": round(ms, 4), "value": round(W * H / ms / 1e3, 1), "unit": "Mpix/s", "steps": K,
                  "tiles": "%dx%d of %dx%d" % (info["tileCountW"], info["tileCountH"], info["tileW"], info["tileH"]),
                  "conv_tflops": round(tf, 1), "conv_frac_of_burst_peak": round(tf / peaks["bf16_tflops"], 4)}
      f.release()
    dev.release()
  return out


def reference_cuda_device(sampler, W, H):
  """The kernel to beat: the reference's own CUDA device (devices/cuda: CUTLASS mma.sync kernels
  compiled for sm_100) on the same frame size through its own oidnBenchmark, in this run, under the same clock
  sampler. baseline/_ref holds the unmodified reference build (tools/build_reference_cuda.sh); absent -> None."""
  exe = os.path.join(ROOT, "baseline", "_ref", "bin", "oidnBenchmark")
  if not os.path.exists(exe):
    return None
  env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "baseline", "_ref", "lib"))
  t0 = time.time()
  try:
    r = subprocess.run([exe, "-d", "cuda", "-r", r"RT\.hdr_alb_nrm\.%dx%d" % (W, H), "-q", "high"], capture_output=True, text=True,
                       timeout=120, env=env)
  except Exception as e:  # noqa: BLE001
    return {"unavailable": str(e)[:200]}
  t1 = time.time()
  import re
  m = re.search(r"([0-9.]+) msec/image", r.stdout)
  if not m:
    return {"unavailable": (r.stdout + r.stderr)[-200:]}
  # the benchmark spends its first seconds building the filter: the clock samples under load are the tail
  return {"ms": float(m.group(1)), "impl": "reference devices/cuda (unmodified; CUTLASS Sm80 mma.sync kernels, fp16 accumulate off "
          "with quality=high), oidnBenchmark -d cuda -r RT.hdr_alb_nrm.%dx%d -q high, device-resident" % (W, H),
          "clocks": sampler.window(max(t0, t1 - 2.0), t1)}


def bench_e2e(api, torch, gpu, imgs, tza, W, H, K, Wm):
  """Frames arrive in pinned host memory and leave to pinned host memory: oidnWriteBufferAsync x3,
  oidnExecuteFilterAsync, oidnReadBufferAsync per frame. Two device/filter sets (own streams)
  alternate so the copies of one frame overlap the convolutions of the other."""
  nb = W * H * 12
  host_in = {k: torch.from_numpy(v).pin_memory() for k, v in imgs.items()}
  host_out = [torch.zeros((H, W, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
  sets = []
  for i in range(2):
    d = api.Device((gpu,)).commit()
    bufs = {k: d.new_buffer(nb) for k in ("color", "albedo", "normal", "output")}
    f = d.new_filter("RT")
    for k, b in bufs.items():
      f.set_image(k, b, api.capi.FORMAT_FLOAT3, W, H)
    f.set("hdr", True); f.set("quality", api.QUALITY_HIGH); f.set_data("weights", tza)
    f.commit()
    sets.append((d, bufs, f))

  L = api.capi.lib()

  def frame(i):
    d, bufs, f = sets[i % 2]
    for k in ("color", "albedo", "normal"):
      L.oidnb200WriteBufferAsync(bufs[k]._h, 0, nb, host_in[k].data_ptr())
    f.execute_async()
    L.oidnb200ReadBufferAsync(bufs["output"]._h, 0, nb, host_out[i % 2].data_ptr())

  for i in range(max(Wm, 2)):
    frame(i)
  for d, _, _ in sets:
    d.sync()
  t0 = time.perf_counter()
  for i in range(K):
    if i >= 2:
      sets[i % 2][0].sync()      # the host consumes frame i-2's result before its buffers are reused
    frame(i)
  for d, _, _ in sets:
    d.sync()
  dt = (time.perf_counter() - t0) / K
  for d, bufs, f in sets:
    f.release()
    for b in bufs.values():
      b.release()
    d.release()
  return {"value": round(W * H / dt / 1e6, 1), "unit": "Mpix/s", "ms_per_step": round(dt * 1e3, 4),
          "h2d_bytes_per_step": 3 * nb, "d2h_bytes_per_step": nb,
          "how": "pinned host fp32 images, oidnb200WriteBufferAsync x3 + ExecuteFilterAsync + ReadBufferAsync per frame, "
                 "two device/stream sets alternating; wall clock around K frames with device sync at both ends"}


if __name__ == "__main__":
  main()
