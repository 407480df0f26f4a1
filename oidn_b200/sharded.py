"""One frame, several GPUs, one process per GPU (torchrun): tile sharding of the UNet path.

The reference deals tiles round-robin to the engines of ONE device object (core/unet_filter.cpp:219)
and joins them with device->submitBarrier() (:178, :243); its only multi-engine backend (SYCL) shares
USM pointers between the engines. The same scheme with one process per GPU:

  * the frame (color/albedo/normal/output) lives in rank 0's HBM; rank 0 exports the four buffers
    as CUDA IPC handles and every other rank maps them (NVLink peer mappings);
  * every rank builds the same tile plan (tile count % world == 0) and executes the tiles whose
    index % world == rank: its input-process kernel loads the tile straight from rank 0's buffers
    (P2P loads), its output-process kernel stores the interior rectangle straight into rank 0's
    output (P2P stores). Tiles write disjoint rectangles, so no reduction is needed;
  * the one global value, the autoexposure scale, is computed by rank 0 and broadcast (4 bytes,
    NCCL, stream ordered); a 4-byte all-reduce at the end of the frame is the join (the
    submitBarrier of the single-process device).
"""
import ctypes as C

import numpy as np

from . import api, capi


def tiles_of_rank(H, W, large, world, rank, max_tile_pixels=7680 * 4352, policy=1):
  """(plan, [tile rects of this rank]) -- host logic only, usable without a GPU. Defaults = the
  device defaults (maxTilePixels, tilePolicy), so this is the plan every rank's filter builds."""
  plan, tiles = api.plan_tiles(H, W, large, 1, world, max_tile_pixels, policy)
  return plan, [t for i, t in enumerate(tiles) if i % world == rank]


def broadcast_object(dist, obj, src=0):
  box = [obj]
  dist.broadcast_object_list(box, src=src)
  return box[0]


class ShardedFilter:
  """RT filter over a frame resident on rank 0, executed by all ranks of the process group.

  stage=True (default): a peer rank moves its tiles with the copy engines -- input rectangles
  (with overlap) rank 0 -> local staging images before the tile runs, output rectangles local ->
  rank 0 after it -- so NVLink traffic occupies no SM and overlaps the convolutions of another
  frame in flight. stage=False: the input/output-process kernels dereference rank 0's memory
  directly (P2P loads/stores)."""

  def __init__(self, dist, torch, device, W, H, tza, hdr=True, quality=api.QUALITY_HIGH, clean_aux=False,
               aux=True, frame=None, stage=True):
    self.dist, self.torch, self.dev = dist, torch, device
    self.rank, self.world = dist.get_rank(), dist.get_world_size()
    self.W, self.H, self.hdr = W, H, hdr
    nb = W * H * 12
    names = ("color", "albedo", "normal", "output") if aux else ("color", "output")
    self.inputs = names[:-1]
    self.bufs = {}
    if self.rank == 0:
      for n in names:
        self.bufs[n] = device.new_buffer(nb)
        if frame is not None and n in frame:
          self.bufs[n].write(frame[n])
      handles = {n: self.bufs[n].ipc_handle() for n in names}
    else:
      handles = None
    handles = broadcast_object(dist, handles, 0)
    if self.rank != 0:
      for n in names:
        self.bufs[n] = device.import_buffer(handles[n], nb)
    self.staged = bool(stage) and self.rank != 0
    self.local = {n: device.new_buffer(nb) for n in names} if self.staged else self.bufs
    self.scale = torch.ones(1, dtype=torch.float32, device="cuda")
    self.token = torch.zeros(1, dtype=torch.float32, device="cuda")
    f = device.new_filter("RT")
    for n in names:
      f.set_image(n, self.local[n], capi.FORMAT_FLOAT3, W, H)
    f.set("hdr", bool(hdr)); f.set("quality", quality); f.set("cleanAux", bool(clean_aux))
    f.set("numShards", self.world); f.set("shardIndex", self.rank)
    if hdr:
      f.set_input_scale_ptr(self.scale.data_ptr())
    f.set_data("weights", tza)
    f.commit()
    self.filter = f
    info = f.info()
    plan, self.tiles = tiles_of_rank(H, W, bool(info["largeModel"]), self.world, self.rank,
                                     device.get("maxTilePixels"), device.get("tilePolicy"))
    assert (plan["tileCountH"], plan["tileCountW"], plan["tileH"], plan["tileW"]) == \
           (info["tileCountH"], info["tileCountW"], info["tileH"], info["tileW"]), (plan, info)
    if self.rank == 0 and hdr:
      L = capi.lib()
      self.ae_scratch = torch.zeros(L.oidnb200_autoexposure_scratch_bytes(H, W), dtype=torch.uint8, device="cuda")
      self.ae_img = capi.Image(self.bufs["color"].data, capi.FORMAT_FLOAT3, W, H, 12, 12 * W)

  def _copy_rects(self, names, to_local):
    pitch = self.W * 12
    for t in self.tiles:
      h, w, nh, nw = (t["hSrc"], t["wSrc"], t["H1"], t["W1"]) if to_local else (t["hDst"], t["wDst"], t["H2"], t["W2"])
      off = h * pitch + w * 12
      for n in names:
        loc, rem = self.local[n].data + off, self.bufs[n].data + off
        if to_local:
          self.dev.copy_rect_async(loc, pitch, rem, pitch, nw * 12, nh)
        else:
          self.dev.copy_rect_async(rem, pitch, loc, pitch, nw * 12, nh)

  def execute_async(self):
    """Enqueues one frame on the device's stream of every rank."""
    torch, dist = self.torch, self.dist
    if self.hdr:
      if self.rank == 0:
        rc = capi.lib().oidnb200_autoexposure_launch(C.byref(self.ae_img), self.ae_scratch.data_ptr(),
                                                     self.scale.data_ptr(), torch.cuda.current_stream().cuda_stream)
        if rc != 0:
          raise RuntimeError(capi.lib().oidnb200_last_error().decode())
      dist.broadcast(self.scale, src=0)       # also orders the peers after rank 0's frame upload
    else:
      dist.all_reduce(self.token)             # frame-start ordering without a scale
    if self.staged:
      self._copy_rects(self.inputs, True)     # NVLink DMA: rank 0 -> local tile inputs (with overlap)
    self.filter.execute_async()
    if self.staged:
      self._copy_rects(("output",), False)    # NVLink DMA: local interior rectangles -> rank 0's output
    dist.all_reduce(self.token)               # join: every rank's rectangles are in rank 0's output

  def release(self):
    self.filter.release()
    if self.staged:
      for b in self.local.values():
        b.release()
    self.dist.barrier()
    if self.rank != 0:
      for b in self.bufs.values():
        b.release()
    self.dist.barrier()
    if self.rank == 0:
      for b in self.bufs.values():
        b.release()


def bench_main(args, rank, world, local_rank):
  """bench.py's N>1 arm (launched by torchrun, one rank per GPU)."""
  import json
  import os
  import time

  import torch
  import torch.distributed as dist

  import bench as B
  from . import synth, weights

  torch.cuda.set_device(local_rank)
  dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
  W, H, K, Wm = args.width, args.height, args.steps, args.warmup
  tza = weights.model_tza("base", 9, seed=0)
  peaks, peaks_src = B.load_peaks()
  sampler = B.ClockSampler(local_rank) if rank == 0 else None
  frame = synth.benchmark_images(W, H, hdr=True, seed=1) if rank == 0 else None

  # Two frames in flight (a renderer double-buffers its frame): two device/stream/filter sets, frames
  # alternate between them, so the peers' NVLink reads of frame f+1 overlap the convolutions of frame f.
  # Collectives are issued in frame order by every rank.
  sets = []
  for i in range(2):
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
      dev = api.Device((local_rank,), streams=[stream.cuda_stream]).commit()
      sf = ShardedFilter(dist, torch, dev, W, H, tza, hdr=True, frame=frame)
    sets.append((stream, dev, sf))
  info = sets[0][2].filter.info()
  ntiles = info["tileCountH"] * info["tileCountW"]

  def run_frame(i):
    stream, _, sf = sets[i % 2]
    with torch.cuda.stream(stream):
      sf.execute_async()

  sA, sB = sets[0][0], sets[1][0]
  for i in range(max(Wm, 2)):
    run_frame(i)
  torch.cuda.synchronize(); dist.barrier()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  join = torch.cuda.Event()
  t0 = time.time()
  e0.record(sA)
  sB.wait_event(e0)                 # neither stream starts a timed frame before e0
  for i in range(K):
    run_frame(i)
  join.record(sB)
  sA.wait_event(join)
  e1.record(sA)                     # after the last frame of both streams
  torch.cuda.synchronize(); dist.barrier()
  t1 = time.time()
  ms = torch.tensor([e0.elapsed_time(e1) / K], device="cuda")
  dist.all_reduce(ms, op=dist.ReduceOp.MAX)
  ms = float(ms.item())

  # per-op times on this rank's tiles (roofline of the dominant kernel, rank 0's share): one set alone
  stream, dev, sf = sets[0]
  with torch.cuda.stream(stream):
    dev.set("profile", 1)
    sf.execute_async(); torch.cuda.synchronize(); sf.filter.profile()
    for _ in range(K):
      sf.execute_async()
    torch.cuda.synchronize()
    prof = sf.filter.profile()
    dev.set("profile", 0)
  dist.barrier()

  # end to end: frames arrive in rank 0's pinned host memory and the results return there; the two
  # sets alternate so the PCIe copies of one frame overlap the other frame's execution
  e2e = None
  if not args.no_e2e:
    nb = W * H * 12
    if rank == 0:
      hin = {k: torch.from_numpy(v).pin_memory() for k, v in frame.items()}
      hout = [torch.zeros((H, W, 3), dtype=torch.float32).pin_memory() for _ in range(2)]
    L = capi.lib()

    def e2e_frame(i):
      stream, _, sf = sets[i % 2]
      with torch.cuda.stream(stream):
        if rank == 0:
          for k in ("color", "albedo", "normal"):
            L.oidnb200WriteBufferAsync(sf.bufs[k]._h, 0, nb, hin[k].data_ptr())
        sf.execute_async()
        if rank == 0:
          L.oidnb200ReadBufferAsync(sf.bufs["output"]._h, 0, nb, hout[i % 2].data_ptr())
    for i in range(2):
      e2e_frame(i)
    torch.cuda.synchronize(); dist.barrier()
    w0 = time.perf_counter()
    for i in range(K):
      if i >= 2:
        sets[i % 2][0].synchronize()   # the host consumes frame i-2's result before its buffers are reused
      e2e_frame(i)
    torch.cuda.synchronize(); dist.barrier()
    dt = torch.tensor([(time.perf_counter() - w0) / K], device="cuda")
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dt = float(dt.item())
    e2e = {"value": round(W * H / dt / 1e6, 1), "unit": "Mpix/s", "ms_per_step": round(dt * 1e3, 4),
           "h2d_bytes_per_step": 3 * nb, "d2h_bytes_per_step": nb,
           "how": "rank 0: pinned host fp32 frame -> WriteBufferAsync x3, all ranks: sharded execute, rank 0: ReadBufferAsync; "
                  "two frame sets alternating (copies of one frame overlap the other's execution); wall clock, max over ranks"}

  if rank == 0:
    clocks = sampler.window(t0, t1); sampler.stop()
    conv_ms = sum(m for _, kind, _, m in prof if kind == 0) / K
    conv_launches = sum(n for _, kind, n, _ in prof if kind == 0) // K
    # rank 0's tiles: algorithmic FLOPs of the pixels it outputs
    _, mine = tiles_of_rank(H, W, False, world, 0)
    my_px = sum(t["H2"] * t["W2"] for t in mine)
    conv_tf = weights.flops_per_pixel("base", 9) * my_px / (conv_ms * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": "conv3x3_tc_kernel on rank 0 (%d launches/frame)" % conv_launches,
                "achieved": round(conv_tf, 1), "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": round(conv_tf / peaks["bf16_tflops_sustained"], 4), "traffic": None,
                "peak_source": peaks_src + ", sustained dense bf16", "conv_ms_per_frame": round(conv_ms, 4),
                "alg_flop_per_launch_avg": weights.flops_per_pixel("base", 9) * my_px / max(conv_launches, 1),
                "avg_launch_ms": round(conv_ms / max(conv_launches, 1), 5)}
    line = {
      "metric": B.METRIC, "value": round(W * H / (ms * 1e-3) / 1e6, 1), "unit": "Mpix/s", "n_gpus": world, "steps": K, "warmup": Wm,
      "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
      "dtype": "f16", "accumulate": "f32", "data": "synthetic",
      "config": dict(B.workload_config(args, world), tiles="%dx%d of %dx%d, %d per rank" % (
        info["tileCountW"], info["tileCountH"], info["tileW"], info["tileH"], ntiles // world)),
      "roofline": roofline, "cpu_baseline": None, "e2e": e2e,
      "gpu_launches": K * (1 + ntiles * info["numOps"]), "clocks": clocks,   # all ranks: autoexposure + every tile's ops
      "exchange": "CUDA IPC peer mappings of rank 0's frame; peers stage their tile rectangles with copy-engine transfers over "
                  "NVLink (in: tile + overlap, out: interior); NCCL broadcast(4 B) + all_reduce(4 B) per frame; two frames in flight",
    }
    print(json.dumps(line))
  for _, dev, sf in sets:
    sf.release()
    dev.release()
  dist.barrier()
  dist.destroy_process_group()
