"""One frame, several GPUs, one process per GPU (torchrun): tile sharding of the UNet path.

The reference deals tiles round-robin to the engines of ONE device object (core/unet_filter.cpp:219)
and joins them with device->submitBarrier() (:178, :243); its only multi-engine backend (SYCL) shares
USM pointers between the engines. The same scheme with one process per GPU: every rank builds the
same tile plan (tile count % world == 0) and executes the tiles whose index % world == rank. Tiles
write disjoint rectangles of the output, so there is no reduction on the data path. Two ways of
holding the frame:

  source="distributed" (bench.py): every GPU holds the pixels of its own tiles (tile + overlap) --
    what a multi-GPU renderer produces, and what each rank pulls from host memory over its own PCIe
    link in the end-to-end path. The one global value, the autoexposure scale, is exchanged exactly:
    each rank computes log2(mean luminance) of the <=16x16 bins of its tiles
    (oidnb200_autoexposure_bins_launch) into its bin array, a copy kernel stores those bin rectangles
    into every peer's array (CUDA IPC mapping; ~0.5 MB per array at 8K), and once every rank's
    rectangles have arrived every rank folds the array in the same fixed order
    (oidnb200_autoexposure_reduce_launch) -- bit-identical to one GPU. Output rectangles are
    assembled in rank 0's output buffer with copy-engine peer writes over NVLink, or go straight to
    the host frame in the end-to-end path. The two ordering points of a frame ("all bins are here",
    "the frame is complete everywhere") are peer flags: oidnb200_flag_signal_launch stores the
    frame's sequence number into this rank's slot on every rank, oidnb200_flag_wait_launch polls
    the local slots -- no collective on the data path (exchange="peer", default). exchange="nccl"
    does the same with an all-reduce of a zero-filled bin array (x + 0 = x) and a 4-byte all-reduce.

  source="rank0": the whole frame (color/albedo/normal/output) lives in rank 0's HBM (the
    single-pointer contract of oidnSetSharedFilterImage); rank 0 exports the four buffers as CUDA IPC
    handles, runs the autoexposure and broadcasts the scale (4 bytes); the peers either stage their
    tile rectangles with copy-engine transfers over NVLink (stage=True) or dereference rank 0's
    memory directly from the input/output-process kernels (stage=False). Rank 0's NVLink egress
    bounds this mode beyond 2 GPUs (profiles/README.md).
"""
import ctypes as C
import mmap
import os
import uuid

import numpy as np

from . import api, capi


def tiles_of_rank(H, W, large, world, rank, max_tile_pixels=7680 * 4352, policy=1):
  """(plan, [tile rects of this rank]) -- host logic only, usable without a GPU. Defaults = the
  device defaults (maxTilePixels, tilePolicy), so this is the plan every rank's filter builds."""
  plan, tiles = api.plan_tiles(H, W, large, 1, world, max_tile_pixels, policy)
  return plan, [t for i, t in enumerate(tiles) if i % world == rank]


def bin_grid(H, W):
  """Autoexposure bin grid of an HxW image (core/autoexposure.h:20-24)."""
  return (H + 15) // 16, (W + 15) // 16


def _first_bin_at_or_after(x, n, size):
  """Smallest bin index i (0..n) whose first pixel i*size//n is >= x."""
  i = min(n, (x * n + size - 1) // size)
  while i > 0 and (i - 1) * size // n >= x:
    i -= 1
  while i < n and i * size // n < x:
    i += 1
  return i


def bins_of_tile(t, H, W):
  """(bh0, bh1, bw0, bw1): the autoexposure bins owned by a tile = the bins whose first pixel lies in
  the tile's destination rectangle. Destination rectangles partition the image, so these rectangles
  partition the bin grid; a bin is at most 16 px and the tile's source rectangle extends >= 96 px
  past every interior edge, so all pixels of an owned bin are inside the tile's source rectangle."""
  nbh, nbw = bin_grid(H, W)
  bh0 = _first_bin_at_or_after(t["hDst"], nbh, H); bh1 = _first_bin_at_or_after(t["hDst"] + t["H2"], nbh, H)
  bw0 = _first_bin_at_or_after(t["wDst"], nbw, W); bw1 = _first_bin_at_or_after(t["wDst"] + t["W2"], nbw, W)
  assert bh0 * H // nbh >= t["hSrc"] and bh1 * H // nbh <= t["hSrc"] + t["H1"], (t, bh0, bh1)
  assert bw0 * W // nbw >= t["wSrc"] and bw1 * W // nbw <= t["wSrc"] + t["W1"], (t, bw0, bw1)
  return bh0, bh1, bw0, bw1


def broadcast_object(dist, obj, src=0):
  box = [obj]
  dist.broadcast_object_list(box, src=src)
  return box[0]


class SharedHostFrame:
  """Images of one frame in host memory shared by all ranks of the node: a POSIX shared-memory file
  mapped by every process and page-locked in each (cudaHostRegister), so every GPU moves its own
  tiles over its own PCIe link. Falls back to private pinned memory per rank (`shared` False) when
  /dev/shm cannot hold the frame."""

  def __init__(self, dist, torch, names, H, W):
    self.names, self.H, self.W = tuple(names), H, W
    self.nb = H * W * 12
    rank = dist.get_rank()
    total = self.nb * len(self.names)
    path = None
    if rank == 0:
      path = "/dev/shm/oidnb200_%s" % uuid.uuid4().hex
      try:
        fd = os.open(path, os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
        os.posix_fallocate(fd, 0, total)
        os.close(fd)
      except OSError:
        try:
          os.unlink(path)
        except OSError:
          pass
        path = None
    path = broadcast_object(dist, path, 0)
    self.shared = path is not None
    self._registered = None
    self._torch = torch
    if self.shared:
      fd = os.open(path, os.O_RDWR)
      self._map = mmap.mmap(fd, total)
      os.close(fd)
      dist.barrier()
      if rank == 0:
        os.unlink(path)   # the mappings keep it alive; nothing is left behind if a rank dies
      base = np.frombuffer(self._map, dtype=np.uint8)
      ptr = base.ctypes.data
      rc = torch.cuda.cudart().cudaHostRegister(ptr, total, 0)
      if int(rc) != 0:
        raise RuntimeError("cudaHostRegister of the shared host frame failed: %s" % rc)
      self._registered = ptr
      self.images = {n: base[i * self.nb:(i + 1) * self.nb].view(np.float32).reshape(H, W, 3) for i, n in enumerate(self.names)}
    else:
      self._pinned = {n: torch.zeros((H, W, 3), dtype=torch.float32).pin_memory() for n in self.names}
      self.images = {n: t.numpy() for n, t in self._pinned.items()}

  def ptr(self, name):
    return self.images[name].ctypes.data

  def release(self):
    if self._registered is not None:
      self._torch.cuda.cudart().cudaHostUnregister(self._registered)
      self._registered = None
    self.images = {}


def bin_rect_copies(bin_rects, nbw):
  """The 2D copies (byte offset, pitch, width in bytes, rows) that carry a rank's bin rectangles from its own bin
  array (fp32, nbw per row) to the same place in a peer's."""
  pitch = nbw * 4
  return [(bh0 * pitch + bw0 * 4, pitch, (bw1 - bw0) * 4, bh1 - bh0) for (bh0, bh1, bw0, bw1) in bin_rects
          if bh1 > bh0 and bw1 > bw0]


class ShardedFilter:
  """RT filter over one frame, executed by all ranks of the process group (see the module docstring).

  frame: dict of full-frame float32 HxWx3 arrays. source="rank0": given on rank 0 only and uploaded
  there. source="distributed": given on every rank (or None to leave the local tile inputs to
  upload_tiles()); each rank uploads only the source rectangles of its own tiles."""

  def __init__(self, dist, torch, device, W, H, tza, hdr=True, quality=api.QUALITY_HIGH, clean_aux=False,
               aux=True, frame=None, stage=True, source="rank0", own_groups=True, exchange="peer"):
    assert source in ("rank0", "distributed") and exchange in ("peer", "nccl")
    # exchange (distributed frames): "peer" = no collective on the data path -- bin rectangles go to every rank's bin
    # array by copy-engine peer writes, frame steps are joined with peer flags (oidnb200_flag_signal / _wait: one tiny
    # block that runs next to a persistent conv CTA); "nccl" = all-reduce of the bin array + 4-byte all-reduce join
    # (round 1; a collective's kernel holds SMs while it waits for the slowest rank).
    exchange = os.environ.get("OIDN_B200_EXCHANGE", exchange)    # A/B on one box
    self.exchange = exchange if source == "distributed" else "nccl"
    self.dist, self.torch, self.dev = dist, torch, device
    self._check_stream()
    self.rank, self.world = dist.get_rank(), dist.get_world_size()
    self.W, self.H, self.hdr, self.source = W, H, hdr, source
    nb = W * H * 12
    names = ("color", "albedo", "normal", "output") if aux else ("color", "output")
    self.inputs = names[:-1]
    self.bufs = {}
    shared_names = names if source == "rank0" else ("output",)   # buffers rank 0 exports to the peers
    if self.rank == 0:
      for n in shared_names:
        self.bufs[n] = device.new_buffer(nb)
        if source == "rank0" and frame is not None and n in frame:
          self.bufs[n].write(frame[n])
      handles = {n: self.bufs[n].ipc_handle() for n in shared_names}
    else:
      handles = None
    handles = broadcast_object(dist, handles, 0)
    if self.rank != 0:
      for n in shared_names:
        self.bufs[n] = device.import_buffer(handles[n], nb)
    if source == "rank0":
      self.staged = bool(stage) and self.rank != 0
      self.local = {n: device.new_buffer(nb) for n in names} if self.staged else self.bufs
    else:
      self.staged = False
      self.local = {n: device.new_buffer(nb) for n in self.inputs}
      self.local["output"] = self.bufs["output"] if self.rank == 0 else device.new_buffer(nb)
    self.scale = torch.ones(1, dtype=torch.float32, device="cuda")
    self.token = torch.zeros(1, dtype=torch.float32, device="cuda")
    # Own communicators per filter: collectives on ONE communicator execute in issue order whatever
    # stream they are on, so with the default group frame f+1's exchange (start of its stream) would
    # wait for frame f's join (end of the other stream) and two frames in flight would not overlap.
    # new_group is collective: every rank creates its filters in the same order.
    self.pg_exchange = dist.new_group() if own_groups else None   # start-of-frame exchange (bins / scale)
    self.pg_join = dist.new_group() if own_groups else None       # end-of-frame join
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
    L = capi.lib()
    if hdr and source == "rank0" and self.rank == 0:
      self.ae_scratch = torch.zeros(L.oidnb200_autoexposure_scratch_bytes(H, W), dtype=torch.uint8, device="cuda")
      self.ae_img = capi.Image(self.bufs["color"].data, capi.FORMAT_FLOAT3, W, H, 12, 12 * W)
    nbh, nbw = bin_grid(H, W)
    self.nbins, self.nbw = nbh * nbw, nbw
    if hdr and source == "distributed":
      self.bins_local = torch.zeros(self.nbins, dtype=torch.float32, device="cuda")   # zeros outside the own bins
      self.bins_all = torch.zeros(self.nbins, dtype=torch.float32, device="cuda")
      self.bin_rects = [bins_of_tile(t, H, W) for t in self.tiles]
      self.ae_img = capi.Image(self.local["color"].data, capi.FORMAT_FLOAT3, W, H, 12, 12 * W)
    if self.exchange == "peer":
      # one exchange buffer per rank, mapped by every rank (CUDA IPC): [bin array][bins-ready flags][frame-done flags]
      self.off_a = (self.nbins * 4 + 255) // 256 * 256
      self.off_d = self.off_a + 256
      xsize = max(self.off_d + 256, 2 << 20)   # an allocation of its own (small cudaMalloc blocks share a 2 MiB page)
      self.xb = device.new_buffer(xsize)
      self.xb.write(np.zeros(xsize, np.uint8))
      handles = [None] * self.world
      dist.all_gather_object(handles, self.xb.ipc_handle())
      self.xpeer = [self.xb if r == self.rank else device.import_buffer(handles[r], xsize) for r in range(self.world)]
      self.slots_a = (C.c_void_p * self.world)(*[b.data + self.off_a + 4 * self.rank for b in self.xpeer])
      self.slots_d = (C.c_void_p * self.world)(*[b.data + self.off_d + 4 * self.rank for b in self.xpeer])
      self.seq = 0
      self.bin_copies = bin_rect_copies(self.bin_rects, self.nbw) if hdr else []
      self.bin_scatter = []     # (source rectangle in the own array, the same rectangle in a peer's) as 1-channel fp32 images
      for r, peer in enumerate(self.xpeer):
        if r != self.rank:
          for off, pitch, wbytes, rows in self.bin_copies:
            self.bin_scatter.append((capi.Image(self.xb.data + off, capi.FORMAT_FLOAT, wbytes // 4, rows, 4, pitch),
                                     capi.Image(peer.data + off, capi.FORMAT_FLOAT, wbytes // 4, rows, 4, pitch)))
      dist.barrier()
    if source == "distributed" and frame is not None:
      self.upload_tiles({n: frame[n].ctypes.data for n in self.inputs})
      self.dev.sync()

  def _check_stream(self):
    """The autoexposure kernels and the NCCL collectives are issued on torch's current stream, the filter and the
    rectangle copies on the device's engine stream: the frame is only ordered when they are the SAME stream, i.e.
    the device was created as api.Device((gpu,), streams=[s.cuda_stream]) and is used inside torch.cuda.stream(s)."""
    cur = self.torch.cuda.current_stream().cuda_stream
    have = self.dev.streams[0] if getattr(self.dev, "streams", None) else None
    if have is None or int(have) != int(cur):
      raise ValueError("ShardedFilter: the device must be created on torch's current stream "
                       "(api.Device((gpu,), streams=[stream.cuda_stream]) inside torch.cuda.stream(stream)); "
                       "device stream %s, current stream %s" % (have, cur))

  # ---- rectangle transfers (copy engines, stream ordered) ------------------------------------------
  def _rect(self, t, src):
    return (t["hSrc"], t["wSrc"], t["H1"], t["W1"]) if src else (t["hDst"], t["wDst"], t["H2"], t["W2"])

  def _copy_rects(self, names, to_local):
    pitch = self.W * 12
    for t in self.tiles:
      h, w, nh, nw = self._rect(t, to_local)
      off = h * pitch + w * 12
      for n in names:
        loc, rem = self.local[n].data + off, self.bufs[n].data + off
        if to_local:
          self.dev.copy_rect_async(loc, pitch, rem, pitch, nw * 12, nh)
        else:
          self.dev.copy_rect_async(rem, pitch, loc, pitch, nw * 12, nh)

  def upload_tiles(self, host_ptrs):
    """distributed: host frame (full-frame float3 images, pinned or not) -> the source rectangles of
    this rank's tiles in its local images. host_ptrs: {image name: address of pixel (0,0)}."""
    pitch = self.W * 12
    for t in self.tiles:
      h, w, nh, nw = self._rect(t, True)
      off = h * pitch + w * 12
      for n in self.inputs:
        self.dev.copy_rect_async(self.local[n].data + off, pitch, host_ptrs[n] + off, pitch, nw * 12, nh)

  def download_tiles(self, host_output_ptr):
    """distributed: this rank's output rectangles -> the host output frame."""
    pitch = self.W * 12
    for t in self.tiles:
      h, w, nh, nw = self._rect(t, False)
      off = h * pitch + w * 12
      self.dev.copy_rect_async(host_output_ptr + off, pitch, self.local["output"].data + off, pitch, nw * 12, nh)

  # ---- one frame ------------------------------------------------------------------------------------
  def execute_async(self, assemble=True):
    """Enqueues one frame on the device's stream of every rank. assemble=False (distributed only)
    leaves every rank's output rectangles in its local output image (the caller downloads them)."""
    torch, dist, L = self.torch, self.dist, capi.lib()
    self._check_stream()
    st = torch.cuda.current_stream().cuda_stream
    if self.source == "distributed" and self.exchange == "peer":
      def ck(rc):
        if rc != 0:
          raise RuntimeError(L.oidnb200_last_error().decode())
      self.seq += 1
      if self.hdr:
        own = self.xb.data
        for (bh0, bh1, bw0, bw1) in self.bin_rects:
          ck(L.oidnb200_autoexposure_bins_launch(C.byref(self.ae_img), bh0, bh1, bw0, bw1, own, st))
        # the bin rectangles of the own tiles -> every peer's bin array. By a copy KERNEL (peer stores), not the copy
        # engines: those are busy with the frame's bulk transfers (output rectangles, and in the end-to-end path the PCIe
        # uploads of the other frame in flight), and every rank's frame waits for these few kilobytes
        for src_img, dst_img in self.bin_scatter:
          ck(L.oidnb200_image_copy_launch(C.byref(src_img), C.byref(dst_img), st))
        ck(L.oidnb200_flag_signal_launch(self.slots_a, self.world, self.seq, st))
        ck(L.oidnb200_flag_wait_launch(own + self.off_a, self.world, self.seq, 10.0, st))   # every rank's bins are here
        ck(L.oidnb200_autoexposure_reduce_launch(own, self.nbins, self.scale.data_ptr(), st))
      self.filter.execute_async()
      if assemble and self.rank != 0:
        self._copy_rects(("output",), False)    # NVLink DMA: local interior rectangles -> rank 0's output
      ck(L.oidnb200_flag_signal_launch(self.slots_d, self.world, self.seq, st))
      ck(L.oidnb200_flag_wait_launch(self.xb.data + self.off_d, self.world, self.seq, 10.0, st))   # join: the frame is complete everywhere
      return
    if self.source == "distributed":
      if self.hdr:
        for (bh0, bh1, bw0, bw1) in self.bin_rects:
          if L.oidnb200_autoexposure_bins_launch(C.byref(self.ae_img), bh0, bh1, bw0, bw1, self.bins_local.data_ptr(), st) != 0:
            raise RuntimeError(L.oidnb200_last_error().decode())
        self.bins_all.copy_(self.bins_local)
        dist.all_reduce(self.bins_all, group=self.pg_exchange)   # the exchange step: every rank gets the complete bin array
        if L.oidnb200_autoexposure_reduce_launch(self.bins_all.data_ptr(), self.nbins, self.scale.data_ptr(), st) != 0:
          raise RuntimeError(L.oidnb200_last_error().decode())
      self.filter.execute_async()
      if assemble and self.rank != 0:
        self._copy_rects(("output",), False)    # NVLink DMA: local interior rectangles -> rank 0's output
      dist.all_reduce(self.token, group=self.pg_join)   # join: every rank's rectangles are where they belong
      return
    if self.hdr:
      if self.rank == 0:
        rc = L.oidnb200_autoexposure_launch(C.byref(self.ae_img), self.ae_scratch.data_ptr(), self.scale.data_ptr(), st)
        if rc != 0:
          raise RuntimeError(L.oidnb200_last_error().decode())
      dist.broadcast(self.scale, src=0, group=self.pg_exchange)   # also orders the peers after rank 0's frame upload
    else:
      dist.all_reduce(self.token, group=self.pg_exchange)   # frame-start ordering without a scale
    if self.staged:
      self._copy_rects(self.inputs, True)     # NVLink DMA: rank 0 -> local tile inputs (with overlap)
    self.filter.execute_async()
    if self.staged:
      self._copy_rects(("output",), False)    # NVLink DMA: local interior rectangles -> rank 0's output
    dist.all_reduce(self.token, group=self.pg_join)   # join: every rank's rectangles are in rank 0's output

  def release(self):
    self.filter.release()
    own = [b for n, b in self.local.items() if b is not self.bufs.get(n)]
    for b in own:
      b.release()
    self.dist.barrier()
    if self.exchange == "peer":
      for r, b in enumerate(self.xpeer):
        if r != self.rank:
          b.release()
    if self.rank != 0:
      for b in self.bufs.values():
        b.release()
    self.dist.barrier()
    if self.exchange == "peer":
      self.xb.release()
    if self.rank == 0:
      for b in self.bufs.values():
        b.release()
    for g in (self.pg_exchange, self.pg_join):
      if g is not None:
        self.dist.destroy_process_group(g)
    self.pg_exchange = self.pg_join = None


def single_process_device(args, torch, world, tza, host, K, Wm):
  """bench.py's one-process arm (called on rank 0 only): api.Device((0, .., N-1)), one filter, K frames."""
  import time
  W, H = args.width, args.height
  nb = W * H * 12
  res = {}
  # (a) the frame lives in GPU 0's HBM (single-pointer contract): every engine stages its tiles over NVLink.
  #     Engines run on caller-supplied streams here, so each frame joins the main stream before the next starts:
  #     device time per frame from CUDA events on engine 0's stream.
  streams = []
  for g in range(world):
    with torch.cuda.device(g):
      streams.append(torch.cuda.Stream())
  dev = api.Device(tuple(range(world)), streams=[s.cuda_stream for s in streams]).commit()
  with torch.cuda.device(0):
    t = {k: torch.from_numpy(host.images[k]).cuda() for k in ("color", "albedo", "normal")}
    out = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
  f = dev.new_filter("RT")
  for k, v in t.items():
    f.set_image(k, v)
  f.set_image("output", out)
  f.set("hdr", True); f.set("quality", api.QUALITY_HIGH); f.set_data("weights", tza)
  f.commit()
  for _ in range(Wm):
    f.execute_async()
  dev.sync()
  info = f.info()
  with torch.cuda.device(0):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(streams[0])
    for _ in range(K):
      f.execute_async()
    e1.record(streams[0])
  dev.sync()
  ms = e0.elapsed_time(e1) / K
  res["frame_in_gpu0_hbm"] = {"ms_per_step": round(ms, 4), "value": round(W * H / ms / 1e3, 1), "unit": "Mpix/s", "staged": info["staged"],
                              "tiles": "%dx%d of %dx%d" % (info["tileCountW"], info["tileCountH"], info["tileW"], info["tileH"]),
                              "how": "frames one after another (each joins the caller's stream): NVLink stage-in + network + stage-out "
                                     "per frame, CUDA events on engine 0's stream"}
  f.release(); dev.release()
  # (a') the same frame with the engines on their own streams: back-to-back frames pipeline inside the library
  #      (stage-in of frame f+1 / stage-out of frame f-1 under the convolutions of frame f). The frames end on internal
  #      copy-out streams, so this one is the wall clock around K frames between two oidnb200SyncDevice calls.
  dev = api.Device(tuple(range(world))).commit()
  f = dev.new_filter("RT")
  for k, v in t.items():
    f.set_image(k, v)
  f.set_image("output", out)
  f.set("hdr", True); f.set("quality", api.QUALITY_HIGH); f.set_data("weights", tza)
  f.commit()
  for _ in range(Wm):
    f.execute_async()
  dev.sync()
  t0 = time.perf_counter()
  for _ in range(K):
    f.execute_async()
  t_enq = time.perf_counter() - t0
  dev.sync()
  msp = (time.perf_counter() - t0) / K * 1e3
  res["frame_in_gpu0_hbm_pipelined"] = {"ms_per_step": round(msp, 4), "value": round(W * H / msp / 1e3, 1), "unit": "Mpix/s",
                                        "host_enqueue_ms_per_frame": round(t_enq / K * 1e3, 4),
                                        "how": "engines on their own streams, K frames enqueued back to back, wall clock between two device syncs"}
  f.release(); dev.release()
  del t, out
  # (b) end to end: the frame lives in pinned host memory; engines on their own streams, frames pipeline in the library
  import bench as B
  imgs = {k: host.images[k] for k in ("color", "albedo", "normal")}
  res["e2e_host_frame"] = B.bench_e2e_staged(api, torch, 0, imgs, tza, W, H, K, Wm, gpus=tuple(range(world)))
  return res


def bench_main(args, rank, world, local_rank):
  """bench.py's N>1 arm (launched by torchrun, one rank per GPU)."""
  import json
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
  nb = W * H * 12

  # The frame in host memory, shared by all ranks; rank 0 fills it (oidnBenchmark LCG images).
  host = SharedHostFrame(dist, torch, ("color", "albedo", "normal", "output"), H, W)
  if rank == 0 or not host.shared:
    frame = synth.benchmark_images(W, H, hdr=True, seed=1)
    for k, v in frame.items():
      host.images[k][...] = v
    del frame
  dist.barrier()
  frame = {k: host.images[k] for k in ("color", "albedo", "normal")}

  # Two frames in flight (a renderer double-buffers its frame): two device/stream/filter sets, frames
  # alternate between them, so the exchange and the copies of frame f+1 overlap the convolutions of
  # frame f. Collectives are issued in frame order by every rank.
  def make_sets(source, weights_blob=None, clean_aux=False, policy=None, exchange="peer"):
    sets = []
    for _ in range(2):
      stream = torch.cuda.Stream()
      with torch.cuda.stream(stream):
        dev = api.Device((local_rank,), streams=[stream.cuda_stream]).commit()
        if policy is not None:
          dev.set("tilePolicy", policy)
        sf = ShardedFilter(dist, torch, dev, W, H, weights_blob or tza, hdr=True, source=source, clean_aux=clean_aux,
                           exchange=exchange, frame=frame if (source == "distributed" or rank == 0) else None)
      sets.append((stream, dev, sf))
    return sets

  def release_sets(sets):
    for _, dev, sf in sets:
      sf.release(); dev.release()

  def tiles_text(sets):
    i = sets[0][2].filter.info()
    return "%dx%d of %dx%d, %d per rank" % (i["tileCountW"], i["tileCountH"], i["tileW"], i["tileH"],
                                            i["tileCountH"] * i["tileCountW"] // world)

  def conv_profile(sets, kind, frames=3):
    """rank 0's convs inside its frames: union of the conv grids' in-frame intervals (device parameter profile=2)."""
    stream, dev, sf = sets[0]
    with torch.cuda.stream(stream):
      dev.set("profile", 2)
      sf.execute_async(); torch.cuda.synchronize(); sf.filter.profile()
      for _ in range(frames):
        sf.execute_async()
      torch.cuda.synchronize()
      prof = sf.filter.profile()
      dev.set("profile", 0)
    dist.barrier()
    union = sum(m for _, k, _, m in prof if k == 3) / frames
    launches = sum(n for _, k, n, _ in prof if k == 0) // frames
    _, mine = tiles_of_rank(H, W, kind == "large", world, 0, sets[0][1].get("maxTilePixels"), sets[0][1].get("tilePolicy"))
    my_px = sum(t["H2"] * t["W2"] for t in mine)
    tf = weights.flops_per_pixel(kind, 9) * my_px / (union * 1e-3) / 1e12
    return union, launches, my_px, tf

  def timed(sets, frame_fn, wall=False):
    """K frames alternating between the two sets; device time (events) or wall clock, max over ranks."""
    sA, sB = sets[0][0], sets[1][0]
    for i in range(max(Wm, 2)):
      frame_fn(sets, i)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    join = torch.cuda.Event()
    t0 = time.time(); w0 = time.perf_counter()
    e0.record(sA)
    sB.wait_event(e0)                 # neither stream starts a timed frame before e0
    for i in range(K):
      if wall and i >= 2:
        sets[i % 2][0].synchronize()  # the host consumes frame i-2's result before its buffers are reused
      frame_fn(sets, i)
    join.record(sB)
    sA.wait_event(join)
    e1.record(sA)                     # after the last frame of both streams
    torch.cuda.synchronize(); dist.barrier()
    t1 = time.time()
    per = (time.perf_counter() - w0) * 1e3 / K if wall else e0.elapsed_time(e1) / K
    ms = torch.tensor([per], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item()), t0, t1

  def run_frame(sets, i):
    stream, _, sf = sets[i % 2]
    with torch.cuda.stream(stream):
      sf.execute_async()

  # the tile plan: fewest recomputed pixels (tilePolicy 1) against fewest 128-pixel conv strips (tilePolicy 2); the
  # headline is the device default, the other one is reported next to it
  default_policy = api.Device((local_rank,)).get("tilePolicy")
  other_policy = 2 if default_policy == 1 else 1
  plan_alt = None
  if not args.no_rank0:
    alt = make_sets("distributed", policy=other_policy)
    ms_alt, _, _ = timed(alt, run_frame)
    plan_alt = {"tilePolicy": other_policy, "tiles": tiles_text(alt), "ms_per_step": round(ms_alt, 4),
                "value": round(W * H / ms_alt / 1e3, 1), "unit": "Mpix/s"}
    release_sets(alt)

  # round 1's exchange (NCCL all-reduce of the bin array + 4-byte all-reduce join), next to the headline
  nccl_alt = None
  if not args.no_rank0:
    alt = make_sets("distributed", exchange="nccl")
    ms_alt, _, _ = timed(alt, run_frame)
    nccl_alt = {"ms_per_step": round(ms_alt, 4), "value": round(W * H / ms_alt / 1e3, 1), "unit": "Mpix/s",
                "how": "same tiles; bins by NCCL all-reduce, frame join by a 4-byte all-reduce"}
    release_sets(alt)

  sets = make_sets("distributed")
  info = sets[0][2].filter.info()
  ntiles = info["tileCountH"] * info["tileCountW"]
  ms, t0, t1 = timed(sets, run_frame)
  clocks = sampler.window(t0, t1) if rank == 0 else None
  conv_ms, conv_launches, my_px, conv_tf = conv_profile(sets, "base", frames=max(K // 4, 3))

  # end to end: the frame is in (shared, pinned) host memory and the result returns there; every rank
  # moves the rectangles of its own tiles over its own PCIe link
  e2e = None
  if not args.no_e2e:
    hin = {k: host.ptr(k) for k in ("color", "albedo", "normal")}
    hout = host.ptr("output")
    my_in = sum(t["H1"] * t["W1"] for t in sets[0][2].tiles) * 36
    my_out = sum(t["H2"] * t["W2"] for t in sets[0][2].tiles) * 12

    def e2e_frame(sets, i):
      stream, _, sf = sets[i % 2]
      with torch.cuda.stream(stream):
        sf.upload_tiles(hin)
        sf.execute_async(assemble=False)
        sf.download_tiles(hout)

    dt, _, _ = timed(sets, e2e_frame, wall=True)
    traffic = torch.tensor([float(my_in), float(my_out)], device="cuda", dtype=torch.float64)
    dist.all_reduce(traffic)
    e2e = {"value": round(W * H / dt / 1e3, 1), "unit": "Mpix/s", "ms_per_step": round(dt, 4),
           "h2d_bytes_per_step": int(traffic[0].item()), "d2h_bytes_per_step": int(traffic[1].item()),
           "host_frame": "one frame in POSIX shared memory, page-locked by every rank" if host.shared
                         else "private pinned copy of the frame per rank (/dev/shm too small)",
           "how": "every rank: 2D copies of its tiles' source rectangles (with overlap) from the pinned host frame, sharded "
                  "execute (autoexposure bins by peer writes), 2D copies of its output rectangles into the host output frame; "
                  "all ranks' bytes summed; two frame sets alternating; wall clock, max over ranks"}
  nbins = sets[0][2].nbins
  release_sets(sets)

  # the frame held by rank 0 alone (single-pointer contract): reported next to the headline
  on_rank0 = None
  if not args.no_rank0:
    sets0 = make_sets("rank0")
    ms0, _, _ = timed(sets0, run_frame)
    on_rank0 = {"value": round(W * H / ms0 / 1e3, 1), "unit": "Mpix/s", "ms_per_step": round(ms0, 4),
                "how": "whole frame in rank 0's HBM; peers pull tile rectangles / push output rectangles with copy engines over NVLink"}
    release_sets(sets0)

  # BASELINE config 3: the large UNet (cleanAux=true + quality=high: core/unet_filter.cpp:417-436,449-452) on the
  # same 7680x4320 frame, tile-sharded the same way
  large = None
  if not args.no_8k:
    tza_large = weights.model_tza("large", 9, seed=0)
    setsL = make_sets("distributed", weights_blob=tza_large, clean_aux=True)
    assert setsL[0][2].filter.info()["largeModel"] == 1
    msL, _, _ = timed(setsL, run_frame)
    cL, nL, _, tfL = conv_profile(setsL, "large")
    large = {"workload": "RT hdr+calb+cnrm 7680x4320 quality=high (large UNet, 19 convs), tile-sharded across %d ranks" % world,
             "ms_per_step": round(msL, 4), "value": round(W * H / msL / 1e3, 1), "unit": "Mpix/s", "tiles": tiles_text(setsL),
             "rank0_conv_ms_per_frame": round(cL, 4), "rank0_conv_launches": nL, "rank0_conv_tflops": round(tfL, 1),
             "rank0_conv_frac_of_burst_peak": round(tfL / peaks["bf16_tflops"], 4)}
    release_sets(setsL)

  # ONE process, ONE device object over all the GPUs: oidnb200NewCUDADevice(ids, streams, N) -- the reference's own
  # multi-pair signature (include/OpenImageDenoise/oidn.h:150-153). No torch.distributed / NCCL on this path: the
  # library stages every engine's tiles over NVLink (frame in GPU 0's HBM) or over each GPU's PCIe link (frame in
  # pinned host memory) and exchanges the autoexposure bins through peer stores. The other ranks idle at a barrier.
  single = None
  torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
  # the other ranks wait on the rendezvous store (host side): an NCCL barrier would park a spinning kernel on
  # their GPUs, which rank 0 is about to use
  store = dist.distributed_c10d._get_default_store()
  if rank == 0:
    if not args.no_single_process:
      try:
        single = single_process_device(args, torch, world, tza, host, K, max(Wm, 2))
      except Exception as e:  # noqa: BLE001
        single = {"error": str(e)[:300]}
    store.set("oidnb200_single_process_done", "1")
  else:
    store.wait(["oidnb200_single_process_done"])
  dist.barrier()

  if rank == 0:
    sampler.stop()
    roofline = {"bound": "tensor", "kernel": "conv3x3_tc_kernel on rank 0 (%d launches/frame)" % conv_launches,
                "achieved": round(conv_tf, 1), "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": round(conv_tf / peaks["bf16_tflops"], 4), "traffic": None,
                "peak_source": peaks_src + ", burst dense bf16 (the timed region is %.0f ms)" % (ms * K),
                "conv_ms_per_frame": round(conv_ms, 4), "share_of_step": round(min(conv_ms / ms, 1.0), 4),
                "how": "union of rank 0's conv grids' in-frame intervals (%globaltimer stamps), algorithmic FLOP of the pixels rank 0 outputs",
                "alg_flop_per_launch_avg": weights.flops_per_pixel("base", 9) * my_px / max(conv_launches, 1),
                "avg_launch_ms": round(conv_ms / max(conv_launches, 1), 5)}
    line = {
      "metric": B.METRIC, "value": round(W * H / (ms * 1e-3) / 1e6, 1), "unit": "Mpix/s", "n_gpus": world, "steps": K, "warmup": Wm,
      "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
      "dtype": "f16", "accumulate": "f32", "data": "synthetic",
      "config": dict(B.workload_config(args, world), tiles="%dx%d of %dx%d, %d per rank" % (
        info["tileCountW"], info["tileCountH"], info["tileW"], info["tileH"], ntiles // world)),
      "roofline": roofline, "cpu_baseline": None, "e2e": e2e, "frame_on_rank0": on_rank0,
      "tile_plan_alternative": plan_alt, "large_unet_8k": large, "single_process_device": single,
      # all ranks: per tile the filter's ops + one autoexposure-bins launch, per rank one reduce
      # all ranks, per tile: autoexposure bins + input process + the conv launches (a fused pair is one; the output
      # process runs in the last pair's epilogue); per rank one fold of the bin array
      # and four flag launches (bins ready: signal + wait, frame done: signal + wait); per rank and peer one bin-scatter
      # copy kernel per tile
      "gpu_launches": K * (ntiles * (2 + conv_launches // max(ntiles // world, 1)) + 5 * world + (world - 1) * ntiles), "clocks": clocks,
      "exchange": "every rank holds its tiles' inputs (tile + overlap); autoexposure: per-tile bin kernels, the tiles' bin rectangles "
                  "go to every rank's bin array (%d B) by peer stores of a copy kernel, fixed-order fold on every rank; output rectangles "
                  "assembled in rank 0's buffer by copy-engine peer writes over NVLink (CUDA IPC); frame steps joined with peer flags "
                  "(one 32-thread block, no collective on the data path); two frames in flight" % (4 * nbins),
      "exchange_nccl": nccl_alt,
    }
    print(json.dumps(line))
  host.release()
  dist.barrier()
  dist.destroy_process_group()
