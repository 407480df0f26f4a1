"""CPU tier, world_size 2 over gloo: the host-side plumbing of the multi-process tile sharding --
every rank derives the same plan, the ranks' tile sets partition the frame, the handle exchange and
the frame join run over the process group."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
  s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
  return p


def _worker(rank, world, port, q):
  sys.path.insert(0, ROOT)
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  from oidn_b200 import sharded
  out = {}
  for (H, W, large) in ((4320, 7680, False), (4320, 7680, True), (2160, 3840, False), (1000, 3000, False)):
    plan, mine = sharded.tiles_of_rank(H, W, large, world, rank)
    # every rank must see the same plan
    plans = [None] * world
    dist.all_gather_object(plans, plan)
    assert all(p == plans[0] for p in plans)
    assert (plan["tileCountH"] * plan["tileCountW"]) % world == 0
    cover = np.zeros((H, W), np.int32)
    for t in mine:
      cover[t["hDst"]:t["hDst"] + t["H2"], t["wDst"]:t["wDst"] + t["W2"]] += 1
    total = torch.from_numpy(cover)
    dist.all_reduce(total)                     # the join: all ranks' rectangles together
    assert int(total.min()) == 1 and int(total.max()) == 1, "ranks' output rectangles must partition the frame"
    out[(H, W, large)] = len(mine)
    # autoexposure bins: the ranks' bin rectangles partition the bin grid (sharded.bins_of_tile)
    nbh, nbw = sharded.bin_grid(H, W)
    bcover = np.zeros((nbh, nbw), np.int32)
    for t in mine:
      bh0, bh1, bw0, bw1 = sharded.bins_of_tile(t, H, W)
      bcover[bh0:bh1, bw0:bw1] += 1
    btotal = torch.from_numpy(bcover)
    dist.all_reduce(btotal)
    assert int(btotal.min()) == 1 and int(btotal.max()) == 1, "ranks' bin rectangles must partition the bin grid"
    # staged exchange (ShardedFilter stage=True) with an identity "filter": a rank copies its input
    # rectangles from the owner's frame into zeroed local staging, produces its interior rectangles
    # from staging only, and copies them back; the joined output must equal the frame.
    if H * W <= 3840 * 2160:
      frame = np.arange(H * W, dtype=np.float32).reshape(H, W) % 1021.0
      local = np.zeros_like(frame)
      for t in mine:
        local[t["hSrc"]:t["hSrc"] + t["H1"], t["wSrc"]:t["wSrc"] + t["W1"]] = \
          frame[t["hSrc"]:t["hSrc"] + t["H1"], t["wSrc"]:t["wSrc"] + t["W1"]]
      result = np.zeros_like(frame)
      for t in mine:
        result[t["hDst"]:t["hDst"] + t["H2"], t["wDst"]:t["wDst"] + t["W2"]] = \
          local[t["hDst"]:t["hDst"] + t["H2"], t["wDst"]:t["wDst"] + t["W2"]]
      joined = torch.from_numpy(result)
      dist.all_reduce(joined)
      assert np.array_equal(joined.numpy(), frame)
      # the autoexposure exchange of the distributed frame: every rank fills the bins of its tiles from
      # its LOCAL pixels into a zero array, the all-reduce (sum) completes it: x + 0 = x exactly
      def bin_value(img, bi, bj):
        h0, h1 = bi * H // nbh, (bi + 1) * H // nbh
        w0, w1 = bj * W // nbw, (bj + 1) * W // nbw
        return np.float32(img[h0:h1, w0:w1].mean(dtype=np.float64))
      step = max(1, nbh // 6), max(1, nbw // 8)     # a sample of the grid keeps the test fast
      mine_bins = np.zeros((nbh, nbw), np.float32)
      for t in mine:
        bh0, bh1, bw0, bw1 = sharded.bins_of_tile(t, H, W)
        for bi in range(bh0, bh1):
          for bj in range(bw0, bw1):
            if bi % step[0] == 0 and bj % step[1] == 0:
              mine_bins[bi, bj] = bin_value(local, bi, bj)
      allb = torch.from_numpy(mine_bins)
      dist.all_reduce(allb)
      for bi in range(0, nbh, step[0]):
        for bj in range(0, nbw, step[1]):
          assert allb[bi, bj].item() == bin_value(frame, bi, bj), (bi, bj)
      # the same exchange without a collective (the default): each rank's bin rectangles are copied as 2D byte
      # rectangles into every peer's array (sharded.bin_rect_copies gives the copies execute_async issues)
      rects = [sharded.bins_of_tile(t, H, W) for t in mine]
      copies = [None] * world
      dist.all_gather_object(copies, (sharded.bin_rect_copies(rects, nbw), mine_bins.tobytes()))
      peer_view = bytearray(mine_bins.tobytes())
      for r, (cps, raw) in enumerate(copies):
        if r == rank:
          continue
        for off, pitch, wbytes, rows in cps:
          for y in range(rows):
            peer_view[off + y * pitch: off + y * pitch + wbytes] = raw[off + y * pitch: off + y * pitch + wbytes]
      assert np.array_equal(np.frombuffer(bytes(peer_view), np.float32).reshape(nbh, nbw), allb.numpy())
  # handle exchange: rank 0's 64-byte handles reach every rank unchanged
  handles = {"color": bytes(range(64)), "output": bytes(range(64, 128))} if rank == 0 else None
  got = sharded.broadcast_object(dist, handles, 0)
  assert got["color"] == bytes(range(64)) and got["output"] == bytes(range(64, 128))
  # scale broadcast + join token, as in ShardedFilter.execute_async
  scale = torch.tensor([0.125 if rank == 0 else -1.0])
  dist.broadcast(scale, src=0)
  assert float(scale) == 0.125
  q.put((rank, out))
  dist.barrier()
  dist.destroy_process_group()


def test_world_size_2_tile_sharding():
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
  for p in procs:
    p.start()
  res = [q.get(timeout=180) for _ in procs]
  for p in procs:
    p.join(60)
    assert p.exitcode == 0
  counts = dict(res)
  assert counts[0] == counts[1], "round-robin dealing gives both ranks the same number of tiles"
