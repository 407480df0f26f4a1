"""torchrun check (N ranks): a frame denoised tile-sharded across the ranks is bit-identical to the
same frame denoised by rank 0 alone with the same tile plan, and matches the CPU oracle within the
path's tolerance. Run: torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/sharded_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from oidn_b200 import api, capi, sharded, synth, weights  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H = 2400, 1700
tza = weights.model_tza("base", 9, seed=0)
full = synth.benchmark_images(W, H, hdr=True, seed=21)   # deterministic: every rank generates the same frame
frame = full if rank == 0 else None
stream = torch.cuda.Stream()
ok = True
with torch.cuda.stream(stream):
  dev = api.Device((local,), streams=[stream.cuda_stream]).commit()
  dev.set("maxTilePixels", 1000 * 1000)   # force several tiles per rank
  # both exchange modes: copy-engine staging of the tile rectangles, and direct P2P loads/stores
  outs = []
  for stage in (False, True):
    sf0 = sharded.ShardedFilter(dist, torch, dev, W, H, tza, hdr=True, frame=frame, stage=stage)
    for _ in range(2):
      sf0.execute_async()
    torch.cuda.synchronize(); dist.barrier()
    if rank == 0:
      o = np.zeros((H, W, 3), np.float32); sf0.bufs["output"].read(o); outs.append(o)
    if stage:
      sf = sf0
    else:
      sf0.release()
  # distributed frame: every rank holds only its own tiles' source rectangles (the rest of its local
  # images is garbage), autoexposure from every rank's bins, output assembled on rank 0 over NVLink
  # -- by peer writes of the bin rectangles + peer flags (default) and by NCCL all-reduce; both must give the same bits
  garbage = np.full((H, W, 3), 1e30, np.float32)
  scale_d = []
  for exchange in ("nccl", "peer"):
    sfd = sharded.ShardedFilter(dist, torch, dev, W, H, tza, hdr=True, source="distributed", exchange=exchange)
    for n in sfd.inputs:
      sfd.local[n].write(garbage)
    sfd.upload_tiles({n: full[n].ctypes.data for n in sfd.inputs})
    for _ in range(5):
      sfd.execute_async()
    torch.cuda.synchronize(); dist.barrier()
    if rank == 0:
      o = np.zeros((H, W, 3), np.float32); sfd.bufs["output"].read(o); outs.append(o)
    scale_d.append(float(sfd.scale.cpu()[0]))
    sfd.release()
  info = sf.filter.info()
  if rank == 0:
    got = outs[1]
    same_d = all(np.array_equal(o.view(np.uint32), outs[1].view(np.uint32)) for o in outs[2:])
    print("distributed frame (nccl, peer exchange) == frame on rank 0: %s (autoexposure scale %.9g, %.9g vs %.9g)"
          % (same_d, scale_d[0], scale_d[1], float(sf.scale.cpu()[0])))
    ok = np.array_equal(outs[0].view(np.uint32), outs[1].view(np.uint32))
    print("staged == direct P2P: %s" % ok)
    ok = ok and same_d
    # same plan on one GPU: numShards=world keeps the tile grid, but this filter runs every tile
    t = {k: torch.from_numpy(v).cuda() for k, v in frame.items()}
    out = torch.zeros((H, W, 3), device="cuda")
    dev1 = api.Device((local,)).commit(); dev1.set("maxTilePixels", 1000 * 1000)
    f = dev1.new_filter("RT")
    for k, v in t.items():
      f.set_image(k, v)
    f.set_image("output", out); f.set("hdr", True); f.set_data("weights", tza); f.commit(); f.execute()
    single = out.cpu().numpy()
    same = np.array_equal(got.view(np.uint32), single.view(np.uint32))
    import oracle as orc
    ref = np.zeros((H, W, 3), np.float32)
    orc.filter_execute(tza, color=frame["color"], albedo=frame["albedo"], normal=frame["normal"], output=ref, hdr=True)
    peak = np.abs(ref).max(); err = np.abs(got - ref).max() / peak
    psnr = 20 * np.log10(peak / np.sqrt(np.mean((got - ref) ** 2)))
    print("sharded_check world=%d tiles=%dx%d: bit-identical to single GPU: %s; vs oracle max|err|/peak=%.3e PSNR=%.1f dB"
          % (world, info["tileCountW"], info["tileCountH"], same, err, psnr))
    ok = ok and same and err <= 1e-2 and psnr >= 50
    f.release(); dev1.release()
  sf.release(); dev.release()
flag = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.broadcast(flag, 0)
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if float(flag) == 1.0 else 1)
