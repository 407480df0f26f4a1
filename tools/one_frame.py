"""Two 4K RT hdr+alb+nrm frames through the filter C ABI (the workload of bench.py) -- a short command for ncu:
the second frame's launches are the ones to capture (-s <launches of one frame> -c <the same>)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oidn_b200 import api, synth, weights  # noqa: E402

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3840, 2160)
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 2
tza = weights.model_tza("base", 9, seed=0)
imgs = synth.benchmark_images(W, H, hdr=True, seed=1)
dev = api.Device((0,)).commit()
t = {k: torch.from_numpy(v).cuda() for k, v in imgs.items()}
out = torch.zeros((H, W, 3), dtype=torch.float32, device="cuda")
f = dev.new_filter("RT")
for k, v in t.items():
  f.set_image(k, v)
f.set_image("output", out)
f.set("hdr", True); f.set("quality", api.QUALITY_HIGH); f.set_data("weights", tza)
f.commit()
for _ in range(frames):
  f.execute()
print("ops per frame:", f.info()["numOps"])
f.release(); dev.release()
