"""Secondary timings (not the bench.py contract line): estimate3 / estimate0 post_update and pose-VAE vae_update.
Prints wall-clock ms per step (host launch overhead included) next to CUDA-event ms (device time)."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lsps_b200  # noqa

hp = lsps_b200.load_hyperparameters("nnyu")
out = {}
for graphs in (False, True):
  tr = lsps_b200.LSPSTrainerB200(hp, device=0, seed=0, noise="device", graphs=graphs)
  for name, B, fn in (
          ("estimate3_b32", 32, lambda a, la, b, lb: tr.post_update(a, la, b, lb, None, None, 3, hp)),
          ("estimate3_b256", 256, lambda a, la, b, lb: tr.post_update(a, la, b, lb, None, None, 3, hp)),
          ("estimate0_b32", 32, lambda a, la, b, lb: tr.post_update(a, la, b, lb, None, None, 0, hp)),
          ("estimate0_b256", 256, lambda a, la, b, lb: tr.post_update(a, la, b, lb, None, None, 0, hp)),
          ("vae_b64", 64, lambda a, la, b, lb: tr.vae_update(torch.cat((la, lb), 0), hp))):
      g = torch.Generator().manual_seed(1)
      ia, ib, la, lb = (t.cuda() for t in lsps_b200.synthetic_batch(B, 108, g, "hand"))
      for _ in range(5):
          fn(ia, la, ib, lb)
      torch.cuda.synchronize()
      K = 20
      l0 = tr.ops.ctx.launch_count()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      t0 = time.perf_counter()
      e0.record()
      for _ in range(K):
          fn(ia, la, ib, lb)
      e1.record()
      torch.cuda.synchronize()
      wall = (time.perf_counter() - t0) * 1e3 / K
      out[name + ("_graphs" if graphs else "")] = {"wall_ms": wall, "event_ms": e0.elapsed_time(e1) / K, "images_per_s": 2 * B / (wall / 1e3),
                   "launches": (tr.ops.ctx.launch_count() - l0) // K}
print(json.dumps(out, indent=1))
