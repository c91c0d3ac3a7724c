"""Secondary timing (not the bench.py contract line): the pretrain step with train_map=True (Mapping net, ndiv=4
discriminator batch, map losses -- SURVEY 8f n1) next to the plain pretrain step, B=64 per domain, device noise."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lsps_b200  # noqa

B = int(os.environ.get("LSPS_MAP_B", "64"))
out = {}
for name, tm in (("pretrain", False), ("pretrain_train_map", True)):
    hp = dict(lsps_b200.load_hyperparameters("nnyu"), train_map=tm)
    tr = lsps_b200.LSPSTrainerB200(hp, device=0, seed=0, noise="device")
    g = torch.Generator().manual_seed(1)
    ia, ib, la, lb = (t.cuda() for t in lsps_b200.synthetic_batch(B, 108, g, "hand"))

    def step():
        tr.dis_update(ia, la, ib, lb, None, None, hp)
        tr.gen_update(ia, la, ib, lb, hp)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    K = 5
    l0 = tr.ops.ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3 / K
    out[name] = {"batch_per_domain": B, "wall_ms": wall, "event_ms": e0.elapsed_time(e1) / K,
                 "images_per_s": 2 * B / (wall / 1e3), "launches": (tr.ops.ctx.launch_count() - l0) // K,
                 "losses": {k: float(getattr(tr, k)) for k in ("gen_total_loss", "dis_loss", "gen_map_loss", "gen_map_loss2")
                            if hasattr(tr, k)}}
    del tr
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
