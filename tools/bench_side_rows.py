"""Timings of the two rows next to the path: the batch crop augmentation kernel and the device
evaluation sweep at the driver's test batch (32 x batch_size = 1024, depth_train.py:86,186-253)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import lsps_b200  # noqa
from lsps_b200 import _lib  # noqa
from lsps_b200.augment import CropAugmenter, Camera, NYU_CAMERA, com_to_transform, sample_params, AugSample  # noqa

out = {}
# ---- augmentation: kernel alone (parameters resident) and end to end with the host parameter draw
n = 2048
cam = Camera(*NYU_CAMERA)
rng = np.random.RandomState(1)
imgs = (torch.rand(n, 128, 128, device="cuda") * 2 - 1).contiguous()
recs = (AugSample * n)()
t0 = time.perf_counter()
for i in range(n):
    com = np.array([320.0 + rng.uniform(-20, 20), 240.0 + rng.uniform(-20, 20), 600.0 + rng.uniform(0, 100)])
    M = com_to_transform(com, (300.0,) * 3, cam)
    gt = (rng.randn(36, 3) * 40).astype(np.float32)
    recs[i] = sample_params(gt, com, (300.0, 300.0, 300.0), np.asarray(M, np.float32), ["com", "rot", "sc", "none"], cam, rng)[0]
host_s = time.perf_counter() - t0
raw = torch.frombuffer(bytearray(bytes(recs)), dtype=torch.uint8).cuda()
premax, o = torch.empty(n, device="cuda"), torch.empty_like(imgs)
ctx = _lib.context(0)
for _ in range(3):
    ctx.augment_crops(imgs.data_ptr(), raw.data_ptr(), premax.data_ptr(), o.data_ptr(), n)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ctx.augment_crops(imgs.data_ptr(), raw.data_ptr(), premax.data_ptr(), o.data_ptr(), n)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
out["augment_crops"] = {"crops": n, "kernel_ms": ms, "crops_per_s_kernel": n / ms * 1e3,
                        "algorithmic_GBps": n * 2 * 65536 / ms / 1e6, "host_param_us_per_crop": host_s / n * 1e6,
                        "note": "algorithmic bytes = 64 KB in + 64 KB out per crop; the reference pipeline costs ~750 us per crop per host core"}
# ---- evaluation sweep at test batch 1024
hp = lsps_b200.load_hyperparameters("nnyu")
tr = lsps_b200.LSPSTrainerB200(hp, device=0, noise="device")
ev = lsps_b200.PoseEvaluator(tr, domain="b", restricted_joints=lsps_b200.NYU_RESTRICTED_JOINTS)
x = torch.rand(1024, 1, 128, 128, device="cuda") * 2 - 1
y = torch.randn(1024, 108, device="cuda") * 0.3
cube = torch.tensor([300.0, 300.0, 300.0])
for _ in range(3):
    ev.reset(); ev.add_batch(x, y, cube); ev.summary()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    ev.reset(); ev.add_batch(x, y, cube); ev.summary()
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) * 1e3 / 10
out["eval_sweep_b1024"] = {"ms_per_batch": ms, "images_per_s": 1024 / ms * 1e3, "precision": tr.precision,
                           "note": "regress_b -> vae.decode -> per-frame joint errors -> (mean mm, % within 40 mm), one host read"}
print(json.dumps(out, indent=1))
