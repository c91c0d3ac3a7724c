"""One launch of each stem kernel at the step's shapes, for `ncu --set full -k regex:stem_`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lsps_b200 import _lib  # noqa

ctx = _lib.context(0)
for stride, n, split in ((1, 64, 0), (2, 192, 1)):
    h = 128
    ho = h // stride
    oc = 128 if split else 64
    img = torch.rand(n, h, h, device="cuda") * 2 - 1
    w = torch.randn(64, 49, device="cuda") * 0.05
    b = torch.randn(64, device="cuda") * 0.1
    y = torch.empty(n, ho, ho, oc, device="cuda", dtype=torch.bfloat16)
    dy = torch.randn(n, ho, ho, oc, device="cuda").bfloat16()
    dw, db = torch.zeros(64, 49, device="cuda"), torch.zeros(64, device="cuda")
    dimg = torch.zeros(n, h, h, device="cuda")
    sfx = "_split" if split else ""
    for _ in range(2):
        getattr(ctx, "stem_fwd" + sfx)(img.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), n, h, h, stride, 0.01)
        getattr(ctx, "stem_wgrad" + sfx)(img.data_ptr(), dy.data_ptr(), dw.data_ptr(), db.data_ptr(), n, h, h, stride)
        getattr(ctx, "stem_dgrad" + sfx)(dy.data_ptr(), w.data_ptr(), dimg.data_ptr(), n, h, h, stride, 0)
    torch.cuda.synchronize()
