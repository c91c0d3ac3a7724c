"""Cin=1 7x7 stem kernels (forward, weight gradient, data gradient; plain and split-bf16) on the shapes of the step:
generator stems 128x128 stride 1 (64 images per call), discriminator stems 128x128 stride 2 (192 images per call).
Burst timings (3 warm-up + 20 launches) and the HBM bytes each call has to move at least."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lsps_b200 import _lib  # noqa

ctx = _lib.context(0)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    print("| case | fwd us | GB/s | wgrad us | GB/s | dgrad us | GB/s |")
    print("|---|---|---|---|---|---|---|")
    for name, stride, n, split in (("gen stem s1 n64", 1, 64, 0), ("gen stem s1 n128", 1, 128, 0),
                                   ("dis stem s2 n192 split", 2, 192, 1), ("dis stem s2 n128 split", 2, 128, 1),
                                   ("dis stem s2 n192", 2, 192, 0)):
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
        f = getattr(ctx, "stem_fwd" + sfx)
        wg = getattr(ctx, "stem_wgrad" + sfx)
        dg = getattr(ctx, "stem_dgrad" + sfx)
        t_f = timeit(lambda: f(img.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), n, h, h, stride, 0.01))
        t_w = timeit(lambda: wg(img.data_ptr(), dy.data_ptr(), dw.data_ptr(), db.data_ptr(), n, h, h, stride))
        t_d = timeit(lambda: dg(dy.data_ptr(), w.data_ptr(), dimg.data_ptr(), n, h, h, stride, 0))
        byts = img.numel() * 4 + y.numel() * 2
        print("| %s | %.1f | %.0f | %.1f | %.0f | %.1f | %.0f |" % (name, t_f, byts / t_f / 1e3, t_w, byts / t_w / 1e3, t_d,
                                                                     byts / t_d / 1e3))


if __name__ == "__main__":
    main()
