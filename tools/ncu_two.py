"""Two launches for `ncu --set full`: the fused up-sampling kernel (deconv 128->64 @64^2 N=128 forward) and the
resident-weight kernel (stride-2 64->128 @128^2 N=128 forward)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lsps_b200 import _lib  # noqa
from lsps_b200._lib import ConvShape  # noqa

ctx = _lib.context(0)
for kind, n, h, cin, cout in ((2, 128, 64, 128, 64), (1, 128, 128, 64, 128)):
    ho = h // 2 if kind == 1 else 2 * h
    x = torch.randn(n, h, h, cin, device="cuda").bfloat16()
    wf = (torch.randn(9, cout, cin, device="cuda") * 0.05).bfloat16()
    b = torch.zeros(cout, device="cuda")
    y = torch.empty(n, ho, ho, cout, device="cuda", dtype=torch.bfloat16)
    sh = C.byref(ConvShape(kind, n, h, h, cin, cout))
    for _ in range(3):
        ctx.conv_fwd(sh, x.data_ptr(), wf.data_ptr(), b.data_ptr(), y.data_ptr(), 3, 0.01)
    torch.cuda.synchronize()
