"""GPU probe: InstanceNorm fwd/bwd kernel timings at the K1 activation shape (N=128, 32x32, 256 ch)."""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lsps_b200 import _lib
ctx = _lib.context(0)
n, hw, c = 128, 1024, 256
h = torch.randn(n, hw, c, device="cuda").bfloat16()
res = torch.randn(n, hw, c, device="cuda").bfloat16()
dy = torch.randn(n, hw, c, device="cuda").bfloat16()
y = torch.empty_like(h); dh = torch.empty_like(h)
stats = torch.empty(n, c, 2, device="cuda"); db = torch.zeros(c, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def t(fn, name, gb):
    for _ in range(3): fn()
    tot = 0.0
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / 10
    print("%-22s %.1f us  %.0f GB/s algorithmic" % (name, ms * 1e3, gb / ms * 1e3 / 1e3))
mb = n * hw * c * 2 / 1e6
t(lambda: ctx.instnorm_fwd(h.data_ptr(), None, y.data_ptr(), stats.data_ptr(), n, hw, c, 0, 1e-5, 0.01), "in_fwd lrelu", 2 * mb)
t(lambda: ctx.instnorm_fwd(h.data_ptr(), res.data_ptr(), y.data_ptr(), stats.data_ptr(), n, hw, c, 1, 1e-5, 0.01), "in_fwd residual", 3 * mb)
t(lambda: ctx.instnorm_bwd(dy.data_ptr(), h.data_ptr(), stats.data_ptr(), dh.data_ptr(), n, hw, c, 0, 0.01, None), "in_bwd", 3 * mb)
t(lambda: ctx.instnorm_bwd(dy.data_ptr(), h.data_ptr(), stats.data_ptr(), dh.data_ptr(), n, hw, c, 1, 0.01, db.data_ptr()), "in_bwd + db", 3 * mb)
