"""In-step (hot, power-limited) time per C entry point: CUDA events around every lsps_* call of ONE pretrain step
(B=64/domain, side stream off so nothing overlaps).  Complements the ncu launch list, whose per-launch times are
cold-cache, serialised and taken at burst clocks."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lsps_b200  # noqa
from lsps_b200 import _lib  # noqa

B = int(os.environ.get("LSPS_PROFILE_B", "64"))
hp = dict(lsps_b200.load_hyperparameters("nnyu"), train_map=os.environ.get("LSPS_PROFILE_MAP", "0") == "1")
tr = lsps_b200.LSPSTrainerB200(hp, device=0, seed=0, noise="device")
g = torch.Generator().manual_seed(1)
ia, ib, la, lb = (t.cuda() for t in lsps_b200.synthetic_batch(B, 108, g, "hand"))


MODE = os.environ.get("LSPS_PROFILE_MODE", "pretrain")     # pretrain | estimateN


def step():
    if MODE == "pretrain":
        tr.dis_update(ia, la, ib, lb, None, None, hp)
        tr.gen_update(ia, la, ib, lb, hp)
    else:
        tr.post_update(ia, la, ib, lb, None, None, int(MODE[len("estimate"):]), hp)


for _ in range(8 if MODE == "pretrain" else 30):          # long enough to reach the power-limited steady state
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    step()
e1.record()
torch.cuda.synchronize()
plain_ms = e0.elapsed_time(e1) / 5

tr.ops.use_side = False
ev = []
ctx = tr.ops.ctx
names = [n[5:] for n in _lib._SIGS]
for n in names:
    orig = getattr(ctx, n)

    def wrapped(*a, _o=orig, _n=n):
        if _n in ("conv_fwd", "conv_dgrad", "conv_wgrad", "conv_fwd_ex", "conv_dgrad_ex", "conv_wgrad_split", "conv_wgrad_grouped"):
            sh = a[0]._obj
            tag = "%s k%d %dx%d %d->%d n%d" % (_n, sh.kind, sh.h, sh.w, sh.cin, sh.cout, sh.n)
        else:
            tag = _n
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        _o(*a)
        e.record()
        ev.append((tag, s, e))
    ctx.__dict__[n] = wrapped
for _ in range(2):
    ev.clear()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    step()
    t1.record()
    torch.cuda.synchronize()
tot = t0.elapsed_time(t1)
rows = {}
for tag, s, e in ev:
    r = rows.setdefault(tag, [0, 0.0])
    r[0] += 1
    r[1] += s.elapsed_time(e)
inside = sum(r[1] for r in rows.values())
print("step %.2f ms (side stream on, uninstrumented) ; instrumented inline step %.2f ms, %.2f ms inside lsps_* calls, "
      "%.2f ms torch glue / gaps" % (plain_ms, tot, inside, tot - inside))
print("| entry point (conv: kind HxW cin->cout N) | calls | total ms | share | avg us |\n|---|---|---|---|---|")
for tag, (c, ms) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
    print("| %s | %d | %.3f | %.1f%% | %.1f |" % (tag, c, ms, 100 * ms / tot, 1e3 * ms / c))
