"""One launch of every kernel family at its in-step shape, inside a cudaProfilerStart/Stop range, for
   ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_targets python tools/ncu_targets.py
The launch order printed on stdout is the order of the kernels in the report (tools/ncu_summary.py reads both).
Select families with LSPS_NCU_CASES=conv,k1,stem,head,in,adam,noise,norm,split (default: conv,stem,head,in,adam,noise)."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lsps_b200 import _lib  # noqa
from lsps_b200._lib import ConvShape  # noqa

ctx = _lib.context(0)
want = set(os.environ.get("LSPS_NCU_CASES", "conv,stem,head,in,adam,noise").split(","))
cases = []  # (label, algorithmic bytes, algorithmic flop, fn)


def conv_cases(name, kind, n, h, cin, cout):
    ho = h if kind == 0 else (h // 2 if kind == 1 else 2 * h)
    x = torch.randn(n, h, h, cin, device="cuda").bfloat16()
    dy = torch.randn(n, ho, ho, cout, device="cuda").bfloat16()
    wf = (torch.randn(9, cout, cin, device="cuda") * 0.05).bfloat16()
    wd = wf.transpose(1, 2).contiguous()
    b = torch.zeros(cout, device="cuda")
    y, dx = torch.empty_like(dy), torch.empty_like(x)
    dw = torch.zeros(9, cout, cin, device="cuda")
    sh = ConvShape(kind, n, h, h, cin, cout)
    flop = 2.0 * n * (ho * ho if kind != 2 else h * h) * cin * cout * 9
    io = (x.numel() + dy.numel() + wf.numel()) * 2
    keep = (x, dy, wf, wd, b, y, dx, dw, sh)
    cases.append((name + " fwd", io, flop, lambda: ctx.conv_fwd(C.byref(sh), x.data_ptr(), wf.data_ptr(), b.data_ptr(), y.data_ptr(), 3, 0.01), keep))
    cases.append((name + " dgrad+mask", io + x.numel() * 2, flop, lambda: ctx.conv_dgrad(C.byref(sh), dy.data_ptr(), wd.data_ptr(), dx.data_ptr(), x.data_ptr(), None, 4, 0.01), keep))
    cases.append((name + " wgrad", (x.numel() + dy.numel()) * 2 + dw.numel() * 4, flop, lambda: ctx.conv_wgrad(C.byref(sh), x.data_ptr(), dy.data_ptr(), dw.data_ptr()), keep))


if "conv" in want:
    conv_cases("s2 64->128 @128^2 N=128", 1, 128, 128, 64, 128)
    conv_cases("s2 128->256 @64^2 N=128", 1, 128, 64, 128, 256)
    conv_cases("deconv 256->128 @32^2 N=128", 2, 128, 32, 256, 128)
    conv_cases("deconv 128->64 @64^2 N=128", 2, 128, 64, 128, 64)
    conv_cases("s2 1024->2048 @4^2 N=384 (dis tail)", 1, 384, 4, 1024, 2048)
if "k1" in want:
    conv_cases("K1 s1 256->256 @32^2 N=128", 0, 128, 32, 256, 256)
def _sec_stem():
    for stride, n in ((1, 128), (2, 192)):
        img = torch.rand(n, 128, 128, device="cuda") * 2 - 1
        ho = 128 // stride
        w = torch.randn(64, 49, device="cuda") * 0.02
        b = torch.zeros(64, device="cuda")
        y = torch.empty(n, ho, ho, 64, device="cuda", dtype=torch.bfloat16)
        dy = torch.randn(n, ho, ho, 64, device="cuda").bfloat16()
        dw, db, dimg = torch.zeros(64, 49, device="cuda"), torch.zeros(64, device="cuda"), torch.zeros_like(img)
        keep = (img, w, b, y, dy, dw, db, dimg)
        io = img.numel() * 4 + y.numel() * 2
        fl = 2.0 * n * ho * ho * 64 * 49
        cases.append(("stem s%d N=%d fwd" % (stride, n), io, fl, lambda img=img, w=w, b=b, y=y, n=n, stride=stride: ctx.stem_fwd(img.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), n, 128, 128, stride, 0.01), keep))
        cases.append(("stem s%d N=%d wgrad" % (stride, n), io, fl, lambda img=img, dy=dy, dw=dw, db=db, n=n, stride=stride: ctx.stem_wgrad(img.data_ptr(), dy.data_ptr(), dw.data_ptr(), db.data_ptr(), n, 128, 128, stride), keep))
        cases.append(("stem s%d N=%d dgrad" % (stride, n), io, fl, lambda dy=dy, w=w, dimg=dimg, n=n, stride=stride: ctx.stem_dgrad(dy.data_ptr(), w.data_ptr(), dimg.data_ptr(), n, 128, 128, stride, 0), keep))


if "stem" in want:
    _sec_stem()
def _sec_head():
    n = 128
    g2 = torch.randn(n, 128, 128, 64, device="cuda").bfloat16()
    w, b = torch.randn(64, device="cuda") * 0.02, torch.zeros(1, device="cuda")
    out, dout = torch.empty(n, 128, 128, device="cuda"), torch.randn(n, 128, 128, device="cuda")
    dg2, dw, db = torch.empty_like(g2), torch.zeros(64, device="cuda"), torch.zeros(1, device="cuda")
    keep = (g2, w, b, out, dout, dg2, dw, db)
    cases.append(("head fwd N=128", g2.numel() * 2 + out.numel() * 4, 2.0 * out.numel() * 64, lambda: ctx.head_fwd(g2.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), out.numel()), keep))
    cases.append(("head bwd N=128", g2.numel() * 4 + out.numel() * 8, 4.0 * out.numel() * 64, lambda: ctx.head_bwd(g2.data_ptr(), w.data_ptr(), out.data_ptr(), dout.data_ptr(), dg2.data_ptr(), dw.data_ptr(), db.data_ptr(), out.numel(), 0.01), keep))


if "head" in want:
    _sec_head()
def _sec_in():
    n, hw, c = 128, 1024, 256
    h = torch.randn(n, hw, c, device="cuda").bfloat16()
    res, dy = torch.randn_like(h), torch.randn_like(h)
    y, dh = torch.empty_like(h), torch.empty_like(h)
    stats, db = torch.empty(n, c, 2, device="cuda"), torch.zeros(c, device="cuda")
    keep = (h, res, dy, y, dh, stats, db)
    mb = h.numel() * 2
    cases.append(("instnorm fwd lrelu N=128", 2 * mb, 0, lambda: ctx.instnorm_fwd(h.data_ptr(), None, y.data_ptr(), stats.data_ptr(), n, hw, c, 0, 1e-5, 0.01), keep))
    cases.append(("instnorm fwd residual N=128", 3 * mb, 0, lambda: ctx.instnorm_fwd(h.data_ptr(), res.data_ptr(), y.data_ptr(), stats.data_ptr(), n, hw, c, 1, 1e-5, 0.01), keep))
    cases.append(("instnorm bwd N=128", 3 * mb, 0, lambda: ctx.instnorm_bwd(dy.data_ptr(), h.data_ptr(), stats.data_ptr(), dh.data_ptr(), n, hw, c, 1, 0.01, db.data_ptr()), keep))


if "in" in want:
    _sec_in()
def _sec_adam():
    nn_ = 18004482 // 256 * 256
    p, g_, m, v = (torch.randn(nn_, device="cuda") * 0.01 for _ in range(4))
    v = v.abs()
    w16 = torch.empty(nn_, device="cuda", dtype=torch.bfloat16)
    keep = (p, g_, m, v, w16)
    cases.append(("adam 18.0M params", nn_ * (16 + 14), 0, lambda: ctx.adam(p.data_ptr(), g_.data_ptr(), m.data_ptr(), v.data_ptr(), w16.data_ptr(), nn_, 1e-4, 0.5, 0.999, 1e-8, 1e-4, 3, 1.0, None), keep))


if "adam" in want:
    _sec_adam()
def _sec_noise():
    n = 128 * 1024 * 256
    x = torch.randn(n, device="cuda").bfloat16()
    nz = torch.randn(n, device="cuda")
    z, acc = torch.empty_like(x), torch.zeros(4, device="cuda")
    keep = (x, nz, z, acc)
    cases.append(("noise_kl N=128", n * 8, 0, lambda: ctx.noise_kl_fwd(x.data_ptr(), nz.data_ptr(), z.data_ptr(), acc.data_ptr(), n), keep))



if "noise" in want:
    _sec_noise()
def _sec_norm():
    from lsps_b200._lib import ConvExt
    n, hw, c = 128, 1024, 256
    h = torch.randn(n, hw, c, device="cuda").bfloat16()
    res, dy, a1 = torch.randn_like(h), torch.randn_like(h), torch.randn_like(h)
    y, dh = torch.empty_like(h), torch.empty_like(h)
    sums = torch.zeros(n, 2, c, device="cuda")
    sums[:, 1] = 1024.0
    stats, bs = torch.zeros(n, 2, c, device="cuda"), torch.zeros(n, 2, c, device="cuda")
    stats[:, 1] = 1.0
    keep = (h, res, dy, a1, y, dh, sums, stats, bs)
    mb = h.numel() * 2
    cases.append(("norm_apply_fwd lrelu N=128", 2 * mb, 0, lambda: ctx.norm_apply_fwd(h.data_ptr(), None, y.data_ptr(), sums.data_ptr(), stats.data_ptr(), n, hw, c, 0, 1, 1e-5, 0.01, None, None), keep))
    cases.append(("norm_apply_fwd residual N=128", 3 * mb, 0, lambda: ctx.norm_apply_fwd(h.data_ptr(), res.data_ptr(), y.data_ptr(), sums.data_ptr(), stats.data_ptr(), n, hw, c, 1, 1, 1e-5, 0.01, None, None), keep))
    cases.append(("norm_bwd_stats N=128", 2 * mb, 0, lambda: ctx.norm_bwd_stats(dy.data_ptr(), h.data_ptr(), stats.data_ptr(), bs.data_ptr(), n, hw, c, 1, 1, 0.01, None, None), keep))
    cases.append(("norm_bwd_apply N=128", 3 * mb, 0, lambda: ctx.norm_bwd_apply(dy.data_ptr(), h.data_ptr(), stats.data_ptr(), bs.data_ptr(), dh.data_ptr(), n, hw, c, 2, 1, 0.01, None, None), keep))
    # K1 with the statistics / InstanceNorm-backward epilogues
    x4 = h.reshape(n, 32, 32, c)
    wf = (torch.randn(9, c, c, device="cuda") * 0.05).bfloat16()
    b = torch.zeros(c, device="cuda")
    sh = ConvShape(0, n, 32, 32, c, c)
    flop = 2.0 * n * 1024 * c * c * 9
    e1, e2 = ConvExt(), ConvExt()
    e1.sums = sums.data_ptr()
    e2.in_a, e2.bsums = a1.data_ptr(), bs.data_ptr()
    keep2 = keep + (wf, b, sh, e1, e2)
    cases.append(("K1 fwd + EP_STATS N=128", 2 * mb, flop, lambda: ctx.conv_fwd_ex(C.byref(sh), x4.data_ptr(), wf.data_ptr(), b.data_ptr(), y.data_ptr(), 1 | 16, 0.01, C.byref(e1)), keep2))
    cases.append(("K1 dgrad + EP_INBWD N=128", 3 * mb, flop, lambda: ctx.conv_dgrad_ex(C.byref(sh), dy.data_ptr(), wf.data_ptr(), dh.data_ptr(), None, None, 32, 0.01, C.byref(e2)), keep2))
    cases.append(("K1 dgrad plain N=128", 2 * mb, flop, lambda: ctx.conv_dgrad(C.byref(sh), dy.data_ptr(), wf.data_ptr(), dh.data_ptr(), None, None, 0, 0.01), keep2))


if "norm" in want:
    _sec_norm()


def _sec_split():
    from lsps_b200._lib import ConvExt
    for name, n, h, cin, cout in (("split s2 64->128 @64^2 N=192", 192, 64, 64, 128), ("split s2 128->256 @32^2 N=384", 384, 32, 128, 256),
                                  ("split s2 1024->2048 @4^2 N=384", 384, 4, 1024, 2048)):
        x = torch.randn(n, h, h, 2 * cin, device="cuda").bfloat16()
        dy = torch.randn(n, h // 2, h // 2, 2 * cout, device="cuda").bfloat16()
        wf, wl = (torch.randn(9, cout, cin, device="cuda") * 0.05).bfloat16(), (torch.randn(9, cout, cin, device="cuda") * 1e-4).bfloat16()
        b = torch.zeros(cout, device="cuda")
        y, dx = torch.empty_like(dy), torch.empty_like(x)
        dw = torch.zeros(9, cout, cin, device="cuda")
        sh = ConvShape(1, n, h, h, cin, cout)
        ext = ConvExt()
        ext.split, ext.w_lo = 1, wl.data_ptr()
        flop = 3 * 2.0 * n * (h // 2) ** 2 * cin * cout * 9
        io = (x.numel() + dy.numel() + 2 * wf.numel()) * 2
        keep = (x, dy, wf, wl, b, y, dx, dw, sh, ext)
        cases.append((name + " fwd", io, flop, lambda sh=sh, x=x, wf=wf, b=b, y=y, ext=ext: ctx.conv_fwd_ex(C.byref(sh), x.data_ptr(), wf.data_ptr(), b.data_ptr(), y.data_ptr(), 3, 0.01, C.byref(ext)), keep))
        cases.append((name + " dgrad+mask", io + x.numel() * 2, flop, lambda sh=sh, x=x, wf=wf, dy=dy, dx=dx, ext=ext: ctx.conv_dgrad_ex(C.byref(sh), dy.data_ptr(), wf.data_ptr(), dx.data_ptr(), x.data_ptr(), None, 4, 0.01, C.byref(ext)), keep))
        cases.append((name + " wgrad", io, flop, lambda sh=sh, x=x, dy=dy, dw=dw: ctx.conv_wgrad_split(C.byref(sh), x.data_ptr(), dy.data_ptr(), dw.data_ptr()), keep))


if "split" in want:
    _sec_split()

for case_ in cases:      # warm-up: sets kernel attributes, fills the tensor-map cache
    case_[3]()
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
times = []
for case_ in cases:
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); case_[3](); e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
torch.cuda.profiler.start()
for case_ in cases:
    case_[3]()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("| # | case | algorithmic MB | GFLOP | cold event ms | GB/s | TFLOP/s |\n|---|---|---|---|---|---|---|")
for i, (case_, ms) in enumerate(zip(cases, times)):
    print("| %d | %s | %.1f | %.1f | %.4f | %.0f | %.0f |" % (i, case_[0], case_[1] / 1e6, case_[2] / 1e9, ms, case_[1] / ms / 1e6, case_[2] / ms / 1e9))
