"""Per-kernel parity through the C ABI (ctypes) against plain fp32 torch ops on the same GPU.
conv kernels use bf16 operands / fp32 accumulate: inputs are pre-rounded to bf16 so the only difference from the
fp32 reference is the bf16 rounding of the stored output (2^-9 relative) and accumulation order."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF16_L2 = 4e-3     # relative L2 error budget of a bf16-stored result
SLOPE = 0.01


@pytest.fixture(scope="module")
def ctx():
    from lsps_b200 import _lib
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return _lib.context(0)


def rel_l2(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-20)).item()


def nhwc16(t):
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def nchw32(t):
    return t.float().permute(0, 3, 1, 2)


def gen(seed=0):
    return torch.Generator(device="cuda").manual_seed(seed)


CONV_CASES = [  # kind, n, h, w, cin, cout
    (0, 2, 32, 32, 256, 256), (0, 3, 32, 32, 64, 64), (0, 1, 16, 16, 128, 128),
    (1, 2, 128, 128, 64, 128), (1, 3, 64, 64, 128, 256), (1, 5, 16, 16, 256, 512), (1, 9, 8, 8, 512, 1024),
    (1, 40, 4, 4, 1024, 2048), (1, 1, 4, 4, 1024, 2048),
    (2, 2, 32, 32, 256, 128), (2, 2, 64, 64, 128, 64), (2, 3, 8, 8, 64, 64),
    # the fused four-phase up-sampling kernel (N = 64, K <= 128) on more than one tile per CTA and an odd image count
    (2, 37, 64, 64, 128, 64), (1, 21, 128, 128, 64, 128), (1, 5, 32, 32, 64, 64),
    # large enough for the CTA-pair (cta_group::2) conv kernels: >= 2*148 M tiles and >= 27 K-steps per tile
    (0, 40, 32, 32, 256, 256),        # K1 shape, 320 tiles: pair kernel with 128-channel stages
    (1, 297, 16, 16, 256, 512),       # dis trunk shape, 149 M tiles (odd): the pair's phantom tile path
    # kind 3 = 4x4 stride-2 pad-1 ConvTranspose2d (Mapping net layers 1-3, lsps_nets.py:19-23) + odd / larger batches
    (3, 2, 4, 4, 1024, 1024), (3, 2, 8, 8, 1024, 512), (3, 2, 16, 16, 512, 256), (3, 5, 4, 4, 128, 64),
    (3, 150, 16, 16, 512, 256),     # >= 296 tiles: CTA-pair kernels with 16 taps / 4 phases
]


@pytest.mark.parametrize("kind,n,h,w,cin,cout", CONV_CASES)
def test_conv_fwd_dgrad_wgrad(ctx, kind, n, h, w, cin, cout):
    from lsps_b200._lib import ConvShape
    g = gen(kind * 100 + n)
    x = torch.randn(n, cin, h, w, device="cuda", generator=g).bfloat16().float().requires_grad_(True)
    taps = 16 if kind == 3 else 9
    if kind == 3:
        wt = (torch.randn(cin, cout, 4, 4, device="cuda", generator=g) * 0.05).bfloat16().float().requires_grad_(True)
        pack = lambda t: t.permute(2, 3, 1, 0).reshape(16, cout, cin)
        y = F.conv_transpose2d(x, wt, None, stride=2, padding=1)
    elif kind == 2:
        wt = (torch.randn(cin, cout, 3, 3, device="cuda", generator=g) * 0.05).bfloat16().float().requires_grad_(True)
        pack = lambda t: t.permute(2, 3, 1, 0).reshape(9, cout, cin)
        y = F.conv_transpose2d(x, wt, None, stride=2, padding=1, output_padding=1)
    else:
        wt = (torch.randn(cout, cin, 3, 3, device="cuda", generator=g) * 0.05).bfloat16().float().requires_grad_(True)
        pack = lambda t: t.permute(2, 3, 0, 1).reshape(9, cout, cin)
        y = F.conv2d(x, wt, None, stride=1 if kind == 0 else 2, padding=1)
    bias = torch.randn(cout, device="cuda", generator=g)
    dy = torch.randn(y.shape, device="cuda", generator=g).bfloat16().float()
    y.backward(dy)
    sh = C.byref(ConvShape(kind, n, h, w, cin, cout))
    xb, dyb = nhwc16(x.detach()), nhwc16(dy)
    wf = pack(wt.detach()).contiguous().bfloat16()
    wd = pack(wt.detach()).transpose(1, 2).contiguous().bfloat16()
    yb = torch.empty(n, y.shape[2], y.shape[3], cout, device="cuda", dtype=torch.bfloat16)
    ctx.conv_fwd(sh, xb.data_ptr(), wf.data_ptr(), bias.data_ptr(), yb.data_ptr(), 3, SLOPE)
    assert rel_l2(nchw32(yb), F.leaky_relu(y.detach() + bias[None, :, None, None], SLOPE)) < BF16_L2
    mask = nhwc16(torch.randn(n, cin, h, w, device="cuda", generator=g))
    add = nhwc16(torch.randn(n, cin, h, w, device="cuda", generator=g))
    dxb = torch.empty(n, h, w, cin, device="cuda", dtype=torch.bfloat16)
    ctx.conv_dgrad(sh, dyb.data_ptr(), wd.data_ptr(), dxb.data_ptr(), None, None, 0, SLOPE)
    assert rel_l2(nchw32(dxb), x.grad) < BF16_L2
    ctx.conv_dgrad(sh, dyb.data_ptr(), wd.data_ptr(), dxb.data_ptr(), mask.data_ptr(), add.data_ptr(), 12, SLOPE)
    ref = (x.grad + nchw32(add)) * torch.where(nchw32(mask) > 0, 1.0, SLOPE)
    assert rel_l2(nchw32(dxb), ref) < BF16_L2
    ctx.conv_dgrad(sh, dyb.data_ptr(), wd.data_ptr(), dxb.data_ptr(), mask.data_ptr(), None, 4, SLOPE)    # mask only
    assert rel_l2(nchw32(dxb), x.grad * torch.where(nchw32(mask) > 0, 1.0, SLOPE)) < BF16_L2
    dw = torch.zeros(taps, cout, cin, device="cuda")
    ctx.conv_wgrad(sh, xb.data_ptr(), dyb.data_ptr(), dw.data_ptr())
    assert rel_l2(dw, pack(wt.grad)) < 1e-4
    ctx.conv_wgrad(sh, xb.data_ptr(), dyb.data_ptr(), dw.data_ptr())       # accumulates
    assert rel_l2(dw, 2 * pack(wt.grad)) < 1e-4
    db = torch.zeros(cout, device="cuda")
    ctx.colsum_bf16(dyb.data_ptr(), dyb.numel() // cout, cout, db.data_ptr())
    assert rel_l2(db, dy.sum((0, 2, 3))) < 1e-4


@pytest.mark.parametrize("n,split,order", [(6, 3, 0), (5, 2, 1), (80, 40, 0)])
def test_conv_grouped_two_weight_sets(ctx, n, split, order):
    """lsps_conv_{fwd,dgrad}_grouped: images [0, split) through conv A, the rest through conv B, in one launch
    (n=80: the CTA-pair kernel); the second set may sit before or after the first in the flat weight buffer."""
    from lsps_b200._lib import ConvShape
    g = gen(50 + n)
    c = 256
    x = torch.randn(n, c, 32, 32, device="cuda", generator=g).bfloat16().float()
    dy = torch.randn(n, c, 32, 32, device="cuda", generator=g).bfloat16().float()
    ws = [(torch.randn(c, c, 3, 3, device="cuda", generator=g) * 0.05).bfloat16().float() for _ in range(2)]
    bs = [torch.randn(c, device="cuda", generator=g) for _ in range(2)]
    pack = lambda t: t.permute(2, 3, 0, 1).reshape(9, c, c)
    per = 9 * c * c
    gap = 5 * 256                                   # some other tensor in between (multiple of the 256-element rows)
    flat_f = torch.zeros(2 * per + gap, device="cuda", dtype=torch.bfloat16)
    flat_d = torch.zeros(2 * per + gap, device="cuda", dtype=torch.bfloat16)
    offs = (0, per + gap) if order == 0 else (per + gap, 0)
    for i in range(2):
        flat_f[offs[i]:offs[i] + per] = pack(ws[i]).reshape(-1).bfloat16()
        flat_d[offs[i]:offs[i] + per] = pack(ws[i]).transpose(1, 2).reshape(-1).bfloat16()
    sh = C.byref(ConvShape(0, n, 32, 32, c, c))
    xb, dyb = nhwc16(x), nhwc16(dy)
    yb = torch.empty_like(xb)
    ctx.conv_fwd_grouped(sh, xb.data_ptr(), flat_f[offs[0]:].data_ptr(), bs[0].data_ptr(), flat_f[offs[1]:].data_ptr(),
                         bs[1].data_ptr(), split, yb.data_ptr(), 1, SLOPE)
    ref = torch.cat((F.conv2d(x[:split], ws[0], bs[0], padding=1), F.conv2d(x[split:], ws[1], bs[1], padding=1)), 0)
    assert rel_l2(nchw32(yb), ref) < BF16_L2
    assert rel_l2(nchw32(yb)[split:], ref[split:]) < BF16_L2 and rel_l2(nchw32(yb)[:split], ref[:split]) < BF16_L2
    dxb = torch.empty_like(xb)
    ctx.conv_dgrad_grouped(sh, dyb.data_ptr(), flat_d[offs[0]:].data_ptr(), flat_d[offs[1]:].data_ptr(), split,
                           dxb.data_ptr(), None, None, 0, SLOPE)
    refd = torch.cat((F.conv_transpose2d(dy[:split], ws[0], padding=1), F.conv_transpose2d(dy[split:], ws[1], padding=1)), 0)
    assert rel_l2(nchw32(dxb)[:split], refd[:split]) < BF16_L2 and rel_l2(nchw32(dxb)[split:], refd[split:]) < BF16_L2


def test_instnorm_bwd_grouped_bias_gradients(ctx):
    g = gen(61)
    n, c, hw, split = 5, 256, 1024, 2
    h = torch.randn(n, hw, c, device="cuda", generator=g).bfloat16()
    dy = torch.randn(n, hw, c, device="cuda", generator=g).bfloat16()
    y, stats = torch.empty_like(h), torch.empty(n, c, 2, device="cuda")
    ctx.instnorm_fwd(h.data_ptr(), None, y.data_ptr(), stats.data_ptr(), n, hw, c, 0, 1e-5, SLOPE)
    dh, dh2 = torch.empty_like(h), torch.empty_like(h)
    db = torch.zeros(c, device="cuda")
    ctx.instnorm_bwd(dy.data_ptr(), h.data_ptr(), stats.data_ptr(), dh.data_ptr(), n, hw, c, 0, SLOPE, db.data_ptr())
    da, dbb = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
    ctx.instnorm_bwd_grouped(dy.data_ptr(), h.data_ptr(), stats.data_ptr(), dh2.data_ptr(), n, hw, c, 0, SLOPE,
                             da.data_ptr(), dbb.data_ptr(), split)
    assert torch.equal(dh, dh2)
    assert torch.allclose(da, dh[:split].float().sum((0, 1)), atol=1e-3)
    assert torch.allclose(dbb, dh[split:].float().sum((0, 1)), atol=1e-3)


def test_l2_bf16(ctx):
    g = gen(5)
    a = torch.randn(3, 32, 32, 256, device="cuda", generator=g).bfloat16()
    b = torch.randn(3, 32, 32, 256, device="cuda", generator=g).bfloat16()
    gr = torch.empty_like(a)
    acc = torch.zeros(2, device="cuda")
    ctx.l2_bf16(a.data_ptr(), b.data_ptr(), gr.data_ptr(), 0.25, acc.data_ptr(), a.numel())
    d = a.float() - b.float()
    assert abs(acc[0].item() - (d * d).sum().item()) <= 1e-5 * (d * d).sum().item()
    assert rel_l2(gr, 0.25 * d) < BF16_L2 and acc[1].item() == 0.0


@pytest.mark.parametrize("stride,n,size", [(1, 3, 128), (2, 5, 128), (1, 24, 128), (2, 40, 128), (1, 37, 64), (2, 7, 256)])
def test_stem(ctx, stride, n, size):
    """n = 24 / 40 / 37: several tiles per CTA (the register-prefetch pipeline and both accumulators in use)"""
    g = gen(stride)
    img = (torch.rand(n, 1, size, size, device="cuda", generator=g) * 2 - 1).requires_grad_(True)
    w = (torch.randn(64, 1, 7, 7, device="cuda", generator=g) * 0.05).requires_grad_(True)
    b = (torch.randn(64, device="cuda", generator=g) * 0.1).requires_grad_(True)
    pre = F.conv2d(img, w, b, stride=stride, padding=3)
    y = F.leaky_relu(pre, SLOPE)
    ho = size // stride
    yb = torch.empty(n, ho, ho, 64, device="cuda", dtype=torch.bfloat16)
    ctx.stem_fwd(img.data_ptr(), w.data_ptr(), b.data_ptr(), yb.data_ptr(), n, size, size, stride, SLOPE)
    assert rel_l2(nchw32(yb), y) < BF16_L2
    dpre = torch.randn(pre.shape, device="cuda", generator=g).bfloat16().float()
    pre.backward(dpre)
    dw, db = torch.zeros(64, 49, device="cuda"), torch.zeros(64, device="cuda")
    dyb = nhwc16(dpre)
    ctx.stem_wgrad(img.data_ptr(), dyb.data_ptr(), dw.data_ptr(), db.data_ptr(), n, size, size, stride)
    assert rel_l2(dw, w.grad.reshape(64, 49)) < 1e-4 and rel_l2(db, b.grad) < 1e-4
    dimg = torch.full((n, size, size), 0.5, device="cuda")
    # tensor-core dgrad: bf16 weights (like every other conv operand), fp32 accumulation -> weight-rounding noise 2^-9
    ctx.stem_dgrad(dyb.data_ptr(), w.data_ptr(), dimg.data_ptr(), n, size, size, stride, 0)
    assert rel_l2(dimg, img.grad[:, 0]) < BF16_L2
    ctx.stem_dgrad(dyb.data_ptr(), w.data_ptr(), dimg.data_ptr(), n, size, size, stride, 1)
    assert rel_l2(dimg, 2 * img.grad[:, 0]) < BF16_L2


def test_head(ctx):
    g = gen(3)
    n = 3
    x = torch.randn(n, 64, 128, 128, device="cuda", generator=g).bfloat16().float().requires_grad_(True)
    w = (torch.randn(64, 1, 1, 1, device="cuda", generator=g) * 0.1).requires_grad_(True)
    b = torch.tensor([0.05], device="cuda", requires_grad=True)
    out = torch.tanh(F.conv_transpose2d(x, w, b))
    xb = nhwc16(x.detach())
    o = torch.empty(n, 128, 128, device="cuda")
    ctx.head_fwd(xb.data_ptr(), w.data_ptr(), b.data_ptr(), o.data_ptr(), o.numel())
    assert rel_l2(o, out[:, 0]) < 1e-5
    dout = torch.randn(n, 128, 128, device="cuda", generator=g)
    out.backward(dout[:, None])
    dx = torch.empty_like(xb)
    dw, db = torch.zeros(64, device="cuda"), torch.zeros(1, device="cuda")
    ctx.head_bwd(xb.data_ptr(), w.data_ptr(), o.data_ptr(), dout.data_ptr(), dx.data_ptr(), dw.data_ptr(), db.data_ptr(),
                 o.numel(), SLOPE)
    ref = x.grad * torch.where(x.detach() > 0, 1.0, SLOPE)
    assert rel_l2(nchw32(dx), ref) < BF16_L2
    assert rel_l2(dw, w.grad.reshape(-1)) < 1e-4 and rel_l2(db, b.grad) < 1e-4


# (3, 256, 32): the LeakyINSResBlock latent -> register-resident forward; the other shapes -> generic three-pass kernels
@pytest.mark.parametrize("n,c,hw", [(3, 256, 32), (2, 128, 16), (2, 64, 32)])
@pytest.mark.parametrize("mode", [0, 1])
def test_instnorm(ctx, mode, n, c, hw):
    g = gen(4 + mode)
    h = (torch.randn(n, c, hw, hw, device="cuda", generator=g) * 2 + 0.3).bfloat16().float().requires_grad_(True)
    res = torch.randn(n, c, hw, hw, device="cuda", generator=g).bfloat16().float()
    v = F.instance_norm(h, eps=1e-5)
    y = F.leaky_relu(v, SLOPE) if mode == 0 else res + v
    hb, rb = nhwc16(h.detach()), nhwc16(res)
    yb = torch.empty_like(hb)
    stats = torch.empty(n, c, 2, device="cuda")
    ctx.instnorm_fwd(hb.data_ptr(), rb.data_ptr() if mode else None, yb.data_ptr(), stats.data_ptr(), n, hw * hw, c, mode,
                     1e-5, SLOPE)
    assert rel_l2(nchw32(yb), y) < BF16_L2
    assert rel_l2(stats[..., 0], h.detach().mean((2, 3))) < 1e-4
    dy = torch.randn(n, c, hw, hw, device="cuda", generator=g).bfloat16().float()
    y.backward(dy)
    dh = torch.empty_like(hb)
    db = torch.zeros(c, device="cuda")
    ctx.instnorm_bwd(nhwc16(dy).data_ptr(), hb.data_ptr(), stats.data_ptr(), dh.data_ptr(), n, hw * hw, c, mode, SLOPE,
                     db.data_ptr())
    assert rel_l2(nchw32(dh), h.grad) < BF16_L2
    assert torch.allclose(db, dh.float().sum((0, 1, 2)), atol=1e-3)      # fused bias gradient == column sum of dh
    ctx.instnorm_bwd(nhwc16(dy).data_ptr(), hb.data_ptr(), stats.data_ptr(), dh.data_ptr(), n, hw * hw, c, mode, SLOPE, None)


def test_losses_and_heads(ctx):
    g = gen(9)
    dev = "cuda"
    # noise + KL
    x = torch.randn(4096 * 8, device=dev, generator=g).bfloat16()
    nz = torch.randn(4096 * 8, device=dev, generator=g)
    z = torch.empty_like(x)
    acc = torch.zeros(4, device=dev)
    ctx.noise_kl_fwd(x.data_ptr(), nz.data_ptr(), z.data_ptr(), acc.data_ptr(), x.numel())
    assert rel_l2(z, x.float() + nz) < BF16_L2 and abs(acc[0].item() / ((x.float() + nz) ** 2).sum().item() - 1) < 1e-5
    out = torch.empty_like(x)
    ctx.axpy_bf16(x.data_ptr(), z.data_ptr(), 0.25, out.data_ptr(), x.numel())
    assert rel_l2(out, x.float() + 0.25 * z.float()) < BF16_L2
    # L1 (images)
    a, t = torch.randn(50000, device=dev, generator=g), torch.randn(50000, device=dev, generator=g)
    da = torch.ones(50000, device=dev)
    acc.zero_()
    ctx.l1_f32(a.data_ptr(), t.data_ptr(), da.data_ptr(), 0.5, 1, acc.data_ptr(), a.numel())
    assert abs(acc[0].item() / (a - t).abs().sum().item() - 1) < 1e-5
    assert torch.allclose(da, 1 + 0.5 * torch.sign(a - t))
    # D head + BCE
    rows, c = 96, 2048
    f = torch.randn(rows, c, device=dev, generator=g).bfloat16()
    w = torch.randn(c, device=dev, generator=g) * 0.02
    b = torch.tensor([0.1], device=dev)
    lg = torch.empty(rows, device=dev)
    ctx.dhead_fwd(f.data_ptr(), w.data_ptr(), b.data_ptr(), lg.data_ptr(), rows, c)
    ref_lg = f.float() @ w + b
    assert rel_l2(lg, ref_lg) < 1e-5
    for target in (1.0, 0.0):
        acc.zero_()
        dl = torch.empty(rows, device=dev)
        ctx.bce_logits(lg.data_ptr(), target, 0.125, dl.data_ptr(), acc.data_ptr(), rows)
        p = torch.sigmoid(ref_lg)
        ref = F.binary_cross_entropy(p, torch.full_like(p, target), reduction="sum")
        assert abs(acc[0].item() / ref.item() - 1) < 1e-5
        assert rel_l2(dl, 0.125 * (p - target)) < 1e-5
        cnt = (p >= 0.5).sum() if target == 1.0 else (p <= 0.5).sum()
        assert acc[1].item() == cnt.item()
    df = torch.ones(rows, c, device=dev)
    dw, db = torch.zeros(c, device=dev), torch.zeros(1, device=dev)
    ctx.dhead_bwd(f.data_ptr(), w.data_ptr(), dl.data_ptr(), df.data_ptr(), dw.data_ptr(), db.data_ptr(), rows, c)
    assert rel_l2(df, 1 + dl[:, None] * w[None]) < 1e-5 and rel_l2(dw, dl @ f.float()) < 1e-4
    assert abs(db.item() - dl.sum().item()) < 1e-5
    o16 = torch.empty(rows, c, device=dev, dtype=torch.bfloat16)
    ctx.mask_to_bf16(df.data_ptr(), f.data_ptr(), o16.data_ptr(), SLOPE, df.numel())
    assert rel_l2(o16, df * torch.where(f.float() > 0, 1.0, SLOPE)) < BF16_L2
    # feature-matching L1
    f2 = torch.randn(rows, c, device=dev, generator=g).bfloat16()
    d1, d2 = torch.zeros(rows, c, device=dev), torch.zeros(rows, c, device=dev)
    acc.zero_()
    ctx.l1_feat(f.data_ptr(), f2.data_ptr(), d1.data_ptr(), d2.data_ptr(), 0.5, acc.data_ptr(), f.numel())
    dd = f.float() - f2.float()
    assert abs(acc[0].item() / dd.abs().sum().item() - 1) < 1e-5
    assert torch.equal(d1, 0.5 * torch.sign(dd)) and torch.equal(d2, -0.5 * torch.sign(dd))


@pytest.mark.parametrize("xbf16", [0, 1])
def test_linear(ctx, xbf16):
    g = gen(11)
    m, n, k = 37, 20, 8192 if xbf16 else 108
    x = torch.randn(m, k, device="cuda", generator=g)
    if xbf16:
        x = x.bfloat16()
    w = (torch.randn(n, k, device="cuda", generator=g) * 0.05).requires_grad_(True)
    b = torch.randn(n, device="cuda", generator=g).requires_grad_(True)
    xf = x.float().requires_grad_(True)
    for act, fn in ((0, lambda t: t), (1, lambda t: F.leaky_relu(t, SLOPE)), (2, F.softplus)):
        y = torch.empty(m, n, device="cuda")
        ctx.linear_fwd(x.data_ptr(), xbf16, w.data_ptr(), b.data_ptr(), y.data_ptr(), m, n, k, act, SLOPE)
        ref = fn(F.linear(xf, w, b))
        assert rel_l2(y, ref) < 1e-5
        dy = torch.randn(m, n, device="cuda", generator=g)
        for t in (xf, w, b):
            t.grad = None
        ref.backward(dy)
        dpre = dy.clone()
        ctx.act_bwd(dpre.data_ptr(), y.data_ptr(), act, SLOPE, dpre.numel())
        dx, dw, db = torch.empty(m, k, device="cuda"), torch.zeros(n, k, device="cuda"), torch.zeros(n, device="cuda")
        ctx.linear_bwd(x.data_ptr(), xbf16, w.data_ptr(), dpre.data_ptr(), dx.data_ptr(), 0, dw.data_ptr(), db.data_ptr(),
                       m, n, k)
        assert rel_l2(dx, xf.grad) < 1e-4 and rel_l2(dw, w.grad) < 1e-4 and rel_l2(db, b.grad) < 1e-4


def test_vae_reparam_and_mse(ctx):
    g = gen(12)
    n = 16 * 20
    mu = torch.randn(n, device="cuda", generator=g).requires_grad_(True)
    sd = (torch.rand(n, device="cuda", generator=g) + 0.1).requires_grad_(True)
    nz = torch.randn(n, device="cuda", generator=g) * 0.05
    z, acc = torch.empty(n, device="cuda"), torch.zeros(2, device="cuda")
    ctx.vae_reparam(mu.data_ptr(), sd.data_ptr(), nz.data_ptr(), z.data_ptr(), acc.data_ptr(), n)
    zr = mu + sd * nz
    kl = (mu * mu + sd * sd - torch.log(sd * sd)).sum()
    assert rel_l2(z, zr) < 1e-6 and abs(acc[0].item() / kl.item() - 1) < 1e-5
    dz = torch.randn(n, device="cuda", generator=g)
    (0.3 * kl + (zr * dz).sum()).backward()
    dmu, dsd = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
    ctx.vae_reparam_bwd(mu.data_ptr(), sd.data_ptr(), nz.data_ptr(), dz.data_ptr(), dmu.data_ptr(), dsd.data_ptr(), 0.3, n)
    assert rel_l2(dmu, mu.grad) < 1e-5 and rel_l2(dsd, sd.grad) < 1e-5
    p, e = torch.randn(n, device="cuda", generator=g), torch.randn(n, device="cuda", generator=g)
    dp = torch.empty(n, device="cuda")
    acc.zero_()
    ctx.mse(p.data_ptr(), e.data_ptr(), dp.data_ptr(), 0.7, acc.data_ptr(), n)
    assert abs(acc[0].item() / ((p - e) ** 2).sum().item() - 1) < 1e-5 and rel_l2(dp, 0.7 * (p - e)) < 1e-6


def test_adam_matches_torch(ctx):
    g = gen(13)
    n = 100000
    p0 = torch.randn(n, device="cuda", generator=g)
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=1e-4, betas=(0.5, 0.999), weight_decay=1e-4)
    p, m, v = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    w16 = torch.empty(n, device="cuda", dtype=torch.bfloat16)
    for step in range(1, 4):
        gr = torch.randn(n, device="cuda", generator=g) * 0.01
        ref.grad = gr.clone()
        opt.step()
        hyper = None
        if step == 3:   # device-resident step factors (CUDA-graph replay path) must give the same update
            hyper = torch.tensor([1e-4 / (1 - 0.5 ** step), 1.0 / (1 - 0.999 ** step) ** 0.5], device="cuda")
        ctx.adam(p.data_ptr(), gr.data_ptr(), m.data_ptr(), v.data_ptr(), w16.data_ptr(), n, 1e-4, 0.5, 0.999, 1e-8,
                 1e-4, step if hyper is None else 1, 1.0, None if hyper is None else hyper.data_ptr())
        assert (p - ref.detach()).abs().max().item() < 5e-7
    assert torch.equal(w16, p.bfloat16())


def test_pack_dgrad(ctx):
    g = gen(14)
    w = torch.randn(9, 128, 64, device="cuda", generator=g)
    wt = torch.empty(9, 64, 128, device="cuda", dtype=torch.bfloat16)
    ctx.pack_dgrad(w.data_ptr(), wt.data_ptr(), 9, 128, 64)
    assert torch.equal(wt, w.transpose(1, 2).bfloat16())


# ------------------------------------------------------------------ round 2: statistics in the conv epilogues
@pytest.mark.parametrize("n,cin,cout,hw,grouped", [(3, 256, 256, 32, False), (40, 256, 256, 32, False),
                                                    (80, 256, 256, 32, True), (2, 64, 128, 16, False)])
def test_conv_epilogue_statistics_and_norm_apply(ctx, n, cin, cout, hw, grouped):
    """LSPS_EP_STATS: per-(image, channel) sum / sum of squares of the fp32 conv result from the epilogue, then
    lsps_norm_apply_fwd (lrelu / residual) against F.instance_norm on the fp32 conv output of torch."""
    from lsps_b200._lib import ConvShape, ConvExt
    g = gen(300 + n)
    x = torch.randn(n, cin, hw, hw, device="cuda", generator=g).bfloat16().float()
    ws = [(torch.randn(cout, cin, 3, 3, device="cuda", generator=g) * 0.05).bfloat16().float() for _ in range(2)]
    bs = [torch.randn(cout, device="cuda", generator=g) for _ in range(2)]
    pack = lambda t: t.permute(2, 3, 0, 1).reshape(9, cout, cin)
    flat = torch.cat([pack(w_).reshape(-1) for w_ in ws]).bfloat16().contiguous()
    per = 9 * cout * cin
    split = n // 2 if grouped else 0
    if grouped:
        ref = torch.cat((F.conv2d(x[:split], ws[0], bs[0], padding=1), F.conv2d(x[split:], ws[1], bs[1], padding=1)), 0)
    else:
        ref = F.conv2d(x, ws[0], bs[0], padding=1)
    xb = nhwc16(x)
    yb = torch.empty(n, hw, hw, cout, device="cuda", dtype=torch.bfloat16)
    sums = torch.full((n, 2, cout), 7.0, device="cuda")            # the call must zero it
    ext = ConvExt()
    ext.sums = sums.data_ptr()
    if grouped:
        ext.w2, ext.bias2, ext.n_split = flat[per:].data_ptr(), bs[1].data_ptr(), split
    ctx.conv_fwd_ex(C.byref(ConvShape(0, n, hw, hw, cin, cout)), xb.data_ptr(), flat.data_ptr(), bs[0].data_ptr(),
                    yb.data_ptr(), 1 | 16, SLOPE, C.byref(ext))
    assert rel_l2(nchw32(yb), ref) < BF16_L2
    s1, s2 = ref.sum((2, 3)), (ref * ref).sum((2, 3))
    assert (sums[:, 0] - s1).abs().max().item() < 2e-3 * (ref.std().item() * hw)      # sqrt(hw*hw) random-walk scale
    assert rel_l2(sums[:, 1], s2) < 1e-4
    # forward apply, both modes, against instance_norm of the STORED (bf16) conv output with the epilogue's statistics
    hs = nchw32(yb)
    res = torch.randn(n, cout, hw, hw, device="cuda", generator=g).bfloat16()
    stats = torch.empty(n, 2, cout, device="cuda")
    y0, y1 = torch.empty_like(yb), torch.empty_like(yb)
    ctx.norm_apply_fwd(yb.data_ptr(), None, y0.data_ptr(), sums.data_ptr(), stats.data_ptr(), n, hw * hw, cout, 0, 1, 1e-5, SLOPE, None, None)
    resb = res.permute(0, 2, 3, 1).contiguous()
    ctx.norm_apply_fwd(yb.data_ptr(), resb.data_ptr(), y1.data_ptr(), sums.data_ptr(), stats.data_ptr(), n, hw * hw, cout, 1, 1,
                       1e-5, SLOPE, None, None)
    xn = F.instance_norm(ref, eps=1e-5)
    assert rel_l2(nchw32(y0), F.leaky_relu(xn, SLOPE)) < 2 * BF16_L2
    assert rel_l2(nchw32(y1), res.float() + xn) < 2 * BF16_L2
    mean, var = ref.mean((2, 3)), ref.var((2, 3), unbiased=False)
    assert (stats[:, 0] - mean).abs().max().item() < 2e-3 * ref.std().item()
    assert rel_l2(stats[:, 1], torch.rsqrt(var + 1e-5)) < 1e-3
    del hs


@pytest.mark.parametrize("n,grouped", [(3, False), (40, False), (80, True)])
def test_instnorm_backward_through_dgrad_epilogue(ctx, n, grouped):
    """A whole LeakyINSResBlock backward front half on the new kernels: norm_bwd_stats + norm_bwd_apply for `res + IN(h2)`,
    the data gradient with LSPS_EP_INBWD for lrelu(IN(h1)), norm_bwd_apply -- against torch autograd in fp32."""
    from lsps_b200._lib import ConvShape, ConvExt
    g = gen(400 + n)
    c, hw = 256, 32
    split = n // 2 if grouped else 0
    h1 = torch.randn(n, c, hw, hw, device="cuda", generator=g).bfloat16().float().requires_grad_(True)
    ws = [(torch.randn(c, c, 3, 3, device="cuda", generator=g) * 0.05).bfloat16().float() for _ in range(2)]
    a1 = F.leaky_relu(F.instance_norm(h1, eps=1e-5), SLOPE)
    a1r = a1.detach().bfloat16().float().requires_grad_(True)           # the stored activation the convs read
    if grouped:
        h2 = torch.cat((F.conv2d(a1r[:split], ws[0], None, padding=1), F.conv2d(a1r[split:], ws[1], None, padding=1)), 0)
    else:
        h2 = F.conv2d(a1r, ws[0], None, padding=1)
    h2r = h2.detach().bfloat16().float().requires_grad_(True)
    out = F.instance_norm(h2r, eps=1e-5)
    dout = torch.randn(n, c, hw, hw, device="cuda", generator=g).bfloat16().float()
    out.backward(dout)
    dh2_ref = h2r.grad
    h2.backward(dh2_ref.bfloat16().float())
    da1_ref = a1r.grad
    a1.backward(da1_ref)
    dh1_ref = h1.grad
    # --- kernels
    def stats_of(t):
        return torch.stack((t.mean((2, 3)), torch.rsqrt(t.var((2, 3), unbiased=False) + 1e-5)), 1).contiguous()
    st1, st2 = stats_of(h1.detach()), stats_of(h2r.detach())
    h1b, h2b, doutb = nhwc16(h1.detach()), nhwc16(h2r.detach()), nhwc16(dout)
    bs = torch.full((2, n, 2, c), 3.0, device="cuda")
    ctx.norm_bwd_stats(doutb.data_ptr(), h2b.data_ptr(), st2.data_ptr(), bs[1].data_ptr(), n, hw * hw, c, 1, 1, SLOPE, None, None)
    dh2b = torch.empty_like(h2b)
    ctx.norm_bwd_apply(doutb.data_ptr(), h2b.data_ptr(), st2.data_ptr(), bs[1].data_ptr(), dh2b.data_ptr(), n, hw * hw, c, 1, 1, SLOPE, None, None)
    assert rel_l2(nchw32(dh2b), dh2_ref) < BF16_L2
    pack = lambda t: t.permute(2, 3, 0, 1).reshape(9, c, c)
    flat_d = torch.cat([pack(w_).transpose(1, 2).reshape(-1) for w_ in ws]).bfloat16().contiguous()
    ext = ConvExt()
    a1b = nhwc16(a1.detach())
    ext.in_a, ext.bsums = a1b.data_ptr(), bs[0].data_ptr()
    if grouped:
        ext.w2, ext.n_split = flat_d[9 * c * c:].data_ptr(), split
    g1 = torch.empty_like(h1b)
    dh2in = nhwc16(dh2_ref)                     # feed the reference gradient so that errors do not compound
    ctx.conv_dgrad_ex(C.byref(ConvShape(0, n, hw, hw, c, c)), dh2in.data_ptr(), flat_d.data_ptr(), g1.data_ptr(), None, None,
                      32, SLOPE, C.byref(ext))
    xh1 = F.instance_norm(h1.detach(), eps=1e-5)
    g_ref = da1_ref * torch.where(xh1 > 0, 1.0, SLOPE)
    assert rel_l2(nchw32(g1), g_ref) < BF16_L2
    assert rel_l2(bs[0][:, 0], g_ref.sum((2, 3))) < 2e-3 and rel_l2(bs[0][:, 1], (g_ref * xh1).sum((2, 3))) < 2e-3
    dh1b = torch.empty_like(h1b)
    ctx.norm_bwd_apply(g1.data_ptr(), a1b.data_ptr(), st1.data_ptr(), bs[0].data_ptr(), dh1b.data_ptr(), n, hw * hw, c, 2, 1, SLOPE, None, None)
    assert rel_l2(nchw32(dh1b), dh1_ref) < 1.5 * BF16_L2
    # un-fused fallback of the same thing: stats kernel in mode 0 + apply with gmode 0 on the raw gradient
    da1b = nhwc16(da1_ref)
    ctx.norm_bwd_stats(da1b.data_ptr(), h1b.data_ptr(), st1.data_ptr(), bs[0].data_ptr(), n, hw * hw, c, 0, 1, SLOPE, None, None)
    ctx.norm_bwd_apply(da1b.data_ptr(), h1b.data_ptr(), st1.data_ptr(), bs[0].data_ptr(), dh1b.data_ptr(), n, hw * hw, c, 0, 1, SLOPE, None, None)
    assert rel_l2(nchw32(dh1b), dh1_ref) < 1.5 * BF16_L2


# ------------------------------------------------------------------ round 2: split-bf16 ("bf16x3") kernels
SPLIT_L2 = 1e-4     # hi + lo carry 16 mantissa bits; the dropped lo*lo term and the output split are ~2^-17


def split16(t):
    """fp32 -> (hi, lo) bf16 pair and the value hi + lo they represent."""
    hi = t.bfloat16()
    lo = (t - hi.float()).bfloat16()
    return hi, lo, hi.float() + lo.float()


def nhwc_split(t):
    """NCHW fp32 -> NHWC [.., 2c] = (hi | lo) bf16"""
    hi, lo, _ = split16(t.permute(0, 2, 3, 1).contiguous())
    return torch.cat((hi, lo), 3).contiguous()


def unsplit(t):
    c = t.shape[-1] // 2
    return (t[..., :c].float() + t[..., c:].float()).permute(0, 3, 1, 2)


@pytest.mark.parametrize("kind,n,h,w,cin,cout", [(1, 6, 64, 64, 64, 128), (1, 5, 32, 32, 128, 256), (1, 40, 4, 4, 1024, 2048),
                                                 (1, 297, 16, 16, 256, 512), (0, 3, 32, 32, 256, 256), (2, 2, 32, 32, 256, 128)])
def test_split_bf16_conv_fwd_dgrad_wgrad(ctx, kind, n, h, w, cin, cout):
    from lsps_b200._lib import ConvShape, ConvExt
    g = gen(500 + n)
    x = split16(torch.randn(n, cin, h, w, device="cuda", generator=g))[2].requires_grad_(True)
    if kind == 2:
        wt = split16(torch.randn(cin, cout, 3, 3, device="cuda", generator=g) * 0.05)[2].requires_grad_(True)
        pack = lambda t: t.permute(2, 3, 1, 0).reshape(9, cout, cin)
        y = F.conv_transpose2d(x, wt, None, stride=2, padding=1, output_padding=1)
    else:
        wt = split16(torch.randn(cout, cin, 3, 3, device="cuda", generator=g) * 0.05)[2].requires_grad_(True)
        pack = lambda t: t.permute(2, 3, 0, 1).reshape(9, cout, cin)
        y = F.conv2d(x, wt, None, stride=1 if kind == 0 else 2, padding=1)
    bias = torch.randn(cout, device="cuda", generator=g)
    dy = split16(torch.randn(y.shape, device="cuda", generator=g))[2]
    y.backward(dy)
    sh = C.byref(ConvShape(kind, n, h, w, cin, cout))
    xs, dys = nhwc_split(x.detach()), nhwc_split(dy)
    wf_hi, wf_lo, _ = split16(pack(wt.detach()).contiguous())
    wd_hi, wd_lo, _ = split16(pack(wt.detach()).transpose(1, 2).contiguous())
    ys = torch.empty(n, y.shape[2], y.shape[3], 2 * cout, device="cuda", dtype=torch.bfloat16)
    ext = ConvExt()
    ext.split, ext.w_lo = 1, wf_lo.data_ptr()
    ctx.conv_fwd_ex(sh, xs.data_ptr(), wf_hi.data_ptr(), bias.data_ptr(), ys.data_ptr(), 3, SLOPE, C.byref(ext))
    assert rel_l2(unsplit(ys), F.leaky_relu(y.detach() + bias[None, :, None, None], SLOPE)) < SPLIT_L2
    mask = torch.randn(n, cin, h, w, device="cuda", generator=g)
    masks = nhwc_split(mask)
    dxs = torch.empty(n, h, w, 2 * cin, device="cuda", dtype=torch.bfloat16)
    ext.w_lo = wd_lo.data_ptr()
    ctx.conv_dgrad_ex(sh, dys.data_ptr(), wd_hi.data_ptr(), dxs.data_ptr(), masks.data_ptr(), None, 4, SLOPE, C.byref(ext))
    ref = x.grad * torch.where(mask.bfloat16().float() > 0, 1.0, SLOPE)
    assert rel_l2(unsplit(dxs), ref) < SPLIT_L2
    dw = torch.zeros(9, cout, cin, device="cuda")
    ctx.conv_wgrad_split(sh, xs.data_ptr(), dys.data_ptr(), dw.data_ptr())
    assert rel_l2(dw, pack(wt.grad)) < SPLIT_L2
    db = torch.zeros(cout, device="cuda")
    ctx.colsum_bf16_split(dys.data_ptr(), dys.numel() // (2 * cout), cout, db.data_ptr())
    assert rel_l2(db, dy.sum((0, 2, 3))) < 1e-4


@pytest.mark.parametrize("stride,n", [(2, 5), (1, 3), (2, 40), (1, 24)])
def test_split_bf16_stems(ctx, stride, n):
    g = gen(600 + stride)
    img = (torch.rand(n, 1, 128, 128, device="cuda", generator=g) * 2 - 1).requires_grad_(True)
    w = (torch.randn(64, 1, 7, 7, device="cuda", generator=g) * 0.05).requires_grad_(True)
    b = torch.randn(64, device="cuda", generator=g)
    pre = F.conv2d(img, w, b, stride=stride, padding=3)
    y = F.leaky_relu(pre, SLOPE)
    ho = 128 // stride
    dy = split16(torch.randn(n, 64, ho, ho, device="cuda", generator=g))[2]
    pre.backward(dy)
    ys = torch.empty(n, ho, ho, 128, device="cuda", dtype=torch.bfloat16)
    i3 = img.detach().reshape(n, 128, 128).contiguous()
    wk = w.detach().reshape(64, 49).contiguous()
    ctx.stem_fwd_split(i3.data_ptr(), wk.data_ptr(), b.data_ptr(), ys.data_ptr(), n, 128, 128, stride, SLOPE)
    assert rel_l2(unsplit(ys), y.detach()) < SPLIT_L2
    dys = nhwc_split(dy)
    dw, db = torch.zeros(64, 49, device="cuda"), torch.zeros(64, device="cuda")
    ctx.stem_wgrad_split(i3.data_ptr(), dys.data_ptr(), dw.data_ptr(), db.data_ptr(), n, 128, 128, stride)
    assert rel_l2(dw, w.grad.reshape(64, 49)) < SPLIT_L2 and rel_l2(db, dy.sum((0, 2, 3))) < SPLIT_L2
    dimg = torch.zeros(n, 128, 128, device="cuda")
    ctx.stem_dgrad_split(dys.data_ptr(), wk.data_ptr(), dimg.data_ptr(), n, 128, 128, stride, 0)
    assert rel_l2(dimg, img.grad.reshape(n, 128, 128)) < SPLIT_L2


def test_split_bf16_heads_and_feature_losses(ctx):
    """dhead fwd/bwd, l1_feat, mask_to_bf16 and the Post FC on split feature tensors against their plain-tensor forms."""
    g = gen(700)
    n, cf = 6, 2048
    Ff = torch.randn(n, 2, 2, cf, device="cuda", generator=g)
    hi, lo, Fv = split16(Ff)
    Fs = torch.cat((hi, lo), 3).contiguous()
    w, b = torch.randn(cf, device="cuda", generator=g) * 0.02, torch.randn(1, device="cuda", generator=g)
    rows = n * 4
    lg = torch.empty(rows, device="cuda")
    ctx.dhead_fwd_split(Fs.data_ptr(), w.data_ptr(), b.data_ptr(), lg.data_ptr(), rows, cf)
    assert rel_l2(lg, Fv.reshape(rows, cf) @ w + b) < 1e-5
    dl = torch.randn(rows, device="cuda", generator=g)
    dF, dw, db = torch.zeros(rows, cf, device="cuda"), torch.zeros(cf, device="cuda"), torch.zeros(1, device="cuda")
    ctx.dhead_bwd_split(Fs.data_ptr(), w.data_ptr(), dl.data_ptr(), dF.data_ptr(), dw.data_ptr(), db.data_ptr(), rows, cf)
    assert rel_l2(dF, dl[:, None] * w[None]) < 1e-6 and rel_l2(dw, dl @ Fv.reshape(rows, cf)) < 1e-5
    acc = torch.zeros(2, device="cuda")
    dF2 = torch.zeros(n, 4 * cf, device="cuda")
    per = 4 * cf
    ctx.l1_feat_split(Fs.data_ptr(), Fs[3:].data_ptr(), dF2.data_ptr(), dF2[3:].data_ptr(), 0.5, acc.data_ptr(), 3 * per, cf)
    d = Fv[:3] - Fv[3:]
    assert abs(acc[0].item() - d.abs().sum().item()) < 1e-4 * d.abs().sum().item()
    assert torch.equal(dF2[:3].reshape(3, 2, 2, cf), 0.5 * torch.sign(d)) and torch.equal(dF2[3:], -dF2[:3])
    out = torch.empty_like(Fs)
    dFr = torch.randn(n, per, device="cuda", generator=g)
    ctx.mask_to_bf16_split(dFr.data_ptr(), Fs.data_ptr(), out.data_ptr(), SLOPE, dFr.numel(), cf)
    ref = dFr.reshape(n, 2, 2, cf) * torch.where(hi.float() > 0, 1.0, SLOPE)
    assert rel_l2(out[..., :cf].float() + out[..., cf:].float(), ref) < SPLIT_L2
    wp, bp = torch.randn(20, per, device="cuda", generator=g) * 0.01, torch.randn(20, device="cuda", generator=g)
    p = torch.empty(n, 20, device="cuda")
    ctx.linear_fwd(Fs.data_ptr(), cf, wp.data_ptr(), bp.data_ptr(), p.data_ptr(), n, 20, per, 0, SLOPE)
    assert rel_l2(p, Fv.reshape(n, per) @ wp.t() + bp) < 1e-5
    dp = torch.randn(n, 20, device="cuda", generator=g)
    dwp, dbp, dx = torch.zeros(20, per, device="cuda"), torch.zeros(20, device="cuda"), torch.empty(n, per, device="cuda")
    ctx.linear_bwd(Fs.data_ptr(), cf, wp.data_ptr(), dp.data_ptr(), dx.data_ptr(), 0, dwp.data_ptr(), dbp.data_ptr(), n, 20, per)
    assert rel_l2(dwp, dp.t() @ Fv.reshape(n, per)) < 1e-5 and rel_l2(dx, dp @ wp) < 1e-5


def test_pack_dgrad_multi_and_split_store(ctx):
    """ParamStore(split=True): one-launch transposed copies + remainders; Adam refreshes hi and lo forward copies."""
    from lsps_b200.params import ParamStore
    ents = [("a.weight", (128, 64, 3, 3), "conv3", "conv", 576), ("a.bias", (128,), "bias", "bias", 576),
            ("t.weight", (128, 64, 3, 3), "deconv3", "conv", 576), ("m.weight", (64, 128, 4, 4), "deconv4", "conv", 1024),
            ("z.weight", (256, 128, 3, 3), "conv3", "conv", 1152)]
    S = ParamStore(ents, "cuda:0", 1e-3, 1e-4, split=True)
    S.init_(3)
    for k, e in S.entries.items():
        if e.dg_off < 0:
            continue
        taps = 16 if e.kind == "deconv4" else 9
        co, ci = (e.shape[0], e.shape[1]) if e.kind == "conv3" else (e.shape[1], e.shape[0])
        wk = S.W(k).reshape(taps, co, ci)
        hi, lo, _ = split16(wk.transpose(1, 2).contiguous())
        assert torch.equal(S.W16T(k).reshape(taps, ci, co), hi) and torch.equal(S.W16TL(k).reshape(taps, ci, co), lo), k
        hi, lo, _ = split16(wk)
        assert torch.equal(S.W16(k).reshape(taps, co, ci), hi) and torch.equal(S.W16L(k).reshape(taps, co, ci), lo), k
    S.g.normal_(generator=gen(9))
    S.adam_step()
    hi, lo, _ = split16(S.w)
    assert torch.equal(S.w16, hi) and torch.equal(S.w16l, lo)


# ------------------------------------------------------------------ round 2: ResNeXt building blocks
@pytest.mark.parametrize("n,cin,cout,hw", [(3, 256, 256, 32), (5, 256, 512, 32), (40, 512, 256, 32), (2, 64, 128, 16)])
def test_conv1x1_fwd_dgrad_wgrad(ctx, n, cin, cout, hw):
    """LSPS_CONV1X1 (the 1x1 convs of LeakyINSResNeXtBlock, common_net.py:116,122) against F.conv2d."""
    from lsps_b200._lib import ConvShape
    g = gen(800 + n)
    x = torch.randn(n, cin, hw, hw, device="cuda", generator=g).bfloat16().float().requires_grad_(True)
    wt = (torch.randn(cout, cin, 1, 1, device="cuda", generator=g) * 0.05).bfloat16().float().requires_grad_(True)
    bias = torch.randn(cout, device="cuda", generator=g)
    y = F.conv2d(x, wt, None)
    dy = torch.randn(y.shape, device="cuda", generator=g).bfloat16().float()
    y.backward(dy)
    sh = C.byref(ConvShape(4, n, hw, hw, cin, cout))
    xb, dyb = nhwc16(x.detach()), nhwc16(dy)
    wf = wt.detach().reshape(cout, cin).contiguous().bfloat16()
    wd = wf.t().contiguous()
    yb = torch.empty(n, hw, hw, cout, device="cuda", dtype=torch.bfloat16)
    ctx.conv_fwd(sh, xb.data_ptr(), wf.data_ptr(), bias.data_ptr(), yb.data_ptr(), 1, SLOPE)
    assert rel_l2(nchw32(yb), y.detach() + bias[None, :, None, None]) < BF16_L2
    dxb = torch.empty(n, hw, hw, cin, device="cuda", dtype=torch.bfloat16)
    add = nhwc16(torch.randn(n, cin, hw, hw, device="cuda", generator=g))
    ctx.conv_dgrad(sh, dyb.data_ptr(), wd.data_ptr(), dxb.data_ptr(), None, add.data_ptr(), 8, SLOPE)
    assert rel_l2(nchw32(dxb), x.grad + nchw32(add)) < BF16_L2
    dw = torch.zeros(cout, cin, device="cuda")
    ctx.conv_wgrad(sh, xb.data_ptr(), dyb.data_ptr(), dw.data_ptr())
    assert rel_l2(dw, wt.grad.reshape(cout, cin)) < 1e-4


@pytest.mark.parametrize("n,c,groups", [(3, 256, 4), (40, 256, 4), (5, 512, 8), (4, 256, 2), (3, 512, 4)])
def test_grouped_conv3x3_fwd_dgrad_wgrad(ctx, n, c, groups):
    """Grouped 3x3 stride-1 conv (ResNeXt cardinality, common_net.py:118): forward with epilogue statistics, data
    gradient, and the block-diagonal weight gradient against F.conv2d(groups=...)."""
    from lsps_b200._lib import ConvShape, ConvExt
    g = gen(900 + n + groups)
    gw, hw = c // groups, 32
    x = torch.randn(n, c, hw, hw, device="cuda", generator=g).bfloat16().float().requires_grad_(True)
    wt = (torch.randn(c, gw, 3, 3, device="cuda", generator=g) * 0.05).bfloat16().float().requires_grad_(True)
    bias = torch.randn(c, device="cuda", generator=g)
    y = F.conv2d(x, wt, None, padding=1, groups=groups)
    dy = torch.randn(y.shape, device="cuda", generator=g).bfloat16().float()
    y.backward(dy)
    sh = C.byref(ConvShape(0, n, hw, hw, c, c))
    xb, dyb = nhwc16(x.detach()), nhwc16(dy)
    wk = wt.detach().permute(2, 3, 0, 1).reshape(9, c, gw).contiguous()          # [tap][co][i]
    wf = wk.bfloat16()
    wd = wk.reshape(9, groups, gw, gw).transpose(2, 3).reshape(9, c, gw).contiguous().bfloat16()   # [tap][ci][o]
    yb = torch.empty(n, hw, hw, c, device="cuda", dtype=torch.bfloat16)
    sums = torch.empty(n, 2, c, device="cuda")
    ext = ConvExt()
    ext.groups, ext.sums = groups, sums.data_ptr()
    ctx.conv_fwd_ex(sh, xb.data_ptr(), wf.data_ptr(), bias.data_ptr(), yb.data_ptr(), 1 | 16, SLOPE, C.byref(ext))
    ref = y.detach() + bias[None, :, None, None]
    assert rel_l2(nchw32(yb), ref) < BF16_L2
    assert rel_l2(sums[:, 1], (ref * ref).sum((2, 3))) < 1e-4
    dxb = torch.empty_like(xb)
    ext2 = ConvExt()
    ext2.groups = groups
    ctx.conv_dgrad_ex(sh, dyb.data_ptr(), wd.data_ptr(), dxb.data_ptr(), None, None, 0, SLOPE, C.byref(ext2))
    assert rel_l2(nchw32(dxb), x.grad) < BF16_L2
    dw = torch.zeros(9, c, gw, device="cuda")
    ctx.conv_wgrad_grouped(sh, xb.data_ptr(), dyb.data_ptr(), dw.data_ptr(), groups)
    assert rel_l2(dw, wt.grad.permute(2, 3, 0, 1).reshape(9, c, gw)) < 1e-4


@pytest.mark.parametrize("rows,d,h,z", [(16, 108, 50, 20), (37, 108, 50, 20), (8, 48, 50, 20), (130, 63, 40, 23)])
def test_vae_step_fused_matches_autograd(ctx, rows, d, h, z):
    """lsps_vae_step (K11: forward + L1/KL losses + backward of the pose-VAE in one launch) against torch autograd of
    poseVAE.forward (lsps_nets.py:68-83) + the vae_update loss (lsps_trainer.py:59-70); several CTAs and a partial one."""
    g = gen(900 + rows)
    mk = lambda *s_: (torch.randn(*s_, device="cuda", generator=g) * 0.2).requires_grad_(True)
    W1, b1, Wmu, bmu, Wsg, bsg, Wd1, bd1, Wd2, bd2 = (mk(h, d), mk(h), mk(z, h), mk(z), mk(z, h), mk(z), mk(h, z), mk(h),
                                                       mk(d, h), mk(d))
    params = [W1, b1, Wmu, bmu, Wsg, bsg, Wd1, bd1, Wd2, bd2]
    y = torch.rand(rows, d, device="cuda", generator=g) * 2 - 1
    noise = torch.randn(rows, z, device="cuda", generator=g) * 0.05
    ll_scale, kl_scale = 1.0 / (rows * d), 0.001 / rows
    hh = F.leaky_relu(y @ W1.t() + b1, SLOPE)
    mu, sd = hh @ Wmu.t() + bmu, F.softplus(hh @ Wsg.t() + bsg)
    zz = mu + sd * noise
    dec_ref = F.leaky_relu(zz @ Wd1.t() + bd1, SLOPE) @ Wd2.t() + bd2
    kl_ref, l1_ref = (mu * mu + sd * sd - torch.log(sd * sd)).sum(), (dec_ref - y).abs().sum()
    (ll_scale * l1_ref + kl_scale * kl_ref).backward()
    grads = [torch.zeros_like(p_) for p_ in params]
    wp, gp = (C.c_void_p * 10)(), (C.c_void_p * 10)()
    for i in range(10):
        wp[i], gp[i] = params[i].data_ptr(), grads[i].data_ptr()
    dec, acc = torch.empty(rows, d, device="cuda"), torch.zeros(2, device="cuda")
    ctx.vae_step(y.data_ptr(), noise.data_ptr(), wp, gp, dec.data_ptr(), acc.data_ptr(), rows, d, h, z, ll_scale, kl_scale,
                 SLOPE)
    assert rel_l2(dec, dec_ref.detach()) < 1e-5
    assert abs(acc[0].item() - kl_ref.item()) < 1e-4 * abs(kl_ref.item())
    assert abs(acc[1].item() - l1_ref.item()) < 1e-4 * abs(l1_ref.item())
    for p_, g_ in zip(params, grads):
        assert rel_l2(g_, p_.grad) < 2e-4


@pytest.mark.parametrize("n,h,cin", [(5, 64, 128), (3, 32, 64), (37, 64, 128)])
def test_decoder_head_fused_into_the_upsampling_epilogue(ctx, n, h, cin):
    """lsps_conv_ext.head_*: ConvTranspose2d(cin,64,3,2,1,1) + LeakyReLU -> ConvTranspose2d(64,1,1) + Tanh (+ L1 term) in
    one launch == the transposed conv followed by lsps_head_fwd_l1 on its stored output."""
    from lsps_b200._lib import ConvShape, ConvExt
    g = gen(1000 + n)
    x = torch.randn(n, h, h, cin, device="cuda", generator=g).bfloat16()
    wf = (torch.randn(9, 64, cin, device="cuda", generator=g) * 0.05).bfloat16()
    bias = torch.randn(64, device="cuda", generator=g) * 0.1
    hw, hb = torch.randn(64, device="cuda", generator=g) * 0.2, torch.randn(1, device="cuda", generator=g) * 0.1
    sh = C.byref(ConvShape(2, n, h, h, cin, 64))
    px = 4 * h * h
    first, nt = 1, n - 2                                   # L1 over images [1, n-1)
    target = torch.rand(nt, 2 * h, 2 * h, device="cuda", generator=g) * 2 - 1
    # reference: two launches
    y0 = torch.empty(n, 2 * h, 2 * h, 64, device="cuda", dtype=torch.bfloat16)
    ctx.conv_fwd(sh, x.data_ptr(), wf.data_ptr(), bias.data_ptr(), y0.data_ptr(), 3, SLOPE)
    img0, dout0, acc0 = torch.empty(n, 2 * h, 2 * h, device="cuda"), torch.zeros_like(target), torch.zeros(1, device="cuda")
    ctx.head_fwd_l1(y0.data_ptr(), hw.data_ptr(), hb.data_ptr(), img0.data_ptr(), img0.numel(), target.data_ptr(),
                    first * px, nt * px, 0.25, dout0.data_ptr(), acc0.data_ptr())
    # fused
    y1 = torch.empty_like(y0)
    img1, dout1, acc1 = torch.empty_like(img0), torch.zeros_like(target), torch.zeros(1, device="cuda")
    ext = ConvExt()
    ext.head_w, ext.head_b, ext.head_out = hw.data_ptr(), hb.data_ptr(), img1.data_ptr()
    ext.head_target, ext.head_t0, ext.head_tn = target.data_ptr(), first * px, nt * px
    ext.head_scale, ext.head_dout, ext.head_acc = 0.25, dout1.data_ptr(), acc1.data_ptr()
    ctx.conv_fwd_ex(sh, x.data_ptr(), wf.data_ptr(), bias.data_ptr(), y1.data_ptr(), 3, SLOPE, C.byref(ext))
    assert torch.equal(y0, y1)
    assert (img0 - img1).abs().max().item() < 2e-6
    assert abs(acc0.item() - acc1.item()) < 1e-5 * abs(acc0.item())
    close = (img0[first:first + nt] - target).abs() > 1e-5          # sign ties apart, the gradients are identical
    assert torch.equal(dout0[close], dout1[close])
    # a shape the fused kernel does not take must refuse the head instead of dropping it
    sh2 = C.byref(ConvShape(2, 2, 32, 32, 256, 128))
    x2 = torch.randn(2, 32, 32, 256, device="cuda", generator=g).bfloat16()
    w2 = (torch.randn(9, 128, 256, device="cuda", generator=g) * 0.05).bfloat16()
    yb = torch.empty(2, 64, 64, 128, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(Exception):
        ctx.conv_fwd_ex(sh2, x2.data_ptr(), w2.data_ptr(), torch.zeros(128, device="cuda").data_ptr(), yb.data_ptr(), 3, SLOPE,
                        C.byref(ext))
