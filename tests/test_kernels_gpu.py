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


@pytest.mark.parametrize("stride,n", [(1, 3), (2, 5)])
def test_stem(ctx, stride, n):
    g = gen(stride)
    img = (torch.rand(n, 1, 128, 128, device="cuda", generator=g) * 2 - 1).requires_grad_(True)
    w = (torch.randn(64, 1, 7, 7, device="cuda", generator=g) * 0.05).requires_grad_(True)
    b = (torch.randn(64, device="cuda", generator=g) * 0.1).requires_grad_(True)
    pre = F.conv2d(img, w, b, stride=stride, padding=3)
    y = F.leaky_relu(pre, SLOPE)
    ho = 128 // stride
    yb = torch.empty(n, ho, ho, 64, device="cuda", dtype=torch.bfloat16)
    ctx.stem_fwd(img.data_ptr(), w.data_ptr(), b.data_ptr(), yb.data_ptr(), n, 128, 128, stride, SLOPE)
    assert rel_l2(nchw32(yb), y) < BF16_L2
    dpre = torch.randn(pre.shape, device="cuda", generator=g).bfloat16().float()
    pre.backward(dpre)
    dw, db = torch.zeros(64, 49, device="cuda"), torch.zeros(64, device="cuda")
    dyb = nhwc16(dpre)
    ctx.stem_wgrad(img.data_ptr(), dyb.data_ptr(), dw.data_ptr(), db.data_ptr(), n, 128, 128, stride)
    assert rel_l2(dw, w.grad.reshape(64, 49)) < 1e-4 and rel_l2(db, b.grad) < 1e-4
    dimg = torch.full((n, 128, 128), 0.5, device="cuda")
    # tensor-core dgrad: bf16 weights (like every other conv operand), fp32 accumulation -> weight-rounding noise 2^-9
    ctx.stem_dgrad(dyb.data_ptr(), w.data_ptr(), dimg.data_ptr(), n, 128, 128, stride, 0)
    assert rel_l2(dimg, img.grad[:, 0]) < BF16_L2
    ctx.stem_dgrad(dyb.data_ptr(), w.data_ptr(), dimg.data_ptr(), n, 128, 128, stride, 1)
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
