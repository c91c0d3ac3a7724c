"""the normalised conv wrappers of the reference's layer zoo (BatchNorm / InstanceNorm variants,
common_net.py:137-158,183-199,270-379) on the device kernels, against the same stacks built from torch.nn in fp32
(train-mode batch statistics, running statistics, affine / Bias2d parameters, eval mode).  The torch.nn stacks
themselves are pinned to the reference's own classes by tests/test_layers_ref_cpu.py."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu
SLOPE = 0.01


def rel_l2(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-20)).item()


def nhwc16(t):
    return t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def nchw32(t):
    return t.float().permute(0, 3, 1, 2)


class Bias2d(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(c))

    def forward(self, x):
        return x + self.bias[None, :, None, None]


def _ref_stack(norm, cin, cout, k, stride, transposed, slope, device="cuda"):
    conv = (nn.ConvTranspose2d(cin, cout, k, stride, 1, output_padding=1, bias=norm != "bn") if transposed else
            nn.Conv2d(cin, cout, k, stride, k // 2, bias=norm != "bn"))
    layers = [conv]
    if norm == "bn":
        layers.append(nn.BatchNorm2d(cout))
    elif norm == "bnns":
        layers += [nn.BatchNorm2d(cout, affine=False), Bias2d(cout)]
    else:
        layers.append(nn.InstanceNorm2d(cout, affine=False))
    layers.append(nn.LeakyReLU(slope) if slope > 0 else nn.ReLU())
    return nn.Sequential(*layers).to(device)


@pytest.mark.parametrize("norm,k,stride,transposed,cin,cout,slope", [
    ("bn", 3, 1, False, 64, 128, SLOPE), ("bn", 3, 2, False, 128, 256, SLOPE), ("bn", 3, 2, True, 256, 128, SLOPE),
    ("bnns", 3, 2, False, 64, 128, SLOPE), ("bnns", 3, 2, True, 128, 64, SLOPE), ("bn", 1, 1, False, 256, 256, SLOPE),
    ("ins", 3, 2, False, 64, 128, SLOPE), ("ins", 3, 2, True, 128, 64, SLOPE), ("ins", 3, 1, False, 128, 128, 0.0)])
def test_conv_norm_act_matches_torch(norm, k, stride, transposed, cin, cout, slope):
    import lsps_b200
    from lsps_b200 import layers
    from lsps_b200.engine import Ops
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ops = Ops("cuda:0")
    g = torch.Generator(device="cuda").manual_seed(7)
    n, h = 6, 32
    L = layers.ConvNormAct(ops, cin, cout, k, stride, transposed=transposed, norm=norm, slope=slope, seed=3)
    ref = _ref_stack(norm, cin, cout, k, stride, transposed, slope)
    sd = L.state_dict()
    if norm == "bn":       # non-trivial affine parameters
        sd["model.1.weight"] = torch.rand(cout, generator=torch.Generator().manual_seed(1)) + 0.5
        sd["model.1.bias"] = torch.randn(cout, generator=torch.Generator().manual_seed(2)) * 0.3
    if norm == "bnns":
        sd["model.2.bias"] = torch.randn(cout, generator=torch.Generator().manual_seed(2)) * 0.3
    sd["model.0.weight"] = sd["model.0.weight"].bfloat16().float()          # the kernels read bf16 operands
    L.load_state_dict(sd)
    with torch.no_grad():
        ref[0].weight.copy_(sd["model.0.weight"])
        if "model.0.bias" in sd:
            ref[0].bias.copy_(sd["model.0.bias"])
        if norm == "bn":
            ref[1].weight.copy_(sd["model.1.weight"]); ref[1].bias.copy_(sd["model.1.bias"])
        if norm == "bnns":
            ref[2].bias.copy_(sd["model.2.bias"])
    x = torch.randn(n, cin, h, h, device="cuda", generator=g).bfloat16().float().requires_grad_(True)
    y_ref = ref(x)
    dy = torch.randn(y_ref.shape, device="cuda", generator=g).bfloat16().float()
    y_ref.backward(dy)
    y = L.forward(nhwc16(x.detach()))
    assert rel_l2(nchw32(y), y_ref.detach()) < 8e-3
    L.S.zero_grad()
    dx = L.backward(nhwc16(dy))
    assert rel_l2(nchw32(dx), x.grad) < 2.5e-2          # bf16 operands + LeakyReLU-mask flips near zero (see below)
    from lsps_b200.params import from_kernel_layout
    e = L.S.entries["model.0.weight"]
    dw = from_kernel_layout(e.kind, L.S.G("model.0.weight"), e.shape)
    assert rel_l2(dw, ref[0].weight.grad) < 2.5e-2
    # d gamma / d beta are sums of dy * lrelu'(pre-activation): the bf16-stored conv output flips the mask of the few
    # pre-activations within rounding distance of zero, each flip moving one channel's sum by ~1 % of a term
    if norm == "bn":
        assert rel_l2(L.S.G("model.1.weight"), ref[1].weight.grad) < 3e-2 and rel_l2(L.S.G("model.1.bias"), ref[1].bias.grad) < 3e-2
    if norm == "bnns":
        assert rel_l2(L.S.G("model.2.bias"), ref[2].bias.grad) < 3e-2
    if norm != "ins":
        assert rel_l2(L.running_mean, ref[1].running_mean) < 1e-2 and rel_l2(L.running_var, ref[1].running_var) < 1e-2
        L.eval(); ref.eval()
        with torch.no_grad():
            y_ref_e = ref(x)
        assert rel_l2(nchw32(L.forward(nhwc16(x.detach()))), y_ref_e) < 8e-3


@pytest.mark.parametrize("which", ["bn", "ins"])
def test_norm_res_blocks_match_torch(which):
    """LeakyReLUBNNSResBlock (BatchNorm2d(affine=False), bias-free convs) and INSResBlock (InstanceNorm2d, ReLU)."""
    from lsps_b200 import layers
    from lsps_b200.engine import Ops
    torch.backends.cudnn.allow_tf32 = False
    ops = Ops("cuda:0")
    c, n, h = 128, 4, 32
    blk = layers.LeakyReLUBNNSResBlock(ops, c, c, seed=5) if which == "bn" else layers.INSResBlock(ops, c, c, seed=5)
    bn = which == "bn"
    norm = (lambda: nn.BatchNorm2d(c, affine=False)) if bn else (lambda: nn.InstanceNorm2d(c))
    ref = nn.Sequential(nn.Conv2d(c, c, 3, 1, 1, bias=not bn), norm(), nn.LeakyReLU(SLOPE) if bn else nn.ReLU(),
                        nn.Conv2d(c, c, 3, 1, 1, bias=not bn), norm()).cuda()
    for l, idx in ((blk.a, 0), (blk.b, 3)):
        sd = l.state_dict()
        sd["model.0.weight"] = sd["model.0.weight"].bfloat16().float()
        l.load_state_dict(sd)
        if bn:
            l.S.W("model.2.bias").zero_()
        with torch.no_grad():
            ref[idx].weight.copy_(sd["model.0.weight"])
            if not bn:
                ref[idx].bias.copy_(sd["model.0.bias"])
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(n, c, h, h, device="cuda", generator=g).bfloat16().float().requires_grad_(True)
    y_ref = x + ref(x)
    dy = torch.randn(y_ref.shape, device="cuda", generator=g).bfloat16().float()
    y_ref.backward(dy)
    y = blk.forward(nhwc16(x.detach()))
    assert rel_l2(nchw32(y), y_ref.detach()) < 8e-3
    dx = blk.backward(nhwc16(dy))
    assert rel_l2(nchw32(dx), x.grad) < 1.5e-2
