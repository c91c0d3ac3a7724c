"""Pins the torch.nn restatements that tests/test_layers_gpu.py compares the device kernels with to the reference's OWN
layer-zoo classes (src/trainers/common_net.py:137-158,183-199,270-379): same weights, same input -> identical forward
values and gradients.  Runs where the reference is mounted (/root/reference) or staged (oracle/_ref)."""
import os
import sys

import pytest
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_loader  # noqa: E402

pytestmark = pytest.mark.skipif(not (ref_loader.reference_available() or ref_loader.staged_available()),
                                reason="reference sources neither mounted nor staged")

CASES = [  # (test_layers_gpu norm tag, k, stride, transposed, slope, reference class, ctor args after (cin, cout))
    ("bn", 3, 1, False, 0.01, "LeakyReLUBNConv2d", (3, 1, 1)),
    ("bn", 3, 2, False, 0.01, "LeakyReLUBNConv2d", (3, 2, 1)),
    ("bn", 1, 1, False, 0.01, "LeakyReLUBNConv2d", (1, 1, 0)),
    ("bn", 3, 2, True, 0.01, "LeakyReLUBNConvTranspose2d", (3, 2, 1, 1)),
    ("bnns", 3, 2, False, 0.01, "LeakyReLUBNNSConv2d", (3, 2, 1)),
    ("ins", 3, 2, False, 0.01, "LeakyReLUINSConv2d", (3, 2, 1)),
    ("ins", 3, 2, True, 0.01, "LeakyReLUINSConvTranspose2d", (3, 2, 1, 1)),
    ("ins", 3, 1, False, 0.0, "ReLUINSConv2d", (3, 1, 1)),
]


def _common_net():
    ref_loader.load_reference()
    import trainers.common_net as cn
    return cn


def _grads(mod, x, dy):
    x = x.clone().requires_grad_(True)
    y = mod(x)
    y.backward(dy)
    return y.detach(), x.grad, [p.grad.clone() for p in mod.parameters()]


@pytest.mark.parametrize("norm,k,stride,transposed,slope,cls,args", CASES)
def test_nn_stack_equals_reference_class(norm, k, stride, transposed, slope, cls, args):
    import test_layers_gpu as T
    cn = _common_net()
    torch.manual_seed(3)
    cin, cout, n, h = 8, 12, 3, 8
    ref = getattr(cn, cls)(cin, cout, *args)
    mine = T._ref_stack(norm, cin, cout, k, stride, transposed, slope, device="cpu")
    mine.load_state_dict(ref.model.state_dict())           # same Sequential layout: conv, norm, [Bias2d], activation
    if norm == "bn":                                       # non-trivial affine parameters on both sides
        with torch.no_grad():
            for m in (ref.model[1], mine[1]):
                m.weight.copy_(torch.linspace(0.5, 1.5, cout)); m.bias.copy_(torch.linspace(-0.3, 0.3, cout))
    x = torch.randn(n, cin, h, h)
    y_ref = ref(x.clone())
    mine(x.clone())                                        # both sides see the same two train-mode passes
    dy = torch.randn(y_ref.shape)
    for m in (ref, mine):
        m.zero_grad()
    yr, gxr, gpr = _grads(ref, x, dy)
    ym, gxm, gpm = _grads(mine, x, dy)
    assert torch.equal(yr, ym) and torch.equal(gxr, gxm)
    assert len(gpr) == len(gpm) and all(torch.equal(a, b) for a, b in zip(gpr, gpm))
    if norm != "ins":                                      # running statistics after the same two train-mode passes
        assert torch.equal(ref.model[1].running_mean, mine[1].running_mean)
        assert torch.equal(ref.model[1].running_var, mine[1].running_var)


@pytest.mark.parametrize("which", ["bn", "ins"])
def test_res_block_restatement_equals_reference_class(which):
    """x + Sequential(conv, norm, act, conv, norm) as written in test_layers_gpu.test_norm_res_blocks_match_torch."""
    cn = _common_net()
    torch.manual_seed(5)
    c, n, h = 8, 3, 8
    bn = which == "bn"
    ref = cn.LeakyReLUBNNSResBlock(c, c, 3, 1, 1) if bn else cn.INSResBlock(c, c)
    norm = (lambda: nn.BatchNorm2d(c, affine=False)) if bn else (lambda: nn.InstanceNorm2d(c))
    mine = nn.Sequential(nn.Conv2d(c, c, 3, 1, 1, bias=not bn), norm(), nn.LeakyReLU(0.01) if bn else nn.ReLU(),
                         nn.Conv2d(c, c, 3, 1, 1, bias=not bn), norm())
    mine.load_state_dict(ref.model.state_dict())
    x = torch.randn(n, c, h, h)
    dy = torch.randn(n, c, h, h)
    xr, xm = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    yr = ref(xr * 1.0)                                     # (the reference adds the residual in place)
    ym = xm + mine(xm)
    yr.backward(dy); ym.backward(dy)
    assert torch.equal(yr.detach(), ym.detach()) and torch.equal(xr.grad, xm.grad)
    assert all(torch.equal(a.grad, b.grad) for a, b in zip(ref.model.parameters(), mine.parameters()))
