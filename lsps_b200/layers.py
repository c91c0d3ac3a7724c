"""Device versions of the reference's normalised conv wrappers (the part of the layer zoo no shipped config
instantiates): /root/reference/src/trainers/common_net.py

    LeakyReLUBNConv2d :270-281      LeakyReLUBNConvTranspose2d :283-294     conv(bias=False) -> BatchNorm2d(affine) -> lrelu
    LeakyReLUBNNSConv2d :296-308    LeakyReLUBNNSConvTranspose2d :310-322   conv(bias) -> BatchNorm2d(affine=False) -> Bias2d -> lrelu
    LeakyReLUINSConv2d :324-335     LeakyReLUINSConvTranspose2d :337-349    conv(bias) -> InstanceNorm2d -> lrelu
    ReLUINSConv2d :354-365          ReLUINSConvTranspose2d :367-379         ... -> ReLU
    LeakyReLUBNNSResBlock :183-199  INSResBlock :137-158                    two of the above around a residual add

Every one is the same three kernels: the tcgen05 implicit-GEMM conv with the statistics epilogue (LSPS_EP_STATS), one
streaming normalise(+affine / Bias2d)(+activation / +residual) pass, and in the backward pass one reduction + one apply
pass around the conv data / weight gradients.  BatchNorm's batch statistics are the per-image rows of the epilogue's sums
added up (`lsps_norm_reduce_images`); under data parallelism that [2, c] row (and the backward one) is what the ranks
all-reduce -- train-mode BatchNorm is the one layer of the zoo that is NOT per-sample.

Kernel shapes covered: Conv2d 3x3 stride 1 / 2 pad 1, 1x1, ConvTranspose2d 3x3 stride 2 pad 1 output_padding 1; channels
multiples of 64, spatial sizes powers of two.  Activations are bf16 NHWC; parameters live in a flat fp32 ParamStore with
the reference's state_dict keys (model.0.weight, model.1.weight/bias/running_mean/running_var, model.2.bias).
"""
import ctypes as C

import torch

from . import _lib
from ._lib import CONV_S1, CONV_S2, DECONV_S2, CONV1X1, EP_BIAS, EP_STATS, ConvShape, ConvExt
from .params import ParamStore
from .sharding import allreduce_sum_, world_rank

BN_EPS, IN_EPS, BN_MOMENTUM = 1e-5, 1e-5, 0.1


def _kind(kernel_size, stride, transposed):
    if transposed:
        if kernel_size == 3 and stride == 2:
            return DECONV_S2
    elif kernel_size == 3 and stride in (1, 2):
        return CONV_S1 if stride == 1 else CONV_S2
    elif kernel_size == 1 and stride == 1:
        return CONV1X1
    raise NotImplementedError("conv kernels cover 3x3 s1/s2 p1, 1x1 and transposed 3x3 s2 p1 op1 (got k=%d s=%d%s)"
                              % (kernel_size, stride, " transposed" if transposed else ""))


class ConvNormAct:
    """conv -> norm ('bn' = BatchNorm2d affine, 'bnns' = BatchNorm2d(affine=False) + Bias2d, 'ins' = InstanceNorm2d)
    -> LeakyReLU(slope) (slope 0 = ReLU; act=False: no activation; res: y = res + norm(...))."""

    def __init__(self, ops, n_in, n_out, kernel_size, stride, transposed=False, norm="bn", conv_bias=None, slope=0.01,
                 act=True, seed=0, lr=1e-4, weight_decay=1e-4):
        self.ops, self.ctx = ops, ops.ctx
        self.kind = _kind(kernel_size, stride, transposed)
        self.cin, self.cout, self.norm, self.slope, self.act = n_in, n_out, norm, float(slope), act
        self.has_bias = (norm != "bn") if conv_bias is None else conv_bias
        wkind = {CONV_S1: "conv3", CONV_S2: "conv3", DECONV_S2: "deconv3", CONV1X1: "conv1"}[self.kind]
        wshape = (n_in, n_out, 3, 3) if transposed else (n_out, n_in, kernel_size, kernel_size)
        fan = (n_out if transposed else n_in) * kernel_size * kernel_size
        ents = [("model.0.weight", wshape, wkind, "conv", fan)]
        if self.has_bias:
            ents.append(("model.0.bias", (n_out,), "bias", "bias", fan))
        if norm == "bn":
            ents += [("model.1.weight", (n_out,), "bias", "bias", 1), ("model.1.bias", (n_out,), "bias", "bias", 1)]
        if norm == "bnns":
            ents.append(("model.2.bias", (n_out,), "bias", "small", 1))
        self.S = ParamStore(ents, ops.device, lr, weight_decay)
        self.S.init_(seed)
        if norm == "bn":          # nn.BatchNorm2d default init (gaussian_weights_init only touches Conv* classes)
            sd = self.S.state_dict()
            sd["model.1.weight"], sd["model.1.bias"] = torch.ones(n_out), torch.zeros(n_out)
            self.S.load_state_dict(sd)
        self.is_bn = norm in ("bn", "bnns")
        if self.is_bn:
            self.running_mean = torch.zeros(n_out, device=ops.device)
            self.running_var = torch.ones(n_out, device=ops.device)
            self.num_batches_tracked = 0
        self.training = True
        self._saved = None

    # -- reference-compatible parameters / buffers
    def state_dict(self):
        sd = self.S.state_dict()
        if self.is_bn:
            sd["model.1.running_mean"], sd["model.1.running_var"] = self.running_mean.clone(), self.running_var.clone()
            sd["model.1.num_batches_tracked"] = torch.tensor(self.num_batches_tracked)
        return sd

    def load_state_dict(self, sd):
        self.S.load_state_dict({k: v for k, v in sd.items() if k in self.S.entries})
        if self.is_bn and "model.1.running_mean" in sd:
            self.running_mean.copy_(sd["model.1.running_mean"])
            self.running_var.copy_(sd["model.1.running_var"])
            self.num_batches_tracked = int(sd.get("model.1.num_batches_tracked", 0))

    def train(self, mode=True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    def _gamma_beta(self):
        S = self.S
        if self.norm == "bn":
            return S.W("model.1.weight"), S.W("model.1.bias"), S.G("model.1.weight"), S.G("model.1.bias")
        if self.norm == "bnns":
            return None, S.W("model.2.bias"), None, S.G("model.2.bias")
        return None, None, None, None

    def _out_hw(self, h, w):
        return (h, w) if self.kind in (CONV_S1, CONV1X1) else ((h // 2, w // 2) if self.kind == CONV_S2 else (2 * h, 2 * w))

    def forward(self, x, res=None):
        """x bf16 [n,h,w,cin] -> bf16 [n,ho,wo,cout]; res (optional, bf16, output-shaped): y = res + norm(conv(x))."""
        ctx, S, o = self.ctx, self.S, self.ops
        n, h, w, cin = x.shape
        assert cin == self.cin
        ho, wo = self._out_hw(h, w)
        hw, c = ho * wo, self.cout
        hconv = o.empty(n, ho, wo, c)
        sums = o.empty(n, 2, c, dtype=torch.float32)
        ext = ConvExt()
        ext.sums = sums.data_ptr()
        ctx.conv_fwd_ex(C.byref(ConvShape(self.kind, n, h, w, self.cin, c)), x.data_ptr(), S.W16("model.0.weight").data_ptr(),
                        S.W("model.0.bias").data_ptr() if self.has_bias else None, hconv.data_ptr(),
                        (EP_BIAS if self.has_bias else 0) | EP_STATS, self.slope, C.byref(ext))
        gamma, beta, _, _ = self._gamma_beta()
        mode = 1 if res is not None else (0 if self.act else 2)
        y = o.empty(n, ho, wo, c)
        if self.is_bn:
            world, _ = world_rank()
            row = o.empty(2, c, dtype=torch.float32)
            count = float(n * hw * world)
            if self.training:
                ctx.norm_reduce_images(sums.data_ptr(), row.data_ptr(), n, c)
                allreduce_sum_(row)                      # batch statistics are over the GLOBAL batch
                ctx.bn_running_update(row.data_ptr(), self.running_mean.data_ptr(), self.running_var.data_ptr(), c, count,
                                      BN_MOMENTUM)
                self.num_batches_tracked += 1
            else:
                ctx.bn_running_to_sums(self.running_mean.data_ptr(), self.running_var.data_ptr(), row.data_ptr(), c, count)
            stats = o.empty(2, c, dtype=torch.float32)
            # per_image = 0: the kernel divides by n*hw; with several ranks the row already holds the global sums
            ctx.norm_apply_fwd(hconv.data_ptr(), _lib.ptr(res), y.data_ptr(), self._scaled(row, world).data_ptr(),
                               stats.data_ptr(), n, hw, c, mode, 0, BN_EPS, self.slope, _lib.ptr(gamma), _lib.ptr(beta))
        else:
            stats = o.empty(n, 2, c, dtype=torch.float32)
            ctx.norm_apply_fwd(hconv.data_ptr(), _lib.ptr(res), y.data_ptr(), sums.data_ptr(), stats.data_ptr(), n, hw, c,
                               mode, 1, IN_EPS, self.slope, _lib.ptr(gamma), _lib.ptr(beta))
        self._saved = (x, hconv, stats, mode)
        return y

    @staticmethod
    def _scaled(row, world):
        return row if world == 1 else row / world     # local count n*hw times world = global count

    def backward(self, dy, need_dx=True):
        """dy: gradient w.r.t. forward()'s output (for res != None also the gradient of the residual branch, which the
        caller adds).  Accumulates parameter gradients into the store; returns dx (bf16) or None."""
        ctx, S, o = self.ctx, self.S, self.ops
        x, hconv, stats, mode = self._saved
        n, ho, wo, c = hconv.shape
        hw = ho * wo
        gamma, beta, dgamma, dbeta = self._gamma_beta()
        per_image = 0 if self.is_bn else 1
        bs = o.empty(*((2, c) if self.is_bn else (n, 2, c)), dtype=torch.float32)
        smode = 0 if mode == 0 else 1                    # lrelu mask only when an activation followed
        ctx.norm_bwd_stats(dy.data_ptr(), hconv.data_ptr(), stats.data_ptr(), bs.data_ptr(), n, hw, c, smode, per_image,
                           self.slope, _lib.ptr(gamma), _lib.ptr(beta))
        world, _ = world_rank()
        if self.is_bn:
            allreduce_sum_(bs)
            if dbeta is not None:                        # sum g / sum g*xhat ARE d beta / d gamma (already global)
                dbeta += bs[0] / world                   # the gradient allreduce of the store sums over ranks again
            if dgamma is not None:
                dgamma += bs[1] / world
        dh = torch.empty_like(hconv)
        ctx.norm_bwd_apply(dy.data_ptr(), hconv.data_ptr(), stats.data_ptr(), self._scaled(bs, world).data_ptr() if self.is_bn
                           else bs.data_ptr(), dh.data_ptr(), n, hw, c, smode, per_image, self.slope, _lib.ptr(gamma),
                           _lib.ptr(beta))
        sh = C.byref(ConvShape(self.kind, x.shape[0], x.shape[1], x.shape[2], self.cin, c))
        ctx.conv_wgrad(sh, x.data_ptr(), dh.data_ptr(), S.G("model.0.weight").data_ptr())
        # a conv bias in front of a norm has an exactly-zero gradient (the norm removes the mean): left at zero
        if not need_dx:
            return None
        dx = torch.empty_like(x)
        ctx.conv_dgrad(sh, dh.data_ptr(), S.W16T("model.0.weight").data_ptr(), dx.data_ptr(), None, None, 0, self.slope)
        return dx


# ---- the reference's names (same constructor arguments; padding / output_padding are implied by the supported shapes)
def LeakyReLUBNConv2d(ops, n_in, n_out, kernel_size, stride, padding=0, **kw):
    return ConvNormAct(ops, n_in, n_out, kernel_size, stride, norm="bn", **kw)


def LeakyReLUBNConvTranspose2d(ops, n_in, n_out, kernel_size, stride, padding=0, output_padding=0, **kw):
    return ConvNormAct(ops, n_in, n_out, kernel_size, stride, transposed=True, norm="bn", **kw)


def LeakyReLUBNNSConv2d(ops, n_in, n_out, kernel_size, stride, padding=0, **kw):
    return ConvNormAct(ops, n_in, n_out, kernel_size, stride, norm="bnns", **kw)


def LeakyReLUBNNSConvTranspose2d(ops, n_in, n_out, kernel_size, stride, padding=0, **kw):
    return ConvNormAct(ops, n_in, n_out, kernel_size, stride, transposed=True, norm="bnns", **kw)


def LeakyReLUINSConv2d(ops, n_in, n_out, kernel_size, stride, padding=0, **kw):
    return ConvNormAct(ops, n_in, n_out, kernel_size, stride, norm="ins", **kw)


def LeakyReLUINSConvTranspose2d(ops, n_in, n_out, kernel_size, stride, padding=0, output_padding=0, **kw):
    return ConvNormAct(ops, n_in, n_out, kernel_size, stride, transposed=True, norm="ins", **kw)


def ReLUINSConv2d(ops, n_in, n_out, kernel_size, stride, padding=0, **kw):
    return ConvNormAct(ops, n_in, n_out, kernel_size, stride, norm="ins", slope=0.0, **kw)


def ReLUINSConvTranspose2d(ops, n_in, n_out, kernel_size, stride, padding=0, output_padding=0, **kw):
    return ConvNormAct(ops, n_in, n_out, kernel_size, stride, transposed=True, norm="ins", slope=0.0, **kw)


class NormResBlock:
    """x + norm(conv(act(norm(conv(x))))): LeakyReLUBNNSResBlock (common_net.py:183-199: BatchNorm2d(affine=False),
    conv bias=False, LeakyReLU) and INSResBlock (:137-158: InstanceNorm2d, conv bias, ReLU).  state_dict keys model.0.* /
    model.3.* (+ model.1 / model.4 running statistics for BatchNorm)."""

    def __init__(self, ops, n_ch, norm="bn_plain", slope=0.01, seed=0):
        bn = norm == "bn_plain"
        kw = dict(norm="bnns" if bn else "ins", conv_bias=not bn, slope=slope)
        self.a = ConvNormAct(ops, n_ch, n_ch, 3, 1, seed=seed, **kw)
        self.b = ConvNormAct(ops, n_ch, n_ch, 3, 1, seed=seed + 1, act=False, **kw)
        self.bn = bn
        if bn:      # BatchNorm2d(affine=False) WITHOUT a Bias2d: keep the kernels' beta at zero and never step it
            for l in (self.a, self.b):
                l.S.W("model.2.bias").zero_()

    def forward(self, x):
        return self.b.forward(self.a.forward(x), res=x)

    def backward(self, dy):
        dx = self.a.backward(self.b.backward(dy))
        self.ops_add(dx, dy)
        return dx

    def ops_add(self, dx, dy):
        self.a.ctx.axpy_bf16(dx.data_ptr(), dy.data_ptr(), 1.0, dx.data_ptr(), dx.numel())

    def train(self, mode=True):
        self.a.train(mode)
        self.b.train(mode)
        return self

    def state_dict(self):
        out = {}
        for pre, l in (("model.0", self.a), ("model.3", self.b)):
            for k, v in l.state_dict().items():
                if k.startswith("model.0."):
                    out[pre + k[7:]] = v
                elif k.startswith("model.1."):
                    out["model.%d" % (int(pre[-1]) + 1) + k[7:]] = v
        return out


def LeakyReLUBNNSResBlock(ops, n_in, n_out, kernel_size=3, stride=1, padding=1, **kw):
    assert n_in == n_out and kernel_size == 3 and stride == 1
    return NormResBlock(ops, n_in, norm="bn_plain", **kw)


def INSResBlock(ops, inplanes, planes, stride=1, dropout=0.0, **kw):
    assert inplanes == planes and stride == 1 and dropout == 0.0
    return NormResBlock(ops, inplanes, norm="ins", slope=0.0, **kw)
