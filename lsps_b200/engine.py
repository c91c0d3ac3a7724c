"""Hand-scheduled forward/backward of the LSPS networks on the lsps_b200 kernels.

There is no autograd: every update of the reference trainer (/root/reference/src/trainers/lsps_trainer.py:62-262)
is a static schedule over the C-ABI kernels -- forward passes keep exactly the activations
their backward needs, gradients that the reference computes and then discards (generator backward inside
dis_update, discriminator wgrad inside gen_update) are never computed.

Layouts: images fp32 [n,128,128]; activations bf16 NHWC; InstanceNorm statistics, loss sums and all parameter
gradients fp32.  Network structure follows lsps_nets.py:86-160 (SharedDis) and :164-272 (SharedResGen).
"""
import ctypes as C
import os

import torch

from . import _lib
from ._lib import (CONV_S1, CONV_S2, DECONV_S2, DECONV4_S2, CONV1X1, EP_BIAS, EP_LRELU, EP_MASK, EP_ADD, EP_STATS, EP_INBWD,
                   ConvShape, ConvExt)

SLOPE = 0.01   # nn.LeakyReLU() default (common_net.py:169,251)
IN_EPS = 1e-5  # nn.InstanceNorm2d default


def _shape(kind, n, h, w, cin, cout):
    return C.byref(ConvShape(kind, n, h, w, cin, cout))


class Ops:
    """Typed wrappers: torch tensors in, kernels out."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.ctx = _lib.context(self.device.index)
        # weight-gradient kernels run on a side stream: they only feed the optimiser, so the tensor-bound wgrad GEMMs
        # overlap the HBM-bound InstanceNorm-backward / small kernels of the data-gradient chain (join_side() before Adam)
        self.side = torch.cuda.Stream(device=self.device)
        self.use_side = os.environ.get("LSPS_NO_SIDE", "0") != "1"
        self._side_refs = []
        # InstanceNorm statistics taken in the conv epilogues + one streaming apply pass (round 2); LSPS_OLD_IN=1 keeps
        # the stand-alone three-pass kernels for A/B runs
        self.fused_in = os.environ.get("LSPS_OLD_IN", "0") != "1"

    def _on_side(self, fn, *keep):
        if not self.use_side or torch.cuda.is_current_stream_capturing():
            fn()
            return
        ev = torch.cuda.Event()
        ev.record()                      # everything the kernel reads has been enqueued on the main stream
        self.side.wait_event(ev)
        with torch.cuda.stream(self.side):
            fn()
        self._side_refs.append(keep)     # keep the operands alive (and their memory unrecycled) until the join

    def side_event(self):
        """Event that completes when everything enqueued so far on the weight-gradient stream has run."""
        ev = torch.cuda.Event()
        ev.record(self.side if (self.use_side and self._side_refs) else torch.cuda.current_stream())
        return ev

    def join_side(self):
        if self._side_refs:
            ev = torch.cuda.Event()
            ev.record(self.side)
            torch.cuda.current_stream().wait_event(ev)
            self._side_refs = []

    def empty(self, *shape, dtype=torch.bfloat16):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def zeros(self, *shape, dtype=torch.float32):
        """zero-filled tensor through a stream-ordered memset (no library fill kernel on the hot path)"""
        t = torch.empty(shape, dtype=dtype, device=self.device)
        self.ctx.memset(t.data_ptr(), 0, t.numel() * t.element_size())
        return t

    def cat(self, tensors):
        """torch.cat along dim 0 of contiguous same-shaped-tail tensors as plain device-to-device copies"""
        tensors = [t for t in tensors if t is not None]
        if len(tensors) == 1:
            return tensors[0]
        out = torch.empty((sum(t.shape[0] for t in tensors),) + tuple(tensors[0].shape[1:]), dtype=tensors[0].dtype,
                          device=self.device)
        r = 0
        for t in tensors:
            self.copy_into(out[r:r + t.shape[0]], t)
            r += t.shape[0]
        return out

    def copy_into(self, dst, src):
        assert dst.is_contiguous() and src.is_contiguous() and dst.numel() == src.numel() and dst.dtype == src.dtype
        self.ctx.memcpy(dst.data_ptr(), src.data_ptr(), src.numel() * src.element_size())

    def noise_kl(self, x, noise, z, acc):
        """GaussianNoiseLayer + KL sum.  noise: host-drawn fp32 tensor shaped like x (parity mode) or a
        ("philox", seed, offset) token (device mode: drawn inside the kernel)."""
        if isinstance(noise, tuple):
            self.ctx.noise_kl_philox(x.data_ptr(), z.data_ptr(), acc.data_ptr(), x.numel(), noise[1], noise[2])
        else:
            self.ctx.noise_kl_fwd(x.data_ptr(), noise.data_ptr(), z.data_ptr(), acc.data_ptr(), x.numel())

    # ---- 3x3 convs on tcgen05
    @staticmethod
    def _io(S, key, kind):
        e = S.entries[key + ".weight"]
        sh = e.shape
        if e.kind == "gconv3":           # grouped conv: (cout, cin / groups, 3, 3), cin == cout
            return sh[0], sh[0]
        return (sh[1], sh[0]) if kind not in (DECONV_S2, DECONV4_S2) else (sh[0], sh[1])  # (cin, cout)

    @staticmethod
    def _groups(S, key):
        e = S.entries[key + ".weight"]
        return e.shape[0] // e.shape[1] if e.kind == "gconv3" else 1

    # `key` may be a pair (key_a, key_b, split): images [0, split) go through conv key_a, the rest through key_b --
    # ONE grouped launch for forward / data gradient (full waves of CTA pairs instead of two half-size launches)
    @staticmethod
    def head_fusable(kind, cin, cout):
        """The decoder head can ride in the epilogue of the fused up-sampling kernel only (igemm.cu conv_up64_kernel)."""
        return (kind == DECONV_S2 and cout == 64 and cin <= 128 and os.environ.get("LSPS_NO_UP64", "0") != "1"
                and os.environ.get("LSPS_NO_HEAD_FUSE", "0") != "1")

    def conv_fwd(self, S, key, kind, x, lrelu, out=None, sums=None, split=False, head=None):
        """sums: optional fp32 [n,2,cout] -- the epilogue accumulates per-(image, channel) sum / sum of squares of the
        conv result there (LSPS_EP_STATS), which is all the following InstanceNorm needs.
        split: bf16x3 operands -- x / y are [n,h,w,2c] (hi | lo) tensors, the store carries the weight remainders.
        head = (head key, img fp32 [n,ho,wo], l1 or None): ConvTranspose2d(64,1,1) + Tanh (+ the L1 term, l1 as in
        Generator._head) computed in this launch's epilogue; only where head_fusable() says so."""
        n, h, w, cin = x.shape
        k0 = key[0] if isinstance(key, tuple) else key
        ci, co = self._io(S, k0, kind)
        assert (2 * ci if split else ci) == cin, (key, ci, cin)
        ho, wo = (h, w) if kind in (CONV_S1, CONV1X1) else ((h // 2, w // 2) if kind == CONV_S2 else (2 * h, 2 * w))
        y = out if out is not None else self.empty(n, ho, wo, 2 * co if split else co)
        flags = EP_BIAS | (EP_LRELU if lrelu else 0)
        if split:
            ext = ConvExt()
            ext.split, ext.w_lo = 1, S.W16L(key + ".weight").data_ptr()
            self.ctx.conv_fwd_ex(_shape(kind, n, h, w, ci, co), x.data_ptr(), S.W16(key + ".weight").data_ptr(),
                                 S.W(key + ".bias").data_ptr(), y.data_ptr(), flags, SLOPE, C.byref(ext))
            return y
        if sums is not None:
            ext = ConvExt()
            ext.sums = sums.data_ptr()
            ka = key
            if isinstance(key, tuple):
                ka, kb, split = key
                ext.w2, ext.bias2, ext.n_split = S.W16(kb + ".weight").data_ptr(), S.W(kb + ".bias").data_ptr(), split
            else:
                ext.groups = self._groups(S, key)
            self.ctx.conv_fwd_ex(_shape(kind, n, h, w, ci, co), x.data_ptr(), S.W16(ka + ".weight").data_ptr(),
                                 S.W(ka + ".bias").data_ptr(), y.data_ptr(), flags | EP_STATS, SLOPE, C.byref(ext))
            return y
        if head is not None:
            hk, img, l1 = head
            ext = ConvExt()
            ext.head_w, ext.head_b, ext.head_out = S.W(hk + ".weight").data_ptr(), S.W(hk + ".bias").data_ptr(), img.data_ptr()
            if l1 is not None:
                target, first, scale, dout, acc = l1
                px = img.shape[1] * img.shape[2]
                ext.head_target, ext.head_t0, ext.head_tn = target.data_ptr(), first * px, target.shape[0] * px
                ext.head_scale, ext.head_dout, ext.head_acc = scale, _lib.ptr(dout), acc.data_ptr()
            self.ctx.conv_fwd_ex(_shape(kind, n, h, w, ci, co), x.data_ptr(), S.W16(key + ".weight").data_ptr(),
                                 S.W(key + ".bias").data_ptr(), y.data_ptr(), flags, SLOPE, C.byref(ext))
            return y
        if isinstance(key, tuple):
            ka, kb, split = key
            self.ctx.conv_fwd_grouped(_shape(kind, n, h, w, ci, co), x.data_ptr(), S.W16(ka + ".weight").data_ptr(),
                                      S.W(ka + ".bias").data_ptr(), S.W16(kb + ".weight").data_ptr(),
                                      S.W(kb + ".bias").data_ptr(), split, y.data_ptr(), flags, SLOPE)
        else:
            self.ctx.conv_fwd(_shape(kind, n, h, w, ci, co), x.data_ptr(), S.W16(key + ".weight").data_ptr(),
                              S.W(key + ".bias").data_ptr(), y.data_ptr(), flags, SLOPE)
        return y

    def conv_dgrad(self, S, key, kind, dy, x_shape, mask=None, add=None, out=None, inbwd=None, split=False):
        """inbwd = (a, bsums): the gradient lands on a = lrelu(IN(h)); the epilogue applies lrelu'(xhat) and accumulates
        sum g / sum g*xhat per (image, channel) into bsums (LSPS_EP_INBWD; xhat is recovered from a)."""
        n, h, w, cin = x_shape
        k0 = key[0] if isinstance(key, tuple) else key
        ci, co = self._io(S, k0, kind)
        dx = out if out is not None else self.empty(n, h, w, 2 * ci if split else ci)
        flags = (EP_MASK if mask is not None else 0) | (EP_ADD if add is not None else 0)
        if split:
            assert add is None and inbwd is None
            ext = ConvExt()
            ext.split, ext.w_lo = 1, S.W16TL(key + ".weight").data_ptr()
            self.ctx.conv_dgrad_ex(_shape(kind, n, h, w, ci, co), dy.data_ptr(), S.W16T(key + ".weight").data_ptr(),
                                   dx.data_ptr(), _lib.ptr(mask), None, flags, SLOPE, C.byref(ext))
            return dx
        if inbwd is not None:
            assert mask is None
            ext = ConvExt()
            ext.in_a, ext.bsums = inbwd[0].data_ptr(), inbwd[1].data_ptr()
            ka = key
            if isinstance(key, tuple):
                ka, kb, split = key
                ext.w2, ext.n_split = S.W16T(kb + ".weight").data_ptr(), split
            else:
                ext.groups = self._groups(S, key)
            self.ctx.conv_dgrad_ex(_shape(kind, n, h, w, ci, co), dy.data_ptr(), S.W16T(ka + ".weight").data_ptr(),
                                   dx.data_ptr(), None, _lib.ptr(add), flags | EP_INBWD, SLOPE, C.byref(ext))
            return dx
        if isinstance(key, tuple):
            ka, kb, split = key
            self.ctx.conv_dgrad_grouped(_shape(kind, n, h, w, ci, co), dy.data_ptr(), S.W16T(ka + ".weight").data_ptr(),
                                        S.W16T(kb + ".weight").data_ptr(), split, dx.data_ptr(), _lib.ptr(mask),
                                        _lib.ptr(add), flags, SLOPE)
        else:
            self.ctx.conv_dgrad(_shape(kind, n, h, w, ci, co), dy.data_ptr(), S.W16T(key + ".weight").data_ptr(),
                                dx.data_ptr(), _lib.ptr(mask), _lib.ptr(add), flags, SLOPE)
        return dx

    def conv_wgrad(self, S, key, kind, x, dy, bias=True, split=False):
        if isinstance(key, tuple):       # weight gradients stay per conv: one launch per half
            ka, kb, nsplit = key
            self.conv_wgrad(S, ka, kind, x[:nsplit], dy[:nsplit], bias)
            self.conv_wgrad(S, kb, kind, x[nsplit:], dy[nsplit:], bias)
            return
        n, h, w, cin = x.shape
        ci, co = self._io(S, key, kind)

        groups = self._groups(S, key)

        def launch():
            if groups > 1:
                self.ctx.conv_wgrad_grouped(_shape(kind, n, h, w, ci, co), x.data_ptr(), dy.data_ptr(),
                                            S.G(key + ".weight").data_ptr(), groups)
                return
            if split:
                self.ctx.conv_wgrad_split(_shape(kind, n, h, w, ci, co), x.data_ptr(), dy.data_ptr(),
                                          S.G(key + ".weight").data_ptr())
                if bias:
                    self.ctx.colsum_bf16_split(dy.data_ptr(), dy.numel() // (2 * co), co, S.G(key + ".bias").data_ptr())
                return
            self.ctx.conv_wgrad(_shape(kind, n, h, w, ci, co), x.data_ptr(), dy.data_ptr(), S.G(key + ".weight").data_ptr())
            if bias:
                self.ctx.colsum_bf16(dy.data_ptr(), dy.numel() // co, co, S.G(key + ".bias").data_ptr())
        self._on_side(launch, x, dy)

    # ---- InstanceNorm
    def in_fwd(self, h, mode, res=None, out=None):
        n, hh, ww, c = h.shape
        stats = self.empty(n, c, 2, dtype=torch.float32)
        y = out if out is not None else torch.empty_like(h)
        self.ctx.instnorm_fwd(h.data_ptr(), _lib.ptr(res), y.data_ptr(), stats.data_ptr(), n, hh * ww, c, mode, IN_EPS,
                              SLOPE)
        return y, stats

    def in_bwd(self, dy, h, stats, mode, db=None, db2=None, split=0):
        """db: optional fp32 [c] accumulator for the bias gradient of the conv that produced h (fused column sum);
        db2/split: images [split, n) came from a second conv (grouped res block) and accumulate into db2."""
        n, hh, ww, c = h.shape
        dh = torch.empty_like(h)
        if db2 is not None:
            self.ctx.instnorm_bwd_grouped(dy.data_ptr(), h.data_ptr(), stats.data_ptr(), dh.data_ptr(), n, hh * ww, c,
                                          mode, SLOPE, db.data_ptr(), db2.data_ptr(), split)
        else:
            self.ctx.instnorm_bwd(dy.data_ptr(), h.data_ptr(), stats.data_ptr(), dh.data_ptr(), n, hh * ww, c, mode,
                                  SLOPE, _lib.ptr(db))
        return dh

    # ---- LeakyINSResBlock (common_net.py:160-181)
    @staticmethod
    def _sub(key, suffix):
        """conv key of a res block; a (block_a, block_b, split) pair maps to the pair of conv keys."""
        if isinstance(key, tuple):
            return (key[0] + suffix, key[1] + suffix, key[2])
        return key + suffix

    def res_fwd(self, S, key, x, save, out=None):
        if not isinstance(key, tuple) and (key + ".model.6.weight") in S.entries:
            return self.resx_fwd(S, key, x, save, out)
        if not self.fused_in:
            return self._res_fwd_old(S, key, x, save, out)
        n, hh, ww, c = x.shape
        ctx = self.ctx
        sums = self.empty(2, n, 2, c, dtype=torch.float32)         # epilogue accumulators of the two convs
        stats = self.empty(2, n, 2, c, dtype=torch.float32)        # (mean, rstd) rows, kept for the backward pass
        h1 = self.conv_fwd(S, self._sub(key, ".model.0"), CONV_S1, x, False, sums=sums[0])
        a1 = torch.empty_like(h1)
        ctx.norm_apply_fwd(h1.data_ptr(), None, a1.data_ptr(), sums[0].data_ptr(), stats[0].data_ptr(), n, hh * ww, c, 0, 1,
                           IN_EPS, SLOPE, None, None)
        h2 = self.conv_fwd(S, self._sub(key, ".model.3"), CONV_S1, a1, False, sums=sums[1])
        y = out if out is not None else torch.empty_like(h2)
        ctx.norm_apply_fwd(h2.data_ptr(), x.data_ptr(), y.data_ptr(), sums[1].data_ptr(), stats[1].data_ptr(), n, hh * ww, c,
                           1, 1, IN_EPS, SLOPE, None, None)
        if save is not None:
            save.append((key, x, None, stats[0], a1, h2, stats[1]))     # h1 is not needed: xhat1 is recovered from a1
        return y

    def res_bwd(self, S, saved, dout, wgrad=True, mask=None, out=None):
        """The biases of both convs feed an InstanceNorm: their gradient is exactly zero and is left at zero (the
        reference accumulates fp32 rounding noise there)."""
        if saved[0] == "resx":
            return self.resx_bwd(S, saved, dout, wgrad, mask, out)
        if not self.fused_in:
            return self._res_bwd_old(S, saved, dout, wgrad, mask, out)
        key, x, h1, st1, a1, h2, st2 = saved
        k0, k3 = self._sub(key, ".model.0"), self._sub(key, ".model.3")
        n, hh, ww, c = x.shape
        ctx, hw = self.ctx, hh * ww
        bs = self.empty(2, n, 2, c, dtype=torch.float32)
        # IN #2 (res + xhat): the gradient is dout itself; its two sums need one reduction pass
        ctx.norm_bwd_stats(dout.data_ptr(), h2.data_ptr(), st2.data_ptr(), bs[1].data_ptr(), n, hw, c, 1, 1, SLOPE, None, None)
        dh2 = torch.empty_like(h2)
        ctx.norm_bwd_apply(dout.data_ptr(), h2.data_ptr(), st2.data_ptr(), bs[1].data_ptr(), dh2.data_ptr(), n, hw, c, 1, 1,
                           SLOPE, None, None)
        if wgrad:
            self.conv_wgrad(S, k3, CONV_S1, a1, dh2, bias=False)
        # IN #1 (lrelu(xhat)): mask + sums in the data-gradient epilogue, then one apply pass
        g1 = self.conv_dgrad(S, k3, CONV_S1, dh2, a1.shape, inbwd=(a1, bs[0]))
        dh1 = torch.empty_like(a1)
        ctx.norm_bwd_apply(g1.data_ptr(), a1.data_ptr(), st1.data_ptr(), bs[0].data_ptr(), dh1.data_ptr(), n, hw, c, 2, 1,
                           SLOPE, None, None)
        if wgrad:
            self.conv_wgrad(S, k0, CONV_S1, x, dh1, bias=False)
        return self.conv_dgrad(S, k0, CONV_S1, dh1, x.shape, mask=mask, add=dout, out=out)

    # ---- LeakyINSResNeXtBlock (common_net.py:111-132): 1x1 -> IN -> lrelu -> grouped 3x3 -> IN -> lrelu -> 1x1 -> IN, + x.
    #      Same building blocks as the plain res block: statistics in each conv's epilogue, streaming apply passes,
    #      the lrelu(IN) backward front halves in the data-gradient epilogues.
    def resx_fwd(self, S, key, x, save, out=None):
        n, hh, ww, c = x.shape
        ctx, hw = self.ctx, hh * ww
        k0, k3, k6 = key + ".model.0", key + ".model.3", key + ".model.6"
        cm = S.entries[k0 + ".weight"].shape[0]                        # k * c channels inside the block
        sums_m, stats_m = self.empty(2, n, 2, cm, dtype=torch.float32), self.empty(2, n, 2, cm, dtype=torch.float32)
        sums_o, stats_o = self.empty(n, 2, c, dtype=torch.float32), self.empty(n, 2, c, dtype=torch.float32)
        h1 = self.conv_fwd(S, k0, CONV1X1, x, False, sums=sums_m[0])
        a1 = torch.empty_like(h1)
        ctx.norm_apply_fwd(h1.data_ptr(), None, a1.data_ptr(), sums_m[0].data_ptr(), stats_m[0].data_ptr(), n, hw, cm, 0, 1,
                           IN_EPS, SLOPE, None, None)
        h2 = self.conv_fwd(S, k3, CONV_S1, a1, False, sums=sums_m[1])
        a2 = torch.empty_like(h2)
        ctx.norm_apply_fwd(h2.data_ptr(), None, a2.data_ptr(), sums_m[1].data_ptr(), stats_m[1].data_ptr(), n, hw, cm, 0, 1,
                           IN_EPS, SLOPE, None, None)
        h3 = self.conv_fwd(S, k6, CONV1X1, a2, False, sums=sums_o)
        y = out if out is not None else torch.empty_like(h3)
        ctx.norm_apply_fwd(h3.data_ptr(), x.data_ptr(), y.data_ptr(), sums_o.data_ptr(), stats_o.data_ptr(), n, hw, c, 1, 1,
                           IN_EPS, SLOPE, None, None)
        if save is not None:
            save.append(("resx", key, x, a1, stats_m[0], a2, stats_m[1], h3, stats_o))
        return y

    def resx_bwd(self, S, saved, dout, wgrad=True, mask=None, out=None):
        _, key, x, a1, st1, a2, st2, h3, st3 = saved
        k0, k3, k6 = key + ".model.0", key + ".model.3", key + ".model.6"
        n, hh, ww, c = x.shape
        cm = a1.shape[-1]
        ctx, hw = self.ctx, hh * ww
        bs_o, bs_m = self.empty(n, 2, c, dtype=torch.float32), self.empty(2, n, 2, cm, dtype=torch.float32)
        ctx.norm_bwd_stats(dout.data_ptr(), h3.data_ptr(), st3.data_ptr(), bs_o.data_ptr(), n, hw, c, 1, 1, SLOPE, None, None)
        dh3 = torch.empty_like(h3)
        ctx.norm_bwd_apply(dout.data_ptr(), h3.data_ptr(), st3.data_ptr(), bs_o.data_ptr(), dh3.data_ptr(), n, hw, c, 1, 1, SLOPE, None, None)
        if wgrad:
            self.conv_wgrad(S, k6, CONV1X1, a2, dh3, bias=False)
        g2 = self.conv_dgrad(S, k6, CONV1X1, dh3, a2.shape, inbwd=(a2, bs_m[1]))
        dh2 = torch.empty_like(a2)
        ctx.norm_bwd_apply(g2.data_ptr(), a2.data_ptr(), st2.data_ptr(), bs_m[1].data_ptr(), dh2.data_ptr(), n, hw, cm, 2, 1, SLOPE, None, None)
        if wgrad:
            self.conv_wgrad(S, k3, CONV_S1, a1, dh2, bias=False)
        g1 = self.conv_dgrad(S, k3, CONV_S1, dh2, a1.shape, inbwd=(a1, bs_m[0]))
        dh1 = torch.empty_like(a1)
        ctx.norm_bwd_apply(g1.data_ptr(), a1.data_ptr(), st1.data_ptr(), bs_m[0].data_ptr(), dh1.data_ptr(), n, hw, cm, 2, 1, SLOPE, None, None)
        if wgrad:
            self.conv_wgrad(S, k0, CONV1X1, x, dh1, bias=False)
        return self.conv_dgrad(S, k0, CONV1X1, dh1, x.shape, mask=mask, add=dout, out=out)

    # -- the stand-alone InstanceNorm kernels (round 1), kept for A/B runs (LSPS_OLD_IN=1)
    def _res_fwd_old(self, S, key, x, save, out=None):
        h1 = self.conv_fwd(S, self._sub(key, ".model.0"), CONV_S1, x, False)
        a1, st1 = self.in_fwd(h1, 0)
        h2 = self.conv_fwd(S, self._sub(key, ".model.3"), CONV_S1, a1, False)
        y, st2 = self.in_fwd(h2, 1, res=x, out=out)
        if save is not None:
            save.append((key, x, h1, st1, a1, h2, st2))
        return y

    def _res_bwd_old(self, S, saved, dout, wgrad=True, mask=None, out=None):
        key, x, h1, st1, a1, h2, st2 = saved
        k0, k3 = self._sub(key, ".model.0"), self._sub(key, ".model.3")

        def dbs(k):
            if not wgrad:
                return {}
            if isinstance(k, tuple):
                return dict(db=S.G(k[0] + ".bias"), db2=S.G(k[1] + ".bias"), split=k[2])
            return dict(db=S.G(k + ".bias"))
        dh2 = self.in_bwd(dout, h2, st2, 1, **dbs(k3))
        if wgrad:
            self.conv_wgrad(S, k3, CONV_S1, a1, dh2, bias=False)
        da1 = self.conv_dgrad(S, k3, CONV_S1, dh2, a1.shape)
        dh1 = self.in_bwd(da1, h1, st1, 0, **dbs(k0))
        if wgrad:
            self.conv_wgrad(S, k0, CONV_S1, x, dh1, bias=False)
        return self.conv_dgrad(S, k0, CONV_S1, dh1, x.shape, mask=mask, add=dout, out=out)

    # ---- stems / head
    def stem_fwd(self, S, key, img, stride, out=None, split=False):
        n, h, w = img.shape
        y = out if out is not None else self.empty(n, h // stride, w // stride, 128 if split else 64)
        fn = self.ctx.stem_fwd_split if split else self.ctx.stem_fwd
        fn(img.data_ptr(), S.W(key + ".weight").data_ptr(), S.W(key + ".bias").data_ptr(), y.data_ptr(), n, h, w, stride,
           SLOPE)
        return y

    def stem_wgrad(self, S, key, img, dy, stride, split=False):
        n, h, w = img.shape
        fn = self.ctx.stem_wgrad_split if split else self.ctx.stem_wgrad
        self._on_side(lambda: fn(img.data_ptr(), dy.data_ptr(), S.G(key + ".weight").data_ptr(),
                                 S.G(key + ".bias").data_ptr(), n, h, w, stride), img, dy)

    def stem_dgrad(self, S, key, dy, dimg, stride, accumulate, split=False):
        n, h, w = dimg.shape
        fn = self.ctx.stem_dgrad_split if split else self.ctx.stem_dgrad
        fn(dy.data_ptr(), S.W(key + ".weight").data_ptr(), dimg.data_ptr(), n, h, w, stride, 1 if accumulate else 0)


class Generator:
    """SharedResGen forward/backward (lsps_nets.py:164-272)."""

    def __init__(self, ops, store, hp):
        self.ops, self.S, self.p = ops, store, hp
        assert hp["n_enc_front_blk"] == 3 and hp["n_gen_front_blk"] == 3 and hp["ch"] == 64, \
            "kernel set covers the reference configs (exps/nnyu.yaml, nicvl.yaml): 3 front blocks, ch=64"
        self.training = True  # the reference drivers never call gen.eval()
        # encoder-A/B (and cycle decoder-B/A) res blocks as grouped launches on the concatenated batch: any split on an
        # image boundary works, the two weight sets only have to sit in one flat buffer (params.ALIGN)
        self.grouped = os.environ.get("LSPS_NO_GROUP", "0") != "1" and hp.get("name") != "SharedResXGen"

    # -- encoder: 7x7 s1 stem, two 3x3 s2 convs, n_enc_res_blk res blocks
    def enc_fwd(self, dom, img, save, out=None):
        o, S, e = self.ops, self.S, "encode_%s" % dom
        f0 = o.stem_fwd(S, e + ".0.model.0", img, 1)
        f1 = o.conv_fwd(S, e + ".1.model.0", CONV_S2, f0, True)
        f2 = o.conv_fwd(S, e + ".2.model.0", CONV_S2, f1, True)
        blocks = [] if save is not None else None
        x, nres = f2, self.p["n_enc_res_blk"]
        for i in range(nres):
            x = o.res_fwd(S, "%s.%d" % (e, 3 + i), x, blocks, out=out if i == nres - 1 else None)
        if save is not None:
            save.append(dict(dom=dom, img=img, f0=f0, f1=f1, f2=f2, blocks=blocks))
        return x

    def enc_bwd(self, sv, dx, dimg=None):
        o, S, e = self.ops, self.S, "encode_%s" % sv["dom"]
        blocks = sv["blocks"]
        for i in range(len(blocks) - 1, -1, -1):
            dx = o.res_bwd(S, blocks[i], dx, mask=sv["f2"] if i == 0 else None)
        if not blocks:
            raise NotImplementedError("n_enc_res_blk == 0")
        o.conv_wgrad(S, e + ".2.model.0", CONV_S2, sv["f1"], dx)
        d1 = o.conv_dgrad(S, e + ".2.model.0", CONV_S2, dx, sv["f1"].shape, mask=sv["f1"])
        o.conv_wgrad(S, e + ".1.model.0", CONV_S2, sv["f0"], d1)
        d0 = o.conv_dgrad(S, e + ".1.model.0", CONV_S2, d1, sv["f0"].shape, mask=sv["f0"])
        o.stem_wgrad(S, e + ".0.model.0", sv["img"], d0, 1)
        if dimg is not None:
            o.stem_dgrad(S, e + ".0.model.0", d0, dimg, 1, accumulate=True)

    # -- both encoders on (xa | xb) with the res blocks of A and B grouped into single launches.  Same arithmetic as
    #    enc_fwd("A") + enc_fwd("B"): only the launch shape changes (one 2B-image GEMM instead of two B-image ones)
    def enc_pair_fwd(self, xa, xb, save, out):
        o, S = self.ops, self.S
        na = xa.shape[0]
        nres = self.p["n_enc_res_blk"]
        f2 = o.empty(na + xb.shape[0], 32, 32, 4 * self.p["ch"]) if nres else out
        fr = []
        for dom, img, dst in (("A", xa, f2[:na]), ("B", xb, f2[na:])):
            e = "encode_%s" % dom
            f0 = o.stem_fwd(S, e + ".0.model.0", img, 1)
            f1 = o.conv_fwd(S, e + ".1.model.0", CONV_S2, f0, True)
            o.conv_fwd(S, e + ".2.model.0", CONV_S2, f1, True, out=dst)
            fr.append(dict(dom=dom, img=img, f0=f0, f1=f1))
        blocks = [] if save is not None else None
        x = f2
        for i in range(nres):
            key = ("encode_A.%d" % (3 + i), "encode_B.%d" % (3 + i), na)
            x = o.res_fwd(S, key, x, blocks, out=out if i == nres - 1 else None)
        if save is not None:
            save.append(dict(pair=True, fronts=fr, f2=f2, blocks=blocks, na=na))
        return x

    def enc_pair_bwd(self, sv, dx, dimg_a=None, dimg_b=None):
        o, S, na = self.ops, self.S, sv["na"]
        blocks = sv["blocks"]
        if not blocks:
            raise NotImplementedError("n_enc_res_blk == 0")   # the LeakyReLU mask of f2 is applied by the first block's dgrad
        for i in range(len(blocks) - 1, -1, -1):
            dx = o.res_bwd(S, blocks[i], dx, mask=sv["f2"] if i == 0 else None)
        for fr, d2, dimg in ((sv["fronts"][0], dx[:na], dimg_a), (sv["fronts"][1], dx[na:], dimg_b)):
            e = "encode_%s" % fr["dom"]
            o.conv_wgrad(S, e + ".2.model.0", CONV_S2, fr["f1"], d2)
            d1 = o.conv_dgrad(S, e + ".2.model.0", CONV_S2, d2, fr["f1"].shape, mask=fr["f1"])
            o.conv_wgrad(S, e + ".1.model.0", CONV_S2, fr["f0"], d1)
            d0 = o.conv_dgrad(S, e + ".1.model.0", CONV_S2, d1, fr["f0"].shape, mask=fr["f0"])
            o.stem_wgrad(S, e + ".0.model.0", fr["img"], d0, 1)
            if dimg is not None:
                o.stem_dgrad(S, e + ".0.model.0", d0, dimg, 1, accumulate=True)

    # -- two decoders on the halves of x (doms = ("B", "A") in the cycle pass, ("A", "B") for gen.decode): grouped res
    #    blocks, then each domain's transposed convs + head on its half
    def dec_pair_fwd(self, doms, x, save, l1s=(None, None)):
        o, S = self.ops, self.S
        n = x.shape[0] // 2
        blocks = [] if save is not None else None
        nres = self.p["n_gen_res_blk"]
        for i in range(nres):
            x = o.res_fwd(S, ("decode_%s.%d" % (doms[0], i), "decode_%s.%d" % (doms[1], i), n), x, blocks)
        outs, tails = [], []
        for dom, xh, l1 in ((doms[0], x[:n], l1s[0]), (doms[1], x[n:], l1s[1])):
            d = "decode_%s" % dom
            g1 = o.conv_fwd(S, "%s.%d.model.0" % (d, nres), DECONV_S2, xh, True)
            g2, img = self._deconv_head("%s.%d.model.0" % (d, nres + 1), "%s.%d" % (d, nres + 2), g1,
                                        lambda n_, h_, w_: o.empty(n_, h_, w_, dtype=torch.float32), l1)
            outs.append(img)
            tails.append(dict(dom=dom, blocks=[], x3=xh, g1=g1, g2=g2, out=img))
        if save is not None:
            save.append(dict(pair=True, tails=tails, blocks=blocks, n=n))
        return outs

    def dec_pair_bwd(self, sv, douts, out=None):
        """douts: gradients w.r.t. the two image halves.  Returns the gradient w.r.t. the [2n] input of dec_pair_fwd."""
        o, S, n = self.ops, self.S, sv["n"]
        nb = len(sv["blocks"])
        dx = (out if (out is not None and nb == 0) else o.empty(2 * n, 32, 32, 4 * self.p["ch"]))
        self.dec_bwd(sv["tails"][0], douts[0], out=dx[:n])     # tails carry no res blocks: head + transposed convs only
        self.dec_bwd(sv["tails"][1], douts[1], out=dx[n:])
        for i in range(nb - 1, -1, -1):
            dx = o.res_bwd(S, sv["blocks"][i], dx, out=out if i == 0 else None)
        return dx

    # -- shared latent: res blocks + GaussianNoiseLayer (+ KL sum of the noised latent) ; then dec_shared
    def shared_fwd(self, x, noise, kl_acc, save, pre=None):
        """pre: optional (x_pre_noise, eb) from an earlier run of the deterministic part (see forward(front=...))."""
        o, S = self.ops, self.S
        if pre is not None:
            x, eb = pre
        else:
            eb = []
            for i in range(self.p["n_enc_shared_blk"]):
                x = o.res_fwd(S, "enc_shared.%d" % i, x, eb)
        if noise is not None:
            z = torch.empty_like(x)
            o.noise_kl(x, noise, z, kl_acc)
        else:
            z = x
        db = [] if save is not None else None
        y = z
        for i in range(self.p["n_gen_shared_blk"]):
            y = o.res_fwd(S, "dec_shared.%d" % i, y, db)
        if save is not None:
            save.append(dict(eb=eb, db=db, z=z))
        return y, z

    def shared_bwd(self, sv, dy, kl_alpha, dz_extra=None):
        """kl_alpha = d(loss)/d(sum z^2): the KL term contributes 2*kl_alpha*z to dz."""
        o, S = self.ops, self.S
        for blk in reversed(sv["db"]):
            dy = o.res_bwd(S, blk, dy)
        z = sv["z"]
        dz = torch.empty_like(z)
        o.ctx.axpy_bf16(dy.data_ptr(), z.data_ptr(), 2.0 * kl_alpha, dz.data_ptr(), z.numel())
        if dz_extra is not None:
            o.ctx.axpy_bf16(dz.data_ptr(), dz_extra.data_ptr(), 1.0, dz.data_ptr(), z.numel())
        for blk in reversed(sv["eb"]):
            dz = o.res_bwd(S, blk, dz)
        return dz

    # -- decoder: res blocks, two transposed 3x3 s2 convs, 1x1 head + tanh
    def _deconv_head(self, ck, hk, g1, img_of, l1):
        """last transposed conv + decoder head: one launch where the head fits the conv's epilogue, else two.
        img_of(g2 shape) -> the fp32 image tensor to fill.  Returns (g2, img)."""
        o, S = self.ops, self.S
        n, h, w, cin = g1.shape
        _, co = o._io(S, ck, DECONV_S2)
        img = img_of(n, 2 * h, 2 * w)
        if o.head_fusable(DECONV_S2, cin, co):
            return o.conv_fwd(S, ck, DECONV_S2, g1, True, head=(hk, img, l1)), img
        g2 = o.conv_fwd(S, ck, DECONV_S2, g1, True)
        self._head(hk, g2, img, l1)
        return g2, img

    def _head(self, hk, g2, img, l1):
        """decoder head (1x1 transposed conv + tanh); l1 = (target [nt,128,128] fp32, first image, scale, dout, acc): the
        L1 reconstruction loss of images [first, first + nt) and its gradient are taken in the same kernel."""
        o, S = self.ops, self.S
        if l1 is None:
            o.ctx.head_fwd(g2.data_ptr(), S.W(hk + ".weight").data_ptr(), S.W(hk + ".bias").data_ptr(), img.data_ptr(),
                           img.numel())
            return
        target, first, scale, dout, acc = l1
        px = img.shape[1] * img.shape[2]
        o.ctx.head_fwd_l1(g2.data_ptr(), S.W(hk + ".weight").data_ptr(), S.W(hk + ".bias").data_ptr(), img.data_ptr(),
                          img.numel(), target.data_ptr(), first * px, target.shape[0] * px, scale, _lib.ptr(dout),
                          acc.data_ptr())

    def dec_fwd(self, dom, x, save, out=None, l1=None):
        o, S, d = self.ops, self.S, "decode_%s" % dom
        blocks = [] if save is not None else None
        nres = self.p["n_gen_res_blk"]
        for i in range(nres):
            x = o.res_fwd(S, "%s.%d" % (d, i), x, blocks)
        g1 = o.conv_fwd(S, "%s.%d.model.0" % (d, nres), DECONV_S2, x, True)
        g2, img = self._deconv_head("%s.%d.model.0" % (d, nres + 1), "%s.%d" % (d, nres + 2), g1,
                                    lambda n_, h_, w_: out if out is not None else o.empty(n_, h_, w_, dtype=torch.float32), l1)
        if save is not None:
            save.append(dict(dom=dom, blocks=blocks, x3=x, g1=g1, g2=g2, out=img))
        return img

    def dec_bwd(self, sv, dout, out=None):
        o, S, d = self.ops, self.S, "decode_%s" % sv["dom"]
        nres = self.p["n_gen_res_blk"]
        g2, g1, x3 = sv["g2"], sv["g1"], sv["x3"]
        hk = "%s.%d" % (d, nres + 2)
        dg2 = torch.empty_like(g2)
        o.ctx.head_bwd(g2.data_ptr(), S.W(hk + ".weight").data_ptr(), sv["out"].data_ptr(), dout.data_ptr(),
                       dg2.data_ptr(), S.G(hk + ".weight").data_ptr(), S.G(hk + ".bias").data_ptr(), dout.numel(), SLOPE)
        k4, k3 = "%s.%d.model.0" % (d, nres + 1), "%s.%d.model.0" % (d, nres)
        o.conv_wgrad(S, k4, DECONV_S2, g1, dg2)
        dg1 = o.conv_dgrad(S, k4, DECONV_S2, dg2, g1.shape, mask=g1)
        o.conv_wgrad(S, k3, DECONV_S2, x3, dg1)
        nb = len(sv["blocks"])
        dx = o.conv_dgrad(S, k3, DECONV_S2, dg1, x3.shape, out=out if nb == 0 else None)
        for i in range(nb - 1, -1, -1):
            dx = o.res_bwd(S, sv["blocks"][i], dx, out=out if i == 0 else None)
        return dx

    # -- full forward (lsps_nets.py:250-258): returns images (x_aa|x_ba) and (x_ab|x_bb) as [2n,128,128] tensors
    def front_fwd(self, xa, xb):
        """The deterministic front of forward(): both encoders and the enc_shared res blocks, up to (not including) the
        GaussianNoiseLayer, with everything their backward needs.  dis_update and gen_update of one training step run
        the generator on the SAME images with the SAME weights (only the noise draw differs, lsps_trainer.py:79,145), so
        the trainer computes this part once per step and hands it to both (exact reuse, not an approximation)."""
        o, S = self.ops, self.S
        na = xa.shape[0] if xa is not None else 0
        nb = xb.shape[0] if xb is not None else 0
        h = o.empty(na + nb, 32, 32, 4 * self.p["ch"])
        se = []
        if na and nb and self.grouped:
            self.enc_pair_fwd(xa, xb, se, h)
        else:
            if na:
                self.enc_fwd("A", xa, se, out=h[:na])
            if nb:
                self.enc_fwd("B", xb, se, out=h[na:])
        eb = []
        x = h
        for i in range(self.p["n_enc_shared_blk"]):
            x = o.res_fwd(S, "enc_shared.%d" % i, x, eb)
        return dict(se=se, eb=eb, x=x, na=na, nb=nb)

    def forward(self, xa, xb, noise, kl_acc, save=None, out_a=None, out_b=None, front=None, l1_a=None, l1_b=None):
        """xa [na,128,128] / xb [nb,128,128] (either may be None).  Returns decode_A and decode_B of ALL na+nb
        latents: oa = (x_aa | x_ba), ob = (x_ab | x_bb), plus the noised shared latent.  out_a / out_b: optional
        [na+nb,128,128] fp32 destinations (slices of the discriminator's input batch: no concatenation copy).
        front: result of front_fwd(xa, xb) to reuse instead of recomputing the encoders."""
        o = self.ops
        if front is None:
            front = self.front_fwd(xa, xb)
        na, nb, se = front["na"], front["nb"], front["se"]
        ss = [] if save is not None else None
        y, z = self.shared_fwd(None, noise, kl_acc, ss, pre=(front["x"], front["eb"]))
        sd = [] if save is not None else None
        oa = self.dec_fwd("A", y, sd, out=out_a, l1=l1_a)
        ob = self.dec_fwd("B", y, sd, out=out_b, l1=l1_b)
        if save is not None:
            save.update(enc=se, shared=ss[0], dec=sd, na=na, nb=nb)
        return oa, ob, z

    def backward(self, save, doa, dob, kl_alpha, dz_extra=None):
        """doa/dob: fp32 gradients w.r.t. the [na+nb,128,128] outputs of decode_A / decode_B; dz_extra: optional bf16
        gradient added at the noised shared latent (the Mapping net's latent-matching term)."""
        na, nb = save["na"], save["nb"]
        dy = self.dec_bwd(save["dec"][0], doa)
        dy2 = self.dec_bwd(save["dec"][1], dob)
        self.ops.ctx.axpy_bf16(dy.data_ptr(), dy2.data_ptr(), 1.0, dy.data_ptr(), dy.numel())
        dh = self.shared_bwd(save["shared"], dy, kl_alpha, dz_extra)
        if save["enc"] and save["enc"][0].get("pair"):
            self.enc_pair_bwd(save["enc"][0], dh)
            return
        i = 0
        if na:
            self.enc_bwd(save["enc"][i], dh[:na])
            i += 1
        if nb:
            self.enc_bwd(save["enc"][i], dh[na:])

    # -- decode of an external latent (lsps_nets.py:239-243, used by the train_map branches): dec_shared on all 2n
    #    latents, decode_A on the first half and decode_B on the second -- the halves the reference keeps
    #    (lsps_trainer.py:92-93); the discarded halves carry no gradient and all layers are per-sample
    def decode_fwd(self, z, save=None):
        o, S = self.ops, self.S
        n = z.shape[0] // 2
        db = [] if save is not None else None
        y = z
        for i in range(self.p["n_gen_shared_blk"]):
            y = o.res_fwd(S, "dec_shared.%d" % i, y, db)
        sd = [] if save is not None else None
        if self.grouped:
            dec_a, dec_b = self.dec_pair_fwd(("A", "B"), y, sd)
        else:
            dec_a = self.dec_fwd("A", y[:n], sd)
            dec_b = self.dec_fwd("B", y[n:], sd)
        if save is not None:
            save.update(db=db, dec=sd, n=n)
        return dec_a, dec_b

    def decode_bwd(self, save, d_a, d_b):
        """Returns the gradient w.r.t. the latent handed to decode_fwd."""
        o, S, n = self.ops, self.S, save["n"]
        if save["dec"][0].get("pair"):
            dy = self.dec_pair_bwd(save["dec"][0], (d_a, d_b))
        else:
            dy = o.empty(2 * n, 32, 32, 4 * self.p["ch"])
            self.dec_bwd(save["dec"][0], d_a, out=dy[:n])
            self.dec_bwd(save["dec"][1], d_b, out=dy[n:])
        for blk in reversed(save["db"]):
            dy = o.res_bwd(S, blk, dy)
        return dy

    # -- cycle passes (lsps_nets.py:260-272), batched: first half a2b (encode_A -> decode_B), second half b2a
    def forward_cycle(self, x_ba, x_ab, noise, kl_acc_bab, kl_acc_aba, save=None, l1s=(None, None)):
        o = self.ops
        n = x_ba.shape[0]
        h = o.empty(2 * n, 32, 32, 4 * self.p["ch"])
        se = [] if save is not None else None
        if self.grouped:
            self.enc_pair_fwd(x_ba, x_ab, se, h)
        else:
            self.enc_fwd("A", x_ba, se, out=h[:n])
            self.enc_fwd("B", x_ab, se, out=h[n:])
        ss = [] if save is not None else None
        # two KL sums (bab / aba halves): run the noise kernel per half via shared_fwd's single call on a split
        y, z = self._shared_fwd_split(h, noise, kl_acc_bab, kl_acc_aba, n, ss)
        sd = [] if save is not None else None
        if self.grouped:
            x_bab, x_aba = self.dec_pair_fwd(("B", "A"), y, sd, l1s)
        else:
            x_bab = self.dec_fwd("B", y[:n], sd, l1=l1s[0])
            x_aba = self.dec_fwd("A", y[n:], sd, l1=l1s[1])
        if save is not None:
            save.update(enc=se, shared=ss[0], dec=sd, n=n)
        return x_bab, x_aba

    def _shared_fwd_split(self, x, noise, acc0, acc1, n, save):
        o, S = self.ops, self.S
        eb = [] if save is not None else None
        for i in range(self.p["n_enc_shared_blk"]):
            x = o.res_fwd(S, "enc_shared.%d" % i, x, eb)
        z = torch.empty_like(x)
        if isinstance(noise, tuple):     # device mode: two draws with distinct Philox offsets
            o.noise_kl(x[:n], noise, z[:n], acc0)
            o.noise_kl(x[n:], (noise[0], noise[1], noise[2] + 1), z[n:], acc1)
        else:
            o.noise_kl(x[:n], noise[:n], z[:n], acc0)
            o.noise_kl(x[n:], noise[n:], z[n:], acc1)
        db = [] if save is not None else None
        y = z
        for i in range(self.p["n_gen_shared_blk"]):
            y = o.res_fwd(S, "dec_shared.%d" % i, y, db)
        if save is not None:
            save.append(dict(eb=eb, db=db, z=z))
        return y, z

    def backward_cycle(self, save, d_bab, d_aba, kl_alpha, dimg_ba, dimg_ab):
        n = save["n"]
        if save["dec"][0].get("pair"):
            dy = self.dec_pair_bwd(save["dec"][0], (d_bab, d_aba))
        else:
            dy = self.ops.empty(2 * n, 32, 32, 4 * self.p["ch"])
            self.dec_bwd(save["dec"][0], d_bab, out=dy[:n])
            self.dec_bwd(save["dec"][1], d_aba, out=dy[n:])
        dh = self.shared_bwd(save["shared"], dy, kl_alpha)
        if save["enc"][0].get("pair"):
            self.enc_pair_bwd(save["enc"][0], dh, dimg_a=dimg_ba, dimg_b=dimg_ab)
        else:
            self.enc_bwd(save["enc"][0], dh[:n], dimg=dimg_ba)
            self.enc_bwd(save["enc"][1], dh[n:], dimg=dimg_ab)


class Mapping:
    """Mapping net (lsps_nets.py:8-31): pose latent (m, 20) -> shared latent (m, 32, 32, 256).  Layer 0 is a transposed
    4x4 conv on a 1x1 input, i.e. a dense layer 20 -> 16*4ch (small dense kernels); layers 1-3 are 4x4 stride-2
    transposed convs on the tcgen05 implicit-GEMM path (16 taps, 4 sub-pixel phases x 2x2 taps each)."""
    KEYS = ("model.0.model.0", "model.1.model.0", "model.2.model.0", "model.3")

    def __init__(self, ops, store, hp):
        self.ops, self.S, self.p = ops, store, hp
        assert hp["output_ch"] % 64 == 0

    def forward(self, e, save=None):
        o, S, (k0, k1, k2, k3) = self.ops, self.S, self.KEYS
        e = e.contiguous().float()
        m, d = e.shape
        c0 = 4 * self.p["output_ch"]
        bias16 = S.W(k0 + ".bias").repeat(16)                 # bias per (r, s, co) output column of the dense form
        y0f = o.empty(m, 16 * c0, dtype=torch.float32)
        o.ctx.linear_fwd(e.data_ptr(), 0, S.W(k0 + ".weight").data_ptr(), bias16.data_ptr(), y0f.data_ptr(), m, 16 * c0,
                         d, _lib.ACT_LRELU, SLOPE)
        y0 = o.empty(m, 4, 4, c0)
        o.ctx.f32_to_bf16(y0f.data_ptr(), y0.data_ptr(), y0.numel())
        y1 = o.conv_fwd(S, k1, DECONV4_S2, y0, True)
        y2 = o.conv_fwd(S, k2, DECONV4_S2, y1, True)
        z = o.conv_fwd(S, k3, DECONV4_S2, y2, False)
        if save is not None:
            save.update(e=e, y0=y0, y1=y1, y2=y2)
        return z

    def backward(self, sv, dz):
        """dz: bf16 gradient w.r.t. the output latent.  The gradient w.r.t. the pose latent is not formed: it would only
        reach the poseVAE encoder, which no optimiser steps from the GAN updates (lsps_trainer.py:26-31)."""
        o, S, (k0, k1, k2, k3) = self.ops, self.S, self.KEYS
        e, y0, y1, y2 = sv["e"], sv["y0"], sv["y1"], sv["y2"]
        o.conv_wgrad(S, k3, DECONV4_S2, y2, dz)
        d2 = o.conv_dgrad(S, k3, DECONV4_S2, dz, y2.shape, mask=y2)
        o.conv_wgrad(S, k2, DECONV4_S2, y1, d2)
        d1 = o.conv_dgrad(S, k2, DECONV4_S2, d2, y1.shape, mask=y1)
        o.conv_wgrad(S, k1, DECONV4_S2, y0, d1)
        d0 = o.conv_dgrad(S, k1, DECONV4_S2, d1, y0.shape, mask=y0)
        m, c0 = y0.shape[0], y0.shape[3]
        o.ctx.colsum_bf16(d0.data_ptr(), m * 16, c0, S.G(k0 + ".bias").data_ptr())
        d0f = o.empty(m, 16 * c0, dtype=torch.float32)
        o.ctx.bf16_to_f32(d0.data_ptr(), d0f.data_ptr(), d0.numel())
        o.ctx.linear_bwd(e.data_ptr(), 0, S.W(k0 + ".weight").data_ptr(), d0f.data_ptr(), None, 0,
                         S.G(k0 + ".weight").data_ptr(), None, m, 16 * c0, e.shape[1])

    def __call__(self, e):
        """Reference-style call: returns (m, 256, 32, 32) fp32."""
        return self.forward(e).permute(0, 3, 1, 2).float()

    def state_dict(self):
        return self.S.state_dict()

    def load_state_dict(self, sd, strict=True):
        self.S.load_state_dict(sd, strict)


class Discriminator:
    """SharedDis forward/backward (lsps_nets.py:86-160).

    With a split store (ParamStore(split=True)) the whole stack runs on the bf16x3 kernels: every activation and
    gradient tensor is [.., 2c] = (bf16 hi | bf16 lo) channel halves, every GEMM accumulates hi*hi + hi*lo + lo*hi.  The
    discriminator is ~1 % of the step's FLOPs but its bf16 rounding is ~80 % of the adversarial-loss deviation from the
    fp32 reference (profiles/r02_precision_ablation.json), so this is where the extra tensor-core passes go."""

    def __init__(self, ops, store, hp):
        self.ops, self.S, self.p = ops, store, hp
        assert hp["n_front_layer"] == 2 and hp["ch"] == 64, "kernel set covers the reference configs"
        self.training = True
        self.split = bool(getattr(store, "split", False))
        self.cf = hp["ch"] * 2 ** (hp["n_front_layer"] - 1 + hp["n_shared_layer"])   # trunk feature channels (2048)
        # called with the conv key right after the LAST trunk layer's weight gradient (75 MB of the 101 MB gradient
        # buffer, the first one the backward pass finishes) has been enqueued: the trainer starts its allreduce there
        self.on_tail_wgrad = None

    def train(self, mode=True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    def _c(self, c):
        return 2 * c if self.split else c

    def front_fwd(self, dom, img, save, out=None):
        o, S, m = self.ops, self.S, "model_%s" % dom
        f0 = o.stem_fwd(S, m + ".0.model.0", img, 2, split=self.split)
        f1 = o.conv_fwd(S, m + ".1.model.0", CONV_S2, f0, True, out=out, split=self.split)
        if save is not None:
            save.append(dict(dom=dom, img=img, f0=f0))
        return f1

    def front_bwd(self, sv, d1, wgrad, dimg=None, f1=None):
        o, S, m = self.ops, self.S, "model_%s" % sv["dom"]
        f0 = sv["f0"]
        n, h, w, _ = f0.shape
        if wgrad:
            o.conv_wgrad(S, m + ".1.model.0", CONV_S2, f0, d1, split=self.split)
        d0 = o.conv_dgrad(S, m + ".1.model.0", CONV_S2, d1, (n, h, w, self.p["ch"]), mask=f0, split=self.split)
        if wgrad:
            o.stem_wgrad(S, m + ".0.model.0", sv["img"], d0, 2, split=self.split)
        if dimg is not None:
            o.stem_dgrad(S, m + ".0.model.0", d0, dimg, 2, accumulate=False, split=self.split)

    def trunk_fwd(self, x, save):
        o, S = self.ops, self.S
        acts = [x]
        for i in range(self.p["n_shared_layer"]):
            x = o.conv_fwd(S, "model_S.%d.model.0" % i, CONV_S2, x, True, split=self.split)
            acts.append(x)
        if save is not None:
            save["acts"] = acts
        return x

    def trunk_bwd(self, acts, d, wgrad):
        """d: gradient w.r.t. the trunk features already multiplied by their LeakyReLU mask.  Returns the gradient
        w.r.t. the pre-activation of the last front conv (masked)."""
        o, S = self.ops, self.S
        for i in range(self.p["n_shared_layer"] - 1, -1, -1):
            key = "model_S.%d.model.0" % i
            if wgrad:
                o.conv_wgrad(S, key, CONV_S2, acts[i], d, split=self.split)
                if i == self.p["n_shared_layer"] - 1 and self.on_tail_wgrad is not None:
                    self.on_tail_wgrad(key)
            n, h, w, c = acts[i].shape
            d = o.conv_dgrad(S, key, CONV_S2, d, (n, h, w, c // 2 if self.split else c), mask=acts[i], split=self.split)
        return d

    def features(self, imgs_a, imgs_b, save=None):
        """model_A / model_B fronts on their image batches, concatenated along batch, shared trunk."""
        o = self.ops
        na = imgs_a.shape[0] if imgs_a is not None else 0
        nb = imgs_b.shape[0] if imgs_b is not None else 0
        x = o.empty(na + nb, 32, 32, self._c(2 * self.p["ch"]))
        fr = [] if save is not None else None
        if na:
            self.front_fwd("A", imgs_a, fr, out=x[:na])
        if nb:
            self.front_fwd("B", imgs_b, fr, out=x[na:])
        if save is not None:
            save.update(fronts=fr, na=na, nb=nb)
        return self.trunk_fwd(x, save)

    def features_bwd(self, save, dF_masked, wgrad, dimg_a=None, dimg_b=None):
        d = self.trunk_bwd(save["acts"], dF_masked, wgrad)
        na, i = save["na"], 0
        if na:
            self.front_bwd(save["fronts"][i], d[:na], wgrad, dimg=dimg_a)
            i += 1
        if save["nb"]:
            self.front_bwd(save["fronts"][i], d[na:], wgrad, dimg=dimg_b)

    # ---- heads and feature losses on the trunk features F [n, 2, 2, cf] (split: [n, 2, 2, 2 cf]).  Gradients w.r.t. F
    #      are accumulated in a plain fp32 [n, 4 cf] buffer (dF) and masked / packed once by mask_grad().
    def rows(self, F):
        return F.shape[0] * F.shape[1] * F.shape[2]

    def _fptr(self, F, img):
        return F.data_ptr() + img * F[0].numel() * F.element_size()

    def logits(self, F):
        o, S = self.ops, self.S
        lg = o.empty(self.rows(F), dtype=torch.float32)
        fn = o.ctx.dhead_fwd_split if self.split else o.ctx.dhead_fwd
        fn(F.data_ptr(), S.W("D.weight").data_ptr(), S.W("D.bias").data_ptr(), lg.data_ptr(), self.rows(F), self.cf)
        return lg

    def logits_bwd(self, F, dlg, dF, wgrad):
        o, S = self.ops, self.S
        fn = o.ctx.dhead_bwd_split if self.split else o.ctx.dhead_bwd
        fn(F.data_ptr(), S.W("D.weight").data_ptr(), dlg.data_ptr(), dF.data_ptr(),
           S.G("D.weight").data_ptr() if wgrad else None, S.G("D.bias").data_ptr() if wgrad else None, self.rows(F), self.cf)

    def head_bce(self, F, rows_per_group, targets, slots, scale, dF, wgrad, acc):
        """D head + sigmoid + BCE + backward in one launch (lsps_dhead_bce): targets[g] is the constant BCE target of
        image group g (-1: the group's logits are not used), its loss sum / correct count go to acc[slots[g]], [+1];
        dF [n, 4 cf] fp32 is written completely (no zero fill)."""
        o, S = self.ops, self.S
        ng = len(targets)
        o.ctx.dhead_bce(F.data_ptr(), S.W("D.weight").data_ptr(), S.W("D.bias").data_ptr(), self.rows(F), self.cf,
                        1 if self.split else 0, rows_per_group, ng, (C.c_float * ng)(*targets), (C.c_int * ng)(*slots), scale,
                        None, dF.data_ptr(), S.G("D.weight").data_ptr() if wgrad else None,
                        S.G("D.bias").data_ptr() if wgrad else None, acc.data_ptr())

    def l1_feat(self, F, img_a, img_b, nimg, dF, scale, acc):
        """acc += sum |F[img_a:img_a+nimg] - F[img_b:img_b+nimg]| ; dF rows get +-scale*sign (lsps_trainer.py:171-177)."""
        o = self.ops
        per = 4 * self.cf
        args = (self._fptr(F, img_a), self._fptr(F, img_b), dF[img_a:].data_ptr(), dF[img_b:].data_ptr(), scale,
                acc.data_ptr(), nimg * per)
        if self.split:
            o.ctx.l1_feat_split(*args, self.cf)
        else:
            o.ctx.l1_feat(*args)

    def mask_grad(self, dF, F):
        """dF fp32 [n, 4 cf] -> gradient w.r.t. the pre-activation of the last trunk conv, in F's storage format."""
        o = self.ops
        out = torch.empty_like(F)
        if self.split:
            o.ctx.mask_to_bf16_split(dF.data_ptr(), F.data_ptr(), out.data_ptr(), SLOPE, dF.numel(), self.cf)
        else:
            o.ctx.mask_to_bf16(dF.data_ptr(), F.data_ptr(), out.data_ptr(), SLOPE, dF.numel())
        return out

    def post(self, F, img0=0, nimg=None):
        """Post = Conv2d(2048, post_dim, 2) on the 2x2 map == FC 8192 -> post_dim (lsps_nets.py:123,135-145)."""
        o, S = self.ops, self.S
        n = F.shape[0] - img0 if nimg is None else nimg
        k = 4 * self.cf
        pd = self.p["post_dim"]
        out = o.empty(n, pd, dtype=torch.float32)
        o.ctx.linear_fwd(self._fptr(F, img0), self.cf if self.split else 1, S.W("Post.weight").data_ptr(),
                         S.W("Post.bias").data_ptr(), out.data_ptr(), n, pd, k, 0, SLOPE)
        return out

    def post_bwd(self, F, img0, nimg, dp, dF):
        """dF[img0:img0+nimg] = dp . W_post (overwrites those rows) ; Post weight / bias gradients accumulate."""
        o, S = self.ops, self.S
        o.ctx.linear_bwd(self._fptr(F, img0), self.cf if self.split else 1, S.W("Post.weight").data_ptr(), dp.data_ptr(),
                         dF[img0:].data_ptr(), 0, S.G("Post.weight").data_ptr(), S.G("Post.bias").data_ptr(), nimg,
                         self.p["post_dim"], 4 * self.cf)

    # public inference API used by the drivers (depth_train.py:197-206)
    def _regress(self, dom, x):
        img = x.reshape(x.shape[0], x.shape[-2], x.shape[-1]).contiguous().float()
        F = self.features(img if dom == "A" else None, img if dom == "B" else None)
        p = self.post(F).squeeze()
        return p, p, p

    def regress_a(self, x):
        return self._regress("A", x)

    def regress_b(self, x):
        return self._regress("B", x)

    def state_dict(self):
        return self.S.state_dict()

    def load_state_dict(self, sd, strict=True):
        self.S.load_state_dict(sd, strict)


class PoseVAE:
    """poseVAE MLP (lsps_nets.py:34-83) on the small dense kernels."""

    def __init__(self, ops, store, hp, noise_fn):
        self.ops, self.S, self.p, self.noise_fn = ops, store, hp, noise_fn

    def _lin(self, key, x, act):
        o, S = self.ops, self.S
        w = S.entries[key + ".weight"].shape
        y = o.empty(x.shape[0], w[0], dtype=torch.float32)
        o.ctx.linear_fwd(x.data_ptr(), 0, S.W(key + ".weight").data_ptr(), S.W(key + ".bias").data_ptr(), y.data_ptr(),
                         x.shape[0], w[0], w[1], act, SLOPE)
        return y

    def _lin_bwd(self, key, x, dy, need_dx, acc_into=None):
        o, S = self.ops, self.S
        w = S.entries[key + ".weight"].shape
        dx = acc_into if acc_into is not None else (o.empty(x.shape[0], w[1], dtype=torch.float32) if need_dx else None)
        o.ctx.linear_bwd(x.data_ptr(), 0, S.W(key + ".weight").data_ptr(), dy.data_ptr(), _lib.ptr(dx),
                         1 if acc_into is not None else 0,
                         S.G(key + ".weight").data_ptr(), S.G(key + ".bias").data_ptr(), x.shape[0], w[0], w[1])
        return dx

    def encode(self, y, kl_acc=None, save=None):
        o = self.ops
        y = y.contiguous().float()
        h = self._lin("en_fc1", y, _lib.ACT_LRELU)
        mu = self._lin("en_mu", h, _lib.ACT_NONE)
        sd = self._lin("en_sigma", h, _lib.ACT_SOFTPLUS)
        noise = self.noise_fn(tuple(mu.shape))
        z = torch.empty_like(mu)
        o.ctx.vae_reparam(mu.data_ptr(), sd.data_ptr(), noise.data_ptr(), z.data_ptr(), _lib.ptr(kl_acc), mu.numel())
        if save is not None:
            save.update(y=y, h=h, mu=mu, sd=sd, noise=noise, z=z)
        return z, mu, sd

    def decode(self, z, save=None):
        z = z.contiguous().float()
        if z.dim() == 1:
            z = z[None]
        h = self._lin("de_fc1.model.0", z, _lib.ACT_LRELU)
        out = self._lin("de_fc2", h, _lib.ACT_NONE)
        if save is not None:
            save.update(dz_in=z, dh=h)
        return out

    def forward(self, y, kl_acc=None, save=None):
        z, mu, sd = self.encode(y, kl_acc, save)
        return self.decode(z, save), z, mu, sd

    def backward(self, sv, ddec, kl_scale):
        """ddec: gradient w.r.t. the decoder output; kl_scale = d(loss)/d(KL sum)."""
        o = self.ops
        dh = self._lin_bwd("de_fc2", sv["dh"], ddec, True)
        o.ctx.act_bwd(dh.data_ptr(), sv["dh"].data_ptr(), _lib.ACT_LRELU, SLOPE, dh.numel())
        dz = self._lin_bwd("de_fc1.model.0", sv["dz_in"], dh, True)
        dmu, dsd = torch.empty_like(dz), torch.empty_like(dz)
        o.ctx.vae_reparam_bwd(sv["mu"].data_ptr(), sv["sd"].data_ptr(), sv["noise"].data_ptr(), dz.data_ptr(),
                              dmu.data_ptr(), dsd.data_ptr(), kl_scale, dz.numel())
        o.ctx.act_bwd(dsd.data_ptr(), sv["sd"].data_ptr(), _lib.ACT_SOFTPLUS, SLOPE, dsd.numel())
        dh1 = self._lin_bwd("en_mu", sv["h"], dmu, True)
        self._lin_bwd("en_sigma", sv["h"], dsd, True, acc_into=dh1)
        o.ctx.act_bwd(dh1.data_ptr(), sv["h"].data_ptr(), _lib.ACT_LRELU, SLOPE, dh1.numel())
        self._lin_bwd("en_fc1", sv["y"], dh1, False)

    _KEYS = ("en_fc1", "en_mu", "en_sigma", "de_fc1.model.0", "de_fc2")

    def step(self, y, ll_scale, kl_scale, acc):
        """forward + L1 / KL losses + backward of one vae_update in ONE launch (lsps_vae_step): the parameter gradients
        are added to the store's (zeroed) gradient buffer, acc[0] += KL sum, acc[1] += L1 sum.  Returns the
        reconstruction.  Same arithmetic as forward() / backward() on the small dense kernels."""
        o, S = self.ops, self.S
        y = y.contiguous().float()
        rows, d = y.shape
        h, z = S.entries["en_fc1.weight"].shape[0], S.entries["en_mu.weight"].shape[0]
        if getattr(self, "_ptrs", None) is None:
            wp, gp = (C.c_void_p * 10)(), (C.c_void_p * 10)()
            for i, k in enumerate(self._KEYS):
                wp[2 * i], wp[2 * i + 1] = S.W(k + ".weight").data_ptr(), S.W(k + ".bias").data_ptr()
                gp[2 * i], gp[2 * i + 1] = S.G(k + ".weight").data_ptr(), S.G(k + ".bias").data_ptr()
            self._ptrs = (wp, gp)
        wp, gp = self._ptrs
        noise = self.noise_fn((rows, z)).contiguous().float()
        dec = o.empty(rows, d, dtype=torch.float32)
        o.ctx.vae_step(y.data_ptr(), noise.data_ptr(), wp, gp, dec.data_ptr(), acc.data_ptr(), rows, d, h, z,
                       ll_scale, kl_scale, SLOPE)
        return dec

    def state_dict(self):
        return self.S.state_dict()

    def load_state_dict(self, sd, strict=True):
        self.S.load_state_dict(sd, strict)
