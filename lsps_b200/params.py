"""Flat parameter stores for the LSPS networks.

Each network (gen / dis / vae) owns ONE fp32 master buffer in *kernel* layout, one fp32 gradient buffer (the
buffer the NCCL allreduce runs on), Adam moments, and bf16 operand copies for the tcgen05 kernels.  `state_dict()`
/ `load_state_dict()` convert to and from the reference's key names and OIHW / IOHW shapes
(/root/reference/src/trainers/lsps_nets.py:34-272 -- key list in SURVEY.md section 8a), so reference snapshots load
unchanged; the packed bf16 tensors are derived data and never serialised.
"""
import math
from collections import OrderedDict

import torch

from . import _lib

ALIGN = 256  # elements: 16-byte alignment in every buffer, and any two conv weights are a whole number of 256-channel
             # GEMM-K rows apart, which lets ONE weight tensor map span two convs (grouped launches, engine.py)
ACC_SLOTS = 64  # float accumulators (loss sums) appended to the gradient buffer so they ride the same allreduce


def _round_up(n, a=ALIGN):
    return (n + a - 1) // a * a


# ------------------------------------------------------------------ layout conversions (reference <-> kernel)
def to_kernel_layout(kind, t):
    if kind in ("conv3", "gconv3"):   # OIHW (co,ci,3,3) -> [tap][co][ci]  (grouped: ci = channels of one group)
        return t.permute(2, 3, 0, 1).reshape(9, t.shape[0], t.shape[1])
    if kind == "deconv3":    # IOHW (ci,co,3,3) -> [tap][co][ci]
        return t.permute(2, 3, 1, 0).reshape(9, t.shape[1], t.shape[0])
    if kind in ("deconv4", "map0"):   # IOHW (ci,co,4,4) -> [tap][co][ci]; map0 (1x1 input) reads it as FC [16*co][ci]
        return t.permute(2, 3, 1, 0).reshape(16, t.shape[1], t.shape[0])
    if kind == "post":       # (20,c,2,2) -> [20][pos*c + ch]
        return t.permute(0, 2, 3, 1).reshape(t.shape[0], -1)
    return t.reshape(-1)     # stem (64,1,7,7)->[64][49]; head (64,1,1,1)->[64]; D (1,c,1,1)->[c]; linear; bias


def from_kernel_layout(kind, flat, shape):
    if kind in ("conv3", "gconv3"):
        co, ci = shape[0], shape[1]
        return flat.reshape(3, 3, co, ci).permute(2, 3, 0, 1).contiguous()
    if kind == "deconv3":
        ci, co = shape[0], shape[1]
        return flat.reshape(3, 3, co, ci).permute(3, 2, 0, 1).contiguous()
    if kind in ("deconv4", "map0"):
        ci, co = shape[0], shape[1]
        return flat.reshape(4, 4, co, ci).permute(3, 2, 0, 1).contiguous()
    if kind == "post":
        o, c = shape[0], shape[1]
        return flat.reshape(o, 2, 2, c).permute(0, 3, 1, 2).contiguous()
    return flat.reshape(shape).clone()


class Entry:
    __slots__ = ("key", "shape", "kind", "law", "fan_in", "off", "numel", "step", "dg_off")

    def __init__(self, key, shape, kind, law, fan_in):
        self.key, self.shape, self.kind, self.law, self.fan_in = key, tuple(shape), kind, law, fan_in
        self.numel = int(math.prod(shape))
        self.off = self.step = 0
        self.dg_off = -1


class ParamStore:
    """entries: list of (key, shape, kind, law, fan_in).
    kind in conv3|deconv3|deconv4|conv1|gconv3|map0|stem|head|dhead|post|linear|bias."""

    def __init__(self, entries, device, lr, weight_decay, betas=(0.5, 0.999), eps=1e-8, split=False):
        """split: keep a second bf16 copy of every operand tensor holding the rounding remainder (w - bf16(w)), for the
        split-bf16 ("bf16x3") kernels -- hi + lo carry 16 mantissa bits."""
        self.device = torch.device(device)
        self.split = bool(split)
        self.entries = OrderedDict()
        off = dg = 0
        for e in entries:
            ent = Entry(*e)
            ent.off = off
            off += _round_up(ent.numel)
            if ent.kind in ("conv3", "deconv3", "deconv4", "conv1", "gconv3"):
                ent.dg_off = dg
                dg += _round_up(ent.numel)
            self.entries[ent.key] = ent
        self.size = off
        kw = dict(device=self.device)
        self.w = torch.zeros(off, dtype=torch.float32, **kw)
        self.gbuf = torch.zeros(off + ACC_SLOTS, dtype=torch.float32, **kw)
        self.g = self.gbuf[:off]
        self.acc = self.gbuf[off:]
        self.m = torch.zeros(off, dtype=torch.float32, **kw)
        self.v = torch.zeros(off, dtype=torch.float32, **kw)
        self.w16 = torch.zeros(off, dtype=torch.bfloat16, **kw)           # forward operands [tap][co][ci]
        self.w16t = torch.zeros(max(dg, 8), dtype=torch.bfloat16, **kw)   # dgrad operands  [tap][ci][co]
        self.w16l = torch.zeros(off, dtype=torch.bfloat16, **kw) if self.split else None
        self.w16tl = torch.zeros(max(dg, 8), dtype=torch.bfloat16, **kw) if self.split else None
        # one-launch transposition of every conv weight (lsps_pack_dgrad_multi): {w_off, wt_off, taps, cout, cin, tile0}
        rows, tile0 = [], 0
        for ent in self.entries.values():
            if ent.dg_off < 0:
                continue
            co, ci = (ent.shape[0], ent.shape[1]) if ent.kind in ("conv3", "conv1", "gconv3") else (ent.shape[1], ent.shape[0])
            taps = {"deconv4": 16, "conv1": 1}.get(ent.kind, 9)
            if ent.kind == "gconv3":     # [tap][group][co in group][ci in group]: taps x groups square transposes
                taps, co = taps * (co // ci), ci
            rows.append([ent.off, ent.dg_off, taps, co, ci, tile0])
            tile0 += taps * ((ci + 31) // 32) * ((co + 31) // 32)
        self._pack_count, self._pack_tiles = len(rows), tile0
        rows.append([0, 0, 0, 0, 0, tile0])
        self._pack_desc = torch.tensor(rows, dtype=torch.int64).to(self.device) if self._pack_count else None
        self.lr, self.base_lr, self.wd, self.betas, self.eps = lr, lr, weight_decay, betas, eps
        self.version = 0     # bumped whenever the weights change (Adam step, load_state_dict): keys weight-derived caches
        self.ctx = _lib.context(self.device.index if self.device.index is not None else torch.cuda.current_device())

    # --- views
    def _view(self, buf, key):
        e = self.entries[key]
        return buf[e.off:e.off + e.numel]

    def W(self, key):
        return self._view(self.w, key)

    def G(self, key):
        return self._view(self.g, key)

    def W16(self, key):
        return self._view(self.w16, key)

    def W16T(self, key):
        e = self.entries[key]
        return self.w16t[e.dg_off:e.dg_off + e.numel]

    def W16L(self, key):
        return self._view(self.w16l, key)

    def W16TL(self, key):
        e = self.entries[key]
        return self.w16tl[e.dg_off:e.dg_off + e.numel]

    # --- reference-compatible (de)serialisation
    def state_dict(self):
        out = OrderedDict()
        for k, e in self.entries.items():
            out[k] = from_kernel_layout(e.kind, self.W(k).detach(), e.shape)
        return out

    def load_state_dict(self, sd, strict=True):
        missing = [k for k in self.entries if k not in sd]
        if strict and missing:
            raise KeyError("missing keys in state_dict: %s" % missing[:4])
        for k, e in self.entries.items():
            if k not in sd:
                continue
            t = sd[k].detach().to(device=self.device, dtype=torch.float32)
            if tuple(t.shape) != e.shape:
                raise ValueError("shape mismatch for %s: %s vs %s" % (k, tuple(t.shape), e.shape))
            self.W(k).copy_(to_kernel_layout(e.kind, t).reshape(-1))
        self.version += 1
        self.refresh_operands()

    def init_(self, seed):
        """The reference's init *laws* (src/trainers/init.py:8-12, lsps_nets.py:55-59): Conv*/ConvTranspose* weights
        N(0, 0.02); biases and Linear weights U(+-1/sqrt(fan_in)); en_mu/en_sigma N(0, 0.002)."""
        g = torch.Generator().manual_seed(seed)
        sd = OrderedDict()
        for k, e in self.entries.items():
            if e.law == "conv":
                t = torch.randn(e.shape, generator=g) * 0.02
            elif e.law == "small":
                t = torch.randn(e.shape, generator=g) * 0.002
            else:
                b = 1.0 / math.sqrt(e.fan_in)
                t = (torch.rand(e.shape, generator=g) * 2 - 1) * b
            sd[k] = t.float()
        self.load_state_dict(sd)

    def refresh_operands(self):
        """bf16 forward copy of everything + transposed dgrad copies of the 3x3 (de)conv weights."""
        if self.split:
            self.ctx.f32_split_bf16(self.w.data_ptr(), self.w16.data_ptr(), self.w16l.data_ptr(), self.size)
        else:
            self.ctx.f32_to_bf16(self.w.data_ptr(), self.w16.data_ptr(), self.size)
        self.refresh_dgrad_operands()

    def refresh_dgrad_operands(self):
        """[tap][ci][co] bf16 copies of every conv weight (and their remainders for a split store): ONE launch."""
        if self._pack_count:
            self.ctx.pack_dgrad_multi(self.w.data_ptr(), self.w16t.data_ptr(),
                                      self.w16tl.data_ptr() if self.split else None, self._pack_desc.data_ptr(),
                                      self._pack_count, self._pack_tiles)

    # --- optimiser
    def zero_grad(self):
        self.ctx.memset(self.gbuf.data_ptr(), 0, self.gbuf.numel() * 4)

    def _segments(self, active):
        """Maximal runs of consecutive active tensors with equal step counts -> [(first, last)] entry indices."""
        ents = list(self.entries.values())
        segs, i = [], 0
        while i < len(ents):
            e = ents[i]
            if active is not None and not active(e.key):
                i += 1
                continue
            j = i
            while j + 1 < len(ents) and ents[j + 1].step == e.step and (active is None or active(ents[j + 1].key)):
                j += 1
            segs.append((i, j))
            i = j + 1
        return segs

    def _hyper(self, step):
        b1, b2 = self.betas
        return self.lr / (1.0 - b1 ** step), 1.0 / math.sqrt(1.0 - b2 ** step)

    def adam_step(self, active=None, grad_scale=1.0, hyper=None):
        """torch.optim.Adam semantics incl. 'parameters without a gradient are skipped' (their step counter does not
        advance).  `active`: None = all, else a predicate key -> bool.  Contiguous active runs with equal step
        counts are fused into one launch over the flat buffer.
        hyper: optional (device float tensor [2*k], host list) pair -- CUDA-graph capture: the launches read their
        step-dependent factors from hyper[2*i:2*i+2]; returns the segment list so that `advance()` can replay it."""
        ents = list(self.entries.values())
        segs = self._segments(active)
        self.version += 1
        for si, (i, j) in enumerate(segs):
            lo, hi = ents[i].off, ents[j].off + _round_up(ents[j].numel)
            step = ents[i].step + 1
            for q in range(i, j + 1):
                ents[q].step = step
            hp = None
            if hyper is not None:
                hp = hyper[2 * si:].data_ptr()
            self.ctx.adam_ex(self.w[lo:].data_ptr(), self.g[lo:].data_ptr(), self.m[lo:].data_ptr(),
                             self.v[lo:].data_ptr(), self.w16[lo:].data_ptr(),
                             self.w16l[lo:].data_ptr() if self.split else None, hi - lo, float(self.lr), self.betas[0],
                             self.betas[1], self.eps, float(self.wd), step, float(grad_scale), hp)
        return segs

    def segments_consistent(self, segs):
        ents = list(self.entries.values())
        return all(ents[q].step == ents[i].step for i, j in segs for q in range(i, j + 1))

    def advance(self, segs, hyper_host):
        """Replay bookkeeping for a captured adam_step: bump the step counters of `segs` and write the factors the
        captured launches will read into hyper_host (pinned float tensor)."""
        ents = list(self.entries.values())
        self.version += 1
        for si, (i, j) in enumerate(segs):
            step = ents[i].step + 1
            for q in range(i, j + 1):
                ents[q].step = step
            a, b = self._hyper(step)
            hyper_host[2 * si], hyper_host[2 * si + 1] = a, b

    def opt_state(self):
        return {"m": self.m.clone(), "v": self.v.clone(), "steps": {k: e.step for k, e in self.entries.items()},
                "lr": self.lr}

    def load_opt_state(self, st):
        self.m.copy_(st["m"]); self.v.copy_(st["v"]); self.lr = st["lr"]
        for k, s in st["steps"].items():
            self.entries[k].step = s


class MultiStepLR:
    """torch.optim.lr_scheduler.MultiStepLR over a ParamStore (lsps_trainer.py:32-34): milestones are counted in
    scheduler steps, the drivers call step() every 1000 (pretrain / pose) or 100 (estimate) iterations."""

    def __init__(self, store, milestones, gamma):
        self.store, self.milestones, self.gamma, self.last_epoch = store, sorted(milestones), gamma, 0

    def step(self):
        self.last_epoch += 1
        k = sum(1 for m in self.milestones if m <= self.last_epoch)
        self.store.lr = self.store.base_lr * (self.gamma ** k)

    def get_lr(self):
        return [self.store.lr]

    get_last_lr = get_lr


class Optimizer:
    """Minimal stand-in for the reference's `*_opt` attributes (state_dict round trip + param_groups lr)."""

    def __init__(self, store):
        self.store = store

    @property
    def param_groups(self):
        return [{"lr": self.store.lr, "betas": self.store.betas, "weight_decay": self.store.wd}]

    def state_dict(self):
        return self.store.opt_state()

    def load_state_dict(self, st):
        self.store.load_opt_state(st)


# ------------------------------------------------------------------ parameter tables (same keys as the reference)
def gen_entries(p):
    """SharedResGen (lsps_nets.py:164-237)."""
    ch, out = p["ch"], []

    def conv(key, co, ci, k, kind):
        shape = (ci, co, k, k) if kind in ("deconv3", "head") else (co, ci, k, k)
        fan = (co if kind in ("deconv3", "head") else ci) * k * k
        out.append((key + ".weight", shape, kind, "conv", fan))
        out.append((key + ".bias", (co,), "bias", "bias", fan))

    resx = p.get("name") == "SharedResXGen"      # LeakyINSResNeXtBlock res blocks (lsps_nets.py:277-343)
    rk, rc = p.get("n_resnext_k", 1), p.get("n_resnext_c", 4)

    def res(prefix, c):
        if resx:
            conv(prefix + ".model.0", rk * c, c, 1, "conv1")
            gw = rk * c // rc
            out.append((prefix + ".model.3.weight", (rk * c, gw, 3, 3), "gconv3", "conv", gw * 9))
            out.append((prefix + ".model.3.bias", (rk * c,), "bias", "bias", gw * 9))
            conv(prefix + ".model.6", c, rk * c, 1, "conv1")
            return
        conv(prefix + ".model.0", c, c, 3, "conv3")
        conv(prefix + ".model.3", c, c, 3, "conv3")

    t = ch
    for dom, cin in (("A", p["input_dim_a"]), ("B", p["input_dim_b"])):
        e = "encode_%s" % dom
        conv("%s.0.model.0" % e, ch, cin, 7, "stem")
        t, idx = ch, 1
        for _ in range(1, p["n_enc_front_blk"]):
            conv("%s.%d.model.0" % (e, idx), 2 * t, t, 3, "conv3")
            t, idx = 2 * t, idx + 1
        for _ in range(p["n_enc_res_blk"]):
            res("%s.%d" % (e, idx), t)
            idx += 1
    for i in range(p["n_enc_shared_blk"]):
        res("enc_shared.%d" % i, t)
    for i in range(p["n_gen_shared_blk"]):
        res("dec_shared.%d" % i, t)
    for dom, cout in (("A", p["input_dim_a"]), ("B", p["input_dim_b"])):
        d = "decode_%s" % dom
        tt, idx = t, 0
        for _ in range(p["n_gen_res_blk"]):
            res("%s.%d" % (d, idx), tt)
            idx += 1
        for _ in range(1, p["n_gen_front_blk"]):
            conv("%s.%d.model.0" % (d, idx), tt // 2, tt, 3, "deconv3")
            tt, idx = tt // 2, idx + 1
        conv("%s.%d" % (d, idx), cout, tt, 1, "head")
    return out


def dis_entries(p):
    """SharedDis (lsps_nets.py:86-126)."""
    ch, out = p["ch"], []

    def conv(key, co, ci, k, kind):
        out.append((key + ".weight", (co, ci, k, k), kind, "conv", ci * k * k))
        out.append((key + ".bias", (co,), "bias", "bias", ci * k * k))

    t = ch
    for dom, cin in (("A", p["input_dim_a"]), ("B", p["input_dim_b"])):
        conv("model_%s.0.model.0" % dom, ch, cin, 7, "stem")
        t = ch
        for i in range(1, p["n_front_layer"]):
            conv("model_%s.%d.model.0" % (dom, i), 2 * t, t, 3, "conv3")
            t *= 2
    if p.get("n_expand_layer", 0):
        raise NotImplementedError("n_expand_layer > 0 is not used by the reference configs")
    for idx in range(p["n_shared_layer"]):
        conv("model_S.%d.model.0" % idx, 2 * t, t, 3, "conv3")
        t *= 2
    conv("D", 1, t, 1, "dhead")
    conv("Post", p["post_dim"], t, 2, "post")
    return out


def vae_entries(p):
    """poseVAE (lsps_nets.py:34-59)."""
    d, z, h = p["input_dim"], p["z_dim"], p["h_dim"]
    out = []
    for key, (o, i), law in (("en_fc1", (h, d), "linear"), ("en_mu", (z, h), "small"), ("en_sigma", (z, h), "small"),
                             ("de_fc1.model.0", (h, z), "linear"), ("de_fc2", (d, h), "linear")):
        out.append((key + ".weight", (o, i), "linear", law, i))
        out.append((key + ".bias", (o,), "bias", "small" if law == "small" else "bias", i))
    return out


def map_entries(p):
    """Mapping (lsps_nets.py:8-25): ConvTranspose2d(20 -> 4ch, k4 s1 p0) on the 1x1 pose latent, then three
    ConvTranspose2d(k4 s2 p1) up to (ch, 32, 32); LeakyReLU after the first three.  `output_dim` is informational in the
    reference too (only stored in an attribute): the four layers always produce 32x32."""
    ch, d, out = p["output_ch"], p["input_dim"], []
    for key, ci, co, kind in (("model.0.model.0", d, 4 * ch, "map0"), ("model.1.model.0", 4 * ch, 4 * ch, "deconv4"),
                              ("model.2.model.0", 4 * ch, 2 * ch, "deconv4"), ("model.3", 2 * ch, ch, "deconv4")):
        out.append((key + ".weight", (ci, co, 4, 4), kind, "conv", co * 16))
        out.append((key + ".bias", (co,), "bias", "bias", co * 16))
    return out
