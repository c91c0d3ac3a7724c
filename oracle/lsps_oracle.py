"""TEST INFRASTRUCTURE ONLY -- CPU oracle ("port") of the LSPS training-step hot path.

A functional fp32 PyTorch restatement of the reference algorithm (masabdi/LSPS),
written against a flat {state_dict key -> tensor} parameter store instead of the
reference's nn.Module zoo.  It is the checker for the CUDA path in `lsps_b200/`
and the `cpu_baseline` / `--impl reference` arm of bench.py.  Only tests/,
__graft_entry__.smoke() and bench.py may import it; the product never does.

Parity status: the reference ships no tests or golden vectors, so
this port is pinned against the *reference itself executed in the build container*
(`oracle/make_golden.py` via `oracle/ref_loader.py`): identical weights + identical
host RNG stream -> losses/outputs equal to float32 round-off; the resulting
vectors are committed under tests/golden/ and re-checked by `pytest -m "not gpu"`.

Reference anchors (paths relative to /root/reference):
  nets      src/trainers/lsps_nets.py:34-83 (poseVAE) :86-160 (SharedDis) :164-272 (SharedResGen)
  layers    src/trainers/common_net.py:32-40 (noise) :160-181 (LeakyINSResBlock) :221-268 (lrelu conv/deconv/linear)
  init      src/trainers/init.py:8-12 ; src/trainers/lsps_nets.py:55-59
  updates   src/trainers/lsps_trainer.py:55-74 (kl, vae_update) :76-141 (gen_update)
            :143-218 (dis_update) :220-262 (post_update) ; optimisers :26-34
The third-party arithmetic underneath (torch.nn.functional conv2d / conv_transpose2d /
instance_norm / leaky_relu / binary_cross_entropy, torch.optim.Adam) is torch 2.11 here;
semantics relied on are listed in SURVEY.md section 8c.
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

LRELU = 0.01          # nn.LeakyReLU() default slope (common_net.py:169,251)
IN_EPS = 1e-5         # nn.InstanceNorm2d default eps (common_net.py:168)


# ----------------------------------------------------------------------------------------
# parameter specs (shapes + init law), keyed exactly like the reference state_dicts
# ----------------------------------------------------------------------------------------
def gen_spec(p):
    """SharedResGen parameter table: lsps_nets.py:164-237; with name == "SharedResXGen" the res blocks are
    LeakyINSResNeXtBlock (common_net.py:111-132; lsps_nets.py:277-343): 1x1 (c -> k*c), grouped 3x3 (k*c, groups =
    cardinality), 1x1 (k*c -> c), an InstanceNorm after each."""
    ch, spec = p["ch"], OrderedDict()
    resx = p.get("name") == "SharedResXGen"
    rk, rc = p.get("n_resnext_k", 1), p.get("n_resnext_c", 4)

    def conv(key, co, ci, k, transposed=False, groups=1):
        cig = ci // groups
        spec[key + ".weight"] = ((ci, co, k, k) if transposed else (co, cig, k, k), "conv", cig * k * k if not transposed else co * k * k)
        spec[key + ".bias"] = ((co,), "bias", cig * k * k if not transposed else co * k * k)

    def res(prefix, c):
        if resx:
            conv(prefix + ".model.0", rk * c, c, 1)
            conv(prefix + ".model.3", rk * c, rk * c, 3, groups=rc)
            conv(prefix + ".model.6", c, rk * c, 1)
            return
        conv(prefix + ".model.0", c, c, 3)
        conv(prefix + ".model.3", c, c, 3)

    for dom, cin in (("A", p["input_dim_a"]), ("B", p["input_dim_b"])):
        e = "encode_%s" % dom
        conv("%s.0.model.0" % e, ch, cin, 7)
        t, idx = ch, 1
        for _ in range(1, p["n_enc_front_blk"]):
            conv("%s.%d.model.0" % (e, idx), 2 * t, t, 3)
            t, idx = 2 * t, idx + 1
        for _ in range(p["n_enc_res_blk"]):
            res("%s.%d" % (e, idx), t)
            idx += 1
    # module registration order in the reference: encode_A, encode_B, enc_shared, dec_shared, decode_A, decode_B
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
            conv("%s.%d.model.0" % (d, idx), tt // 2, tt, 3, transposed=True)
            tt, idx = tt // 2, idx + 1
        conv("%s.%d" % (d, idx), cout, tt, 1, transposed=True)
    # encode_A / encode_B interleave in registration order: A fully, then B (they are separate Sequentials)
    ordered = OrderedDict()
    for pre in ("encode_A", "encode_B", "enc_shared", "dec_shared", "decode_A", "decode_B"):
        for k, v in spec.items():
            if k.startswith(pre + "."):
                ordered[k] = v
    return ordered


def dis_spec(p):
    """SharedDis parameter table: lsps_nets.py:86-126 (n_expand_layer = 0 in both YAMLs)."""
    ch, spec = p["ch"], OrderedDict()

    def conv(key, co, ci, k):
        spec[key + ".weight"] = ((co, ci, k, k), "conv", ci * k * k)
        spec[key + ".bias"] = ((co,), "bias", ci * k * k)

    for dom, cin in (("A", p["input_dim_a"]), ("B", p["input_dim_b"])):
        conv("model_%s.0.model.0" % dom, ch, cin, 7)
        t = ch
        for i in range(1, p["n_front_layer"]):
            conv("model_%s.%d.model.0" % (dom, i), 2 * t, t, 3)
            t *= 2
    idx = 0
    for _ in range(p.get("n_expand_layer", 0)):
        conv("model_S.%d.model.0" % idx, 2 * t, t, 3)
        t, idx = 2 * t, idx + 1
    for _ in range(p["n_shared_layer"]):
        conv("model_S.%d.model.0" % idx, 2 * t, t, 3)
        t, idx = 2 * t, idx + 1
    conv("D", 1, t, 1)
    conv("Post", p["post_dim"], t, 2)
    return spec


def vae_spec(p):
    """poseVAE parameter table: lsps_nets.py:34-59."""
    d, z, h = p["input_dim"], p["z_dim"], p["h_dim"]
    spec = OrderedDict()
    for key, (o, i), law in (("en_fc1", (h, d), "linear"), ("en_mu", (z, h), "small"),
                             ("en_sigma", (z, h), "small"), ("de_fc1.model.0", (h, z), "linear"),
                             ("de_fc2", (d, h), "linear")):
        spec[key + ".weight"] = ((o, i), law, i)
        spec[key + ".bias"] = ((o,), "small" if law == "small" else "bias", i)
    return spec


def map_spec(p):
    """Mapping parameter table: lsps_nets.py:8-25 (ConvTranspose2d weights are (Cin, Cout, 4, 4))."""
    ch, d, spec = p["output_ch"], p["input_dim"], OrderedDict()
    for key, ci, co in (("model.0.model.0", d, 4 * ch), ("model.1.model.0", 4 * ch, 4 * ch),
                        ("model.2.model.0", 4 * ch, 2 * ch), ("model.3", 2 * ch, ch)):
        spec[key + ".weight"] = ((ci, co, 4, 4), "conv", co * 16)
        spec[key + ".bias"] = ((co,), "bias", co * 16)
    return spec


def init_params(spec, seed):
    """Deterministic init with the reference's *laws* (not its RNG stream):
    Conv*/ConvTranspose* weights ~ N(0, 0.02) (init.py:8-12); biases and Linear weights
    keep torch's default U(-1/sqrt(fan_in), 1/sqrt(fan_in)); en_mu/en_sigma ~ N(0, 0.002)
    (lsps_nets.py:55-59)."""
    g = torch.Generator().manual_seed(seed)
    out = OrderedDict()
    for k, (shape, law, fan_in) in spec.items():
        if law == "conv":
            t = torch.randn(shape, generator=g) * 0.02
        elif law == "small":
            t = torch.randn(shape, generator=g) * 0.002
        else:
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        out[k] = t.float().contiguous()
    return out


# ----------------------------------------------------------------------------------------
# functional layers
# ----------------------------------------------------------------------------------------
def _lrelu(x):
    return F.leaky_relu(x, LRELU)


def _conv_lrelu(P, key, x, stride, pad):
    return _lrelu(F.conv2d(x, P[key + ".weight"], P[key + ".bias"], stride=stride, padding=pad))


def _res_block(P, key, x):
    """x + IN(conv(lrelu(IN(conv(x)))))  -- common_net.py:160-181.  A block that owns a `.model.6` conv is a
    LeakyINSResNeXtBlock (common_net.py:111-132): 1x1 -> IN -> lrelu -> grouped 3x3 -> IN -> lrelu -> 1x1 -> IN, + x."""
    if key + ".model.6.weight" in P:
        w3 = P[key + ".model.3.weight"]
        h = F.conv2d(x, P[key + ".model.0.weight"], P[key + ".model.0.bias"])
        h = _lrelu(F.instance_norm(h, eps=IN_EPS))
        h = F.conv2d(h, w3, P[key + ".model.3.bias"], padding=1, groups=w3.shape[0] // w3.shape[1])
        h = _lrelu(F.instance_norm(h, eps=IN_EPS))
        h = F.conv2d(h, P[key + ".model.6.weight"], P[key + ".model.6.bias"])
        return x + F.instance_norm(h, eps=IN_EPS)
    h = F.conv2d(x, P[key + ".model.0.weight"], P[key + ".model.0.bias"], padding=1)
    h = _lrelu(F.instance_norm(h, eps=IN_EPS))
    h = F.conv2d(h, P[key + ".model.3.weight"], P[key + ".model.3.bias"], padding=1)
    return x + F.instance_norm(h, eps=IN_EPS)


class Gen:
    """Functional SharedResGen over a parameter dict (lsps_nets.py:239-272)."""

    def __init__(self, hp_gen, params):
        self.p, self.P = hp_gen, params
        self.training = True   # the drivers never put gen in eval mode

    def _encode_front(self, dom, x):
        p, e = self.p, "encode_%s" % dom
        x = _conv_lrelu(self.P, "%s.0.model.0" % e, x, 1, 3)
        idx = 1
        for _ in range(1, p["n_enc_front_blk"]):
            x = _conv_lrelu(self.P, "%s.%d.model.0" % (e, idx), x, 2, 1)
            idx += 1
        for _ in range(p["n_enc_res_blk"]):
            x = _res_block(self.P, "%s.%d" % (e, idx), x)
            idx += 1
        return x

    def _enc_shared(self, x):
        for i in range(self.p["n_enc_shared_blk"]):
            x = _res_block(self.P, "enc_shared.%d" % i, x)
        if self.training:                               # GaussianNoiseLayer, host RNG (common_net.py:36-40)
            x = x + torch.randn(x.size())
        return x

    def _dec_shared(self, x):
        for i in range(self.p["n_gen_shared_blk"]):
            x = _res_block(self.P, "dec_shared.%d" % i, x)
        return x

    def _decode_back(self, dom, x):
        p, d = self.p, "decode_%s" % dom
        idx = 0
        for _ in range(p["n_gen_res_blk"]):
            x = _res_block(self.P, "%s.%d" % (d, idx), x)
            idx += 1
        for _ in range(1, p["n_gen_front_blk"]):
            x = _lrelu(F.conv_transpose2d(x, self.P["%s.%d.model.0.weight" % (d, idx)],
                                          self.P["%s.%d.model.0.bias" % (d, idx)],
                                          stride=2, padding=1, output_padding=1))
            idx += 1
        x = F.conv_transpose2d(x, self.P["%s.%d.weight" % (d, idx)], self.P["%s.%d.bias" % (d, idx)])
        return torch.tanh(x)

    def decode(self, z):
        """lsps_nets.py:239-243: both decoders on an externally supplied latent (no noise layer)."""
        out = self._dec_shared(z)
        return self._decode_back("A", out), self._decode_back("B", out)

    def forward(self, xa, xb):
        out = torch.cat((self._encode_front("A", xa), self._encode_front("B", xb)), 0)
        shared = self._enc_shared(out)
        out = self._dec_shared(shared)
        oa, ob = self._decode_back("A", out), self._decode_back("B", out)
        n = xa.size(0)
        x_aa, x_ba = torch.split(oa, n, 0)
        x_ab, x_bb = torch.split(ob, n, 0)
        return x_aa, x_ba, x_ab, x_bb, shared

    def forward_a2b(self, xa):
        shared = self._enc_shared(self._encode_front("A", xa))
        return self._decode_back("B", self._dec_shared(shared)), shared

    def forward_b2a(self, xb):
        shared = self._enc_shared(self._encode_front("B", xb))
        return self._decode_back("A", self._dec_shared(shared)), shared


class Dis:
    """Functional SharedDis (lsps_nets.py:128-160)."""

    def __init__(self, hp_dis, params):
        self.p, self.P = hp_dis, params

    def front(self, dom, x):
        x = _conv_lrelu(self.P, "model_%s.0.model.0" % dom, x, 2, 3)
        for i in range(1, self.p["n_front_layer"]):
            x = _conv_lrelu(self.P, "model_%s.%d.model.0" % (dom, i), x, 2, 1)
        return x

    def trunk(self, x):
        idx = 0
        for _ in range(self.p.get("n_expand_layer", 0)):
            x = _conv_lrelu(self.P, "model_S.%d.model.0" % idx, x, 1, 1)
            idx += 1
        for _ in range(self.p["n_shared_layer"]):
            x = _conv_lrelu(self.P, "model_S.%d.model.0" % idx, x, 2, 1)
            idx += 1
        return x

    def regress(self, dom, x):
        f = self.trunk(self.front(dom, x))
        return F.conv2d(f, self.P["Post.weight"], self.P["Post.bias"]).squeeze()

    def feats(self, x_aa, x_ba, x_ab, x_bb):
        f = torch.cat((self.front("A", torch.cat((x_aa, x_ba), 0)),
                       self.front("B", torch.cat((x_ab, x_bb), 0))), 0)
        f = self.trunk(f)
        return torch.split(f, f.size(0) // 4, 0)

    def forward(self, xa, xb):
        f = self.trunk(torch.cat((self.front("A", xa), self.front("B", xb)), 0))
        out = F.conv2d(f, self.P["D.weight"], self.P["D.bias"])
        fa, fb = torch.split(f, f.size(0) // 2, 0)
        oa, ob = torch.split(out, out.size(0) // 2, 0)
        return oa.reshape(-1), ob.reshape(-1), fa, fb


class PoseVAE:
    """Functional poseVAE (lsps_nets.py:68-83)."""

    def __init__(self, hp_vae, params):
        self.p, self.P = hp_vae, params

    def encode(self, y):
        P = self.P
        h = _lrelu(F.linear(y, P["en_fc1.weight"], P["en_fc1.bias"]))
        mu = F.linear(h, P["en_mu.weight"], P["en_mu.bias"])
        sd = F.softplus(F.linear(h, P["en_sigma.weight"], P["en_sigma.bias"]))
        noise = torch.normal(torch.zeros(mu.size()), std=0.05)      # host RNG (lsps_nets.py:77)
        return mu + sd * noise, mu, sd

    def decode(self, z):
        P = self.P
        h = _lrelu(F.linear(z, P["de_fc1.model.0.weight"], P["de_fc1.model.0.bias"]))
        return F.linear(h, P["de_fc2.weight"], P["de_fc2.bias"])

    def forward(self, y):
        z, mu, sd = self.encode(y)
        return self.decode(z), z, mu, sd


class Mapping:
    """Functional Mapping net (lsps_nets.py:8-31): pose latent (n, 20) -> (n, 256, 32, 32)."""

    def __init__(self, hp_map, params):
        self.p, self.P = hp_map, params

    def forward(self, x):
        P = self.P
        x = x.unsqueeze(2).unsqueeze(3)
        x = _lrelu(F.conv_transpose2d(x, P["model.0.model.0.weight"], P["model.0.model.0.bias"], stride=1, padding=0))
        x = _lrelu(F.conv_transpose2d(x, P["model.1.model.0.weight"], P["model.1.model.0.bias"], stride=2, padding=1))
        x = _lrelu(F.conv_transpose2d(x, P["model.2.model.0.weight"], P["model.2.model.0.bias"], stride=2, padding=1))
        return F.conv_transpose2d(x, P["model.3.weight"], P["model.3.bias"], stride=2, padding=1)


# ----------------------------------------------------------------------------------------
# trainer
# ----------------------------------------------------------------------------------------
def _bce_logits_as_reference(logits, target_value):
    """sigmoid then binary_cross_entropy with mean reduction (lsps_trainer.py:107-112,179-192)."""
    prob = torch.sigmoid(logits)
    return F.binary_cross_entropy(prob, torch.full_like(prob, target_value))


class OracleTrainer:
    """CPU restatement of LSPSTrainer.  The Mapping net exists only when `hp["train_map"]` is set (the reference always
    builds it and hands its parameters to gen_opt, lsps_trainer.py:24-28, but Adam skips parameters without a gradient,
    so with train_map False it never changes)."""

    def __init__(self, hp, seed=0, params=None):
        self.hp = hp
        if params is None:
            params = {"gen": init_params(gen_spec(hp["gen"]), seed + 1),
                      "dis": init_params(dis_spec(hp["dis"]), seed + 2),
                      "vae": init_params(vae_spec(hp["vae"]), seed + 3)}
            if hp.get("train_map"):
                params["map"] = init_params(map_spec(hp["map"]), seed + 4)
        self.params = {net: OrderedDict((k, v.clone().requires_grad_(True)) for k, v in d.items())
                       for net, d in params.items()}
        self.gen = Gen(hp["gen"], self.params["gen"])
        self.dis = Dis(hp["dis"], self.params["dis"])
        self.vae = PoseVAE(hp["vae"], self.params["vae"])
        self.map = Mapping(hp["map"], self.params["map"]) if "map" in self.params else None
        lr = hp["lr"]
        adam = torch.optim.Adam
        self.dis_opt = adam(list(self.params["dis"].values()), lr=lr, betas=(0.5, 0.999), weight_decay=1e-4)
        self.gen_opt = adam(list(self.params["gen"].values()) + list(self.params.get("map", {}).values()), lr=lr, betas=(0.5, 0.999), weight_decay=1e-4)
        self.vae_opt = adam(list(self.params["vae"].values()), lr=lr * 10.0, betas=(0.5, 0.999), weight_decay=1e-3)
        msl = torch.optim.lr_scheduler.MultiStepLR
        self.dis_sch = msl(self.dis_opt, milestones=[200, 300, 400, 450], gamma=0.5)
        self.gen_sch = msl(self.gen_opt, milestones=[200, 300, 400, 450], gamma=0.5)
        self.vae_sch = msl(self.vae_opt, milestones=[125, 175], gamma=0.1)

    def state_dict(self, net):
        return OrderedDict((k, v.detach().clone()) for k, v in self.params[net].items())

    def _zero(self, net):
        for v in self.params.get(net, {}).values():
            v.grad = None

    def _map_decode(self, la, lb):
        """lsps_trainer.py:86-93 / :148-155: pose labels -> vae.encode -> Mapping -> both decoders; domain A keeps the
        first half of decode_A, domain B the second half of decode_B.  Also returns nothing else; the mapped latent is
        recomputed by the caller when it needs it."""
        z = self.map.forward(self.vae.encode(torch.cat((la, lb), 0))[0])
        self._z_pose2depth = z
        dec_a, dec_b = self.gen.decode(z)
        n = dec_a.size(0) // 2
        return dec_a[:n], dec_b[n:]

    @staticmethod
    def _kl(mu, sd=None):
        if sd is None:
            return torch.mean(mu * mu)
        return (mu * mu + sd * sd - torch.log(sd * sd)).sum() / mu.size(0)

    # --- lsps_trainer.py:62-74
    def vae_update(self, y, hp=None):
        hp = hp or self.hp
        self._zero("vae")
        dec, z, mu, sd = self.vae.forward(y)
        total = hp["kl_loss_vae"] * self._kl(mu, sd) + hp["ll_loss_vae"] * F.l1_loss(dec, y)
        total.backward()
        self.vae_opt.step()
        self.vae_total_loss = total.item()
        return dec.detach()

    # --- lsps_trainer.py:143-218 (feat_mat branch, train_map False)
    def dis_update(self, ia, la, ib, lb, com_a=None, com_b=None, hp=None, feat_mat=True):
        hp = hp or self.hp
        self._zero("dis")
        x_aa, x_ba, x_ab, x_bb, _ = self.gen.forward(ia, ib)
        if hp["train_map"]:          # :147-158
            dec_a, dec_b = self._map_decode(la, lb)
            da, db, ndiv = torch.cat((ia, x_ba, x_aa, dec_a), 0), torch.cat((ib, x_ab, x_bb, dec_b), 0), 4
        elif feat_mat:
            da, db, ndiv = torch.cat((ia, x_ba, x_aa), 0), torch.cat((ib, x_ab, x_bb), 0), 3
        else:
            da, db, ndiv = torch.cat((ia, x_ba), 0), torch.cat((ib, x_ab), 0), 2
        ra, rb, fa, fb = self.dis.forward(da, db)
        feat = 0.0
        if feat_mat:
            fas, fbs = torch.split(fa, fa.size(0) // ndiv, 0), torch.split(fb, fb.size(0) // ndiv, 0)
            feat = (fbs[1] - fas[2]).abs().mean() + (fas[1] - fbs[2]).abs().mean()
        la_, lb_ = torch.split(ra, ra.size(0) // ndiv, 0), torch.split(rb, rb.size(0) // ndiv, 0)
        ad = (_bce_logits_as_reference(la_[0], 1.0) + _bce_logits_as_reference(la_[1], 0.0) +
              _bce_logits_as_reference(lb_[0], 1.0) + _bce_logits_as_reference(lb_[1], 0.0))
        if hp["train_map"]:          # :201-204
            ad = ad + _bce_logits_as_reference(la_[3], 0.0) + _bce_logits_as_reference(lb_[3], 0.0)
        with torch.no_grad():   # helpers.py:20-32
            self.dis_true_acc = 0.5 * ((torch.sigmoid(la_[0]) >= 0.5).float().mean().item() +
                                       (torch.sigmoid(lb_[0]) >= 0.5).float().mean().item())
            self.dis_fake_acc = 0.5 * ((torch.sigmoid(la_[1]) <= 0.5).float().mean().item() +
                                       (torch.sigmoid(lb_[1]) <= 0.5).float().mean().item())
        loss = hp["gan_w"] * ad + hp["feature_w"] * feat
        loss.backward()
        self._zero("gen")            # the reference clears these in gen_update before use (:77)
        self._zero("map")            # ... and these at :85 ; vae grads are cleared by vae.zero_grad() (:63)
        self._zero("vae")
        self.dis_opt.step()
        self.dis_ad_loss = ad.item()
        self.dis_feat_loss = float(feat.detach()) if torch.is_tensor(feat) else float(feat)
        self.dis_loss = loss.item()

    # --- lsps_trainer.py:76-141 (train_map False)
    def gen_update(self, ia, la, ib, lb, hp=None):
        hp = hp or self.hp
        self._zero("gen")
        x_aa, x_ba, x_ab, x_bb, shared = self.gen.forward(ia, ib)
        x_bab, shared_bab = self.gen.forward_a2b(x_ba)
        x_aba, shared_aba = self.gen.forward_b2a(x_ab)
        map_z, map_ll = 0.0, 0.0
        if hp["train_map"]:          # :84-99
            self._zero("map")
            dec_a, dec_b = self._map_decode(la, lb)
            da, db = torch.cat((x_ba, dec_a), 0), torch.cat((x_ab, dec_b), 0)
            map_z = ((shared - self._z_pose2depth) ** 2).mean()
            map_ll = F.l1_loss(dec_a, ia) + F.l1_loss(dec_b, ib)
        else:
            da, db, dec_a, dec_b = x_ba, x_ab, x_ba, x_ab
        oa, ob, _, _ = self.dis.forward(da, db)
        ad = _bce_logits_as_reference(oa, 1.0) + _bce_logits_as_reference(ob, 1.0)
        enc, enc_bab, enc_aba = self._kl(shared), self._kl(shared_bab), self._kl(shared_aba)
        ll_a, ll_b = F.l1_loss(x_aa, ia), F.l1_loss(x_bb, ib)
        ll_aba, ll_bab = F.l1_loss(x_aba, ia), F.l1_loss(x_bab, ib)
        total = (hp["gan_w"] * ad + hp["ll_direct_link_w"] * (ll_a + ll_b) +
                 hp["ll_cycle_link_w"] * (ll_aba + ll_bab) + hp["kl_direct_link_w"] * (enc + enc) +
                 hp["kl_cycle_link_w"] * (enc_bab + enc_aba) +
                 hp["ll_map_z_w"] * map_z + hp["ll_map_w"] * map_ll)
        total.backward()
        self._zero("vae")            # vae.encode sits inside the graph (:88); vae_opt never sees these gradients
        self.gen_opt.step()
        self.gen_enc_loss, self.gen_enc_loss2 = enc.item(), (enc_aba + enc_bab).item()
        self.gen_ad_loss = ad.item()
        self.gen_ll_loss, self.gen_ll_loss2 = (ll_a + ll_b).item(), (ll_bab + ll_aba).item()
        if hp["train_map"]:
            self.gen_map_loss, self.gen_map_loss2 = map_z.item(), map_ll.item()
        self.gen_total_loss = total.item()
        return tuple(t.detach() for t in (x_aa, x_ba, x_ab, x_bb, x_aba, x_bab, dec_a, dec_b))

    # --- lsps_trainer.py:220-262
    def post_update(self, ia, la, ib, lb, com_a=None, com_b=None, mode=3, hp=None):
        hp = hp or self.hp
        self._zero("dis")
        x_aa, x_ba, x_ab, x_bb = ia, ia, ib, ib
        feat, reg = 0.0, 0.0
        if mode == 0:
            reg = ((self.dis.regress("A", ia) - self.vae.encode(la)[0]) ** 2).mean()
        elif mode == 1:
            reg = ((self.dis.regress("B", ib) - self.vae.encode(lb)[0]) ** 2).mean()
        else:
            x_aa, x_ba, x_ab, x_bb, _ = self.gen.forward(ia[0:4], ib[0:4])
            f_aa, f_ba, f_ab, f_bb = self.dis.feats(x_aa, x_ba, x_ab, x_bb)
            feat = (f_ab - f_aa).abs().mean() + (f_ba - f_bb).abs().mean()
            reg = ((self.dis.regress("A", ia) - self.vae.encode(la)[0]) ** 2).mean()
            if mode == 4:
                reg = reg + ((self.dis.regress("B", ib) - self.vae.encode(lb)[0]) ** 2).mean()
        total = hp["reg_w"] * reg + hp["feature_w_reg"] * feat
        total.backward()
        self._zero("gen")
        self._zero("vae")
        self.dis_opt.step()
        self.dis_reg_loss = float(reg.detach()) if torch.is_tensor(reg) else float(reg)
        self.dis_total_loss = total.item()
        return tuple(t.detach() for t in (x_aa, x_ba, x_ab, x_bb, x_aa, x_bb, x_aa, x_bb))


# ----------------------------------------------------------------------------------------
# synthetic inputs: shared by the oracle, the tests and bench.py
# ----------------------------------------------------------------------------------------
def synthetic_batch(batch, label_dim, generator, kind="uniform"):
    """(ia, ib, la, lb): depth crops in [-1,1] (background +1 for kind='hand'), labels ~ 0.3*N(0,1).

    Item contract: src/data/dataset_hand2.py:352,366 (img (1,128,128) float32, label (J*3,))."""
    def img():
        if kind == "uniform":
            return torch.rand(batch, 1, 128, 128, generator=generator) * 2 - 1
        x = torch.ones(batch, 1, 128, 128)
        yy, xx = torch.meshgrid(torch.arange(128.0), torch.arange(128.0), indexing="ij")
        for i in range(batch):
            ax = 30 + 20 * torch.rand(2, generator=generator)
            m = ((yy - 63.5) / ax[0]) ** 2 + ((xx - 63.5) / ax[1]) ** 2 <= 1.0
            v = (torch.randn(128, 128, generator=generator) * 0.35).clamp(-1, 1)
            x[i, 0][m] = v[m]
        return x
    ia, ib = img(), img()
    la = torch.randn(batch, label_dim, generator=generator) * 0.3
    lb = torch.randn(batch, label_dim, generator=generator) * 0.3
    return ia, ib, la, lb


# ----------------------------------------------------------------------------------------
# evaluation metrics (src/depth_train.py:229-237; src/utils/handpose_evaluation.py:92-97,197-203)
# ----------------------------------------------------------------------------------------
NYU_RESTRICTED_JOINTS = (0, 3, 6, 9, 12, 15, 18, 21, 24, 25, 27, 30, 31, 32)


def evaluation_metrics(gt, pred, cube0, com=None, restricted=None, dist=40.0):
    """numpy restatement: gt/pred (n, J*3) normalised joints -> (mean error mm, % frames with max joint error <= dist)."""
    import numpy as np
    gt = np.asarray(gt, np.float64).reshape(gt.shape[0], -1, 3)
    pred = np.asarray(pred, np.float64).reshape(pred.shape[0], -1, 3)
    if restricted is not None:
        gt, pred = gt[:, list(restricted)], pred[:, list(restricted)]
    scale = np.asarray(cube0, np.float64).reshape(3) / 2.0
    com = np.zeros((gt.shape[0], 1, 3)) if com is None else np.asarray(com, np.float64).reshape(-1, 1, 3)
    g3, p3 = gt * scale + com, pred * scale + com
    e = np.sqrt(np.square(g3 - p3).sum(axis=2))
    return float(np.nanmean(np.nanmean(e, axis=1))), float(100.0 * (np.nanmax(e, axis=1) <= dist).sum() / len(e))
