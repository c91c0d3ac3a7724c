"""LSPSTrainerB200 -- drop-in for the reference's `LSPSTrainer` (src/trainers/lsps_trainer.py:15-347).

Same constructor argument (the YAML `hyperparameters` dict), same update methods and return values, same loss /
accuracy attribute names (read by common.py:71-80 `write_loss`), same state_dict keys -- but every update is a
static kernel schedule on sm_100a (engine.py) and each update does ONE device->host read for all of its scalars.

Multi-GPU: one process per GPU (torch.distributed, NCCL).  Each rank calls the update with ITS shard of the
batch; the only exchange is one sum-allreduce of the flat gradient buffer (loss sums ride in its tail) per update.
Every loss is normalised by the GLOBAL element count so that the summed gradient equals the single-GPU one.
"""
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from .engine import Ops, Generator, Discriminator, PoseVAE, Mapping, SLOPE
from .sharding import shard_rows, feature_sources, allreduce_sum_, world_rank
from .params import ParamStore, MultiStepLR, Optimizer, gen_entries, dis_entries, vae_entries, map_entries

_LATENT = 256 * 32 * 32
_NPIX = 128 * 128


def _world():
    return world_rank()


def _img(t, dev):
    """(B,1,128,128) any device -> contiguous fp32 [B,128,128] on dev."""
    t = t.detach()
    return t.reshape(t.shape[0], t.shape[-2], t.shape[-1]).to(device=dev, dtype=torch.float32).contiguous()


class _JointSchedule:
    """gen_sch when the Mapping store exists: steps both MultiStepLRs (one optimiser in the reference)."""

    def __init__(self, *schedules):
        self.schedules = schedules

    def step(self):
        for s in self.schedules:
            s.step()

    def get_lr(self):
        return self.schedules[0].get_lr()

    get_last_lr = get_lr


class LSPSTrainerB200(object):
    def __init__(self, hyperparameters, device=None, seed=0, noise="host", graphs=False, precision=None):
        """precision: "mixed" (default) runs the discriminator stack on the split-bf16 ("bf16x3") kernels -- ~1 % of the
        step's FLOPs, ~80 % of the adversarial-loss deviation from the fp32 reference when run in plain bf16
        (profiles/r02_precision_ablation.json); "bf16" runs every conv with plain bf16 operands.  LSPS_PRECISION
        overrides the default."""
        hp = hyperparameters
        self.precision = precision or os.environ.get("LSPS_PRECISION", "mixed")
        assert self.precision in ("mixed", "bf16"), self.precision
        if device is None:
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index)
        self.gpu = self.device.index
        self.hp = hp
        self.noise_mode = noise  # "host": reference RNG stream (CPU torch.randn, same draw order); "device": Philox
        self.graphs = graphs     # CUDA-graph replay of post_update / vae_update (device-noise mode only)
        self._graphs = {}
        torch.cuda.set_device(self.device)
        self.ops = Ops(self.device)
        lr = hp["lr"]
        # lsps_trainer.py:26-31 -- Adam betas (0.5, 0.999); weight decay 1e-4 (dis, gen) / 1e-3 (vae); vae lr = 10*lr
        self.gen_store = ParamStore(gen_entries(hp["gen"]), self.device, lr, 1e-4)
        self.dis_store = ParamStore(dis_entries(hp["dis"]), self.device, lr, 1e-4, split=self.precision == "mixed")
        self.vae_store = ParamStore(vae_entries(hp["vae"]), self.device, lr * 10.0, 1e-3)
        self.gen_store.init_(seed + 1)
        self.dis_store.init_(seed + 2)
        self.vae_store.init_(seed + 3)
        self.gen = Generator(self.ops, self.gen_store, hp["gen"])
        self.dis = Discriminator(self.ops, self.dis_store, hp["dis"])
        self.vae = PoseVAE(self.ops, self.vae_store, hp["vae"], self._vae_noise)
        self.gen.state_dict, self.gen.load_state_dict = self.gen_store.state_dict, self.gen_store.load_state_dict
        # The reference always builds Mapping and hands its parameters to gen_opt (lsps_trainer.py:24-28); without
        # train_map they never receive a gradient, so Adam never touches them -- the 27.6 M-parameter store (and its
        # moments) is only allocated when the train_map branches are on.  Same lr / weight decay as the generator.
        self.map_store = self.map = None
        if hp.get("train_map", False):
            self.map_store = ParamStore(map_entries(hp["map"]), self.device, lr, 1e-4)
            self.map_store.init_(seed + 4)
            self.map = Mapping(self.ops, self.map_store, hp["map"])
        self._vae_noise_groups = 1
        self._vae_fused = os.environ.get("LSPS_NO_VAE_FUSED", "0") != "1"
        self.dis_opt, self.gen_opt, self.vae_opt = Optimizer(self.dis_store), Optimizer(self.gen_store), Optimizer(self.vae_store)
        # lsps_trainer.py:32-34
        self.dis_sch = MultiStepLR(self.dis_store, [200, 300, 400, 450], 0.5)
        self.gen_sch = MultiStepLR(self.gen_store, [200, 300, 400, 450], 0.5)
        if self.map_store is not None:      # map parameters sit in gen_opt: one schedule drives both stores
            self.gen_sch = _JointSchedule(self.gen_sch, MultiStepLR(self.map_store, [200, 300, 400, 450], 0.5))
        self.vae_sch = MultiStepLR(self.vae_store, [125, 175], 0.1)
        self._scratch = torch.zeros(8, dtype=torch.float32, device=self.device)
        self.last_outputs = None
        # generator front (encoders + enc_shared, deterministic) of the current step's images: computed by dis_update,
        # reused by gen_update when images and generator weights are unchanged (LSPS_NO_FRONT_CACHE=1 disables)
        self._front = None
        self._front_cache_on = os.environ.get("LSPS_NO_FRONT_CACHE", "0") != "1"
        self._comm = torch.cuda.Stream(device=self.device)
        self._early = {}
        self.dis.on_tail_wgrad = lambda key: self._allreduce_early(self.dis_store, key)
        # device-noise mode: Philox counters of the in-kernel draws; every rank gets its own stream (its samples differ)
        world, rank = _world()
        self._rng_seed = (int(seed) * 7919 + 1000003 * rank + 12345) & 0x7FFFFFFFFFFFFFFF
        self._rng_offset = 0
        if noise == "device":
            torch.cuda.manual_seed(self._rng_seed & 0x7FFFFFFF)     # the torch.randn draws that remain (pose-VAE, graphs)

    # ------------------------------------------------------------------ plumbing
    def cuda(self, gpu=None):
        if gpu is not None and gpu != self.gpu:
            raise RuntimeError("LSPSTrainerB200 lives on cuda:%d; construct it with device=%d" % (self.gpu, gpu))
        return self

    def _latent_noise(self, n, groups=1, shard=True):
        """GaussianNoiseLayer draw (common_net.py:36-40): the reference draws N(0,1) of shape (n,256,32,32) on the HOST.
        "host" mode reproduces that stream (same global draw on every rank, each rank keeps the rows of its samples:
        the global batch is `groups` blocks of world*n/groups rows, rank r owns the r-th slice of every block)."""
        world, rank = _world()
        if self.noise_mode == "host":
            if world == 1 or not shard:
                t = torch.randn(n, 256, 32, 32)
            else:
                t = shard_rows(torch.randn(n * world, 256, 32, 32), groups, world, rank)
            return t.to(self.device).permute(0, 2, 3, 1).contiguous()
        if torch.cuda.is_current_stream_capturing():
            return torch.randn(n, 32, 32, 256, device=self.device)   # graph-safe generator state; a baked counter is not
        self._rng_offset += 2
        return ("philox", self._rng_seed, self._rng_offset)

    def _vae_noise(self, shape):
        """poseVAE.encode draw (lsps_nets.py:77): torch.normal(zeros, std=0.05) on the host."""
        world, rank = _world()
        if self.noise_mode == "host":
            if world == 1:
                return torch.normal(torch.zeros(shape), std=0.05).to(self.device)
            t = torch.normal(torch.zeros((shape[0] * world,) + tuple(shape[1:])), std=0.05)
            return shard_rows(t, self._vae_noise_groups, world, rank).to(self.device)
        return torch.randn(shape, device=self.device) * 0.05

    def _allreduce(self, store):
        """The update's ONE exchange: sum-allreduce of the flat gradient buffer (+ loss sums in its tail).  When a slice
        of it was already started on the communication stream (_allreduce_early), only the rest goes now and the
        compute stream then waits for the early part."""
        world, _ = _world()
        if world <= 1:
            return
        early = self._early.pop(id(store), None)
        if early is None:
            allreduce_sum_(store.gbuf)
            return
        lo, hi, ev = early
        if lo > 0:
            dist.all_reduce(store.gbuf[:lo], op=dist.ReduceOp.SUM)
        dist.all_reduce(store.gbuf[hi:], op=dist.ReduceOp.SUM)
        torch.cuda.current_stream().wait_event(ev)

    def _allreduce_early(self, store, key):
        """Starts the allreduce of one conv's gradient slice (weight + bias) on the communication stream as soon as the
        kernels that produce it have been enqueued, so that it overlaps the rest of the backward pass."""
        world, _ = _world()
        if world <= 1 or torch.cuda.is_current_stream_capturing() or os.environ.get("LSPS_NO_EARLY_AR", "0") == "1":
            return
        from .params import _round_up
        ew, eb = store.entries[key + ".weight"], store.entries[key + ".bias"]
        lo, hi = ew.off, eb.off + _round_up(eb.numel)
        assert hi > lo and eb.off == ew.off + _round_up(ew.numel)
        self._comm.wait_event(self.ops.side_event())
        with torch.cuda.stream(self._comm):
            dist.all_reduce(store.gbuf[lo:hi], op=dist.ReduceOp.SUM)
            done = torch.cuda.Event()
            done.record(self._comm)
        self._early[id(store)] = (lo, hi, done)

    def _p(self, t):
        return t.data_ptr()

    def _front_key(self, ia, ib):
        return (ia.data_ptr(), ib.data_ptr(), ia._version, ib._version, tuple(ia.shape), tuple(ib.shape),
                self.gen_store.version)

    def _gen_front(self, ia, ib, keep):
        """front_fwd of (ia, ib), from the per-step cache when it holds exactly these tensors and weights.  keep: leave
        the result in the cache for the next update of this step (dis_update) or drop it after use (gen_update)."""
        key = self._front_key(ia, ib)
        hit = self._front is not None and self._front[0] == key
        front = self._front[1] if hit else self.gen.front_fwd(ia, ib)
        self._front = (key, front, ia, ib) if (keep and self._front_cache_on) else None
        return front

    def _map_decode(self, labels_a, labels_b, hp, save_map=None, save_dec=None):
        """lsps_trainer.py:86-93 / :148-155: cat(labels) -> vae.encode -> Mapping -> gen.decode; decode_A keeps the
        domain-a half, decode_B the domain-b half.  Returns (z_pose2depth, decode_A, decode_B)."""
        if self.map is None:
            raise RuntimeError("train_map=True needs a trainer constructed with hyperparameters['train_map'] = True")
        la = labels_a.detach().to(device=self.device, dtype=torch.float32)
        lb = labels_b.detach().to(device=self.device, dtype=torch.float32)
        self._vae_noise_groups = 2          # rows are (a-block | b-block): shard the host draw per block
        try:
            e = self.vae.encode(torch.cat((la, lb), 0))[0]
        finally:
            self._vae_noise_groups = 1
        z = self.map.forward(e, save_map)
        dec_a, dec_b = self.gen.decode_fwd(z, save_dec)
        return z, dec_a, dec_b

    # ------------------------------------------------------------------ vae_update (lsps_trainer.py:62-74)
    def vae_update(self, y, hyperparameters=None):
        hp = hyperparameters or self.hp
        S = self.vae_store
        world, _ = _world()
        y = y.detach().to(device=self.device, dtype=torch.float32).contiguous()
        rows, dim = y.shape[0] * world, y.shape[1]

        def body(t):
            ctx = self.ops.ctx
            S.zero_grad()
            if self._vae_fused:      # K11: forward + losses + backward in one launch
                return dict(dec=self.vae.step(t["y"], hp["ll_loss_vae"] / float(rows * dim), hp["kl_loss_vae"] / float(rows),
                                              S.acc[0:]))
            sv = {}
            dec, z, mu, sd = self.vae.forward(t["y"], kl_acc=S.acc[0:], save=sv)
            ddec = torch.empty_like(dec)
            ctx.l1_f32(dec.data_ptr(), t["y"].data_ptr(), ddec.data_ptr(), hp["ll_loss_vae"] / float(rows * dim), 0,
                       S.acc[1:].data_ptr(), dec.numel())
            self.vae.backward(sv, ddec, hp["kl_loss_vae"] / float(rows))
            return dict(dec=dec)

        if self.graphs and self.noise_mode == "device":
            st = self._graphed(("vae", y.shape[0], dim, self._hp_key(hp)), dict(y=y), body, lambda hyper: S.adam_step(hyper=hyper), S)
        else:
            st = body(dict(y=y))
            self._allreduce(S)
            S.adam_step()
        acc = S.acc[:2].cpu().numpy().astype(np.float64)
        self.vae_total_loss = np.float32(hp["kl_loss_vae"] * acc[0] / rows + hp["ll_loss_vae"] * acc[1] / (rows * dim))
        return st["dec"]

    # ------------------------------------------------------------------ dis_update (lsps_trainer.py:143-218)
    def dis_update(self, images_a, labels_a, images_b, labels_b, com_a=None, com_b=None, hyperparameters=None,
                   feat_mat=True):
        hp = hyperparameters or self.hp
        D, ctx, dis = self.dis_store, self.ops.ctx, self.dis
        world, _ = _world()
        ia, ib = _img(images_a, self.device), _img(images_b, self.device)
        B = ia.shape[0]
        Bg = B * world
        D.zero_grad()
        noise = self._latent_noise(2 * B, groups=2)
        train_map = bool(hp.get("train_map", False))
        ndiv = 4 if train_map else 3
        # discriminator input batches, by group: a = (ia | x_aa | x_ba [| dec_a]), b = (ib | x_ab | x_bb [| dec_b]).  The
        # reference concatenates (ia, x_ba, x_aa) (lsps_trainer.py:156-165); the order of the groups inside the batch is
        # immaterial (every layer is per-sample), and this one lets both decoders write straight into the batch.
        o = self.ops
        imgs_a, imgs_b = o.empty(ndiv * B, 128, 128, dtype=torch.float32), o.empty(ndiv * B, 128, 128, dtype=torch.float32)
        o.copy_into(imgs_a[:B], ia)
        o.copy_into(imgs_b[:B], ib)
        self.gen.forward(ia, ib, noise, self._scratch, out_a=imgs_a[B:3 * B], out_b=imgs_b[B:3 * B],
                         front=self._gen_front(ia, ib, keep=True))                                     # gen gets no grads
        if train_map:                                                        # :147-158, no activations kept either
            _, dec_a, dec_b = self._map_decode(labels_a, labels_b, hp)
            o.copy_into(imgs_a[3 * B:], dec_a)
            o.copy_into(imgs_b[3 * B:], dec_b)
        sv = {}
        F = dis.features(imgs_a, imgs_b, sv)                                 # [2*ndiv*B, 2, 2, 2048] (split: 2 x 2048)
        cf = dis.cf
        r = 4 * B                                                            # logits per group
        scale = hp["gan_w"] / float(4 * Bg)
        A_REAL, A_XAA, A_XBA, B_REAL, B_XAB, B_XBB = 0, 1, 2, ndiv, ndiv + 1, ndiv + 2
        # BCE targets per image group: real -> 1 (acc[0..1]); x_ba / x_ab -> 0 (acc[2..3]); the x_aa / x_bb groups only
        # feed the feature-matching term; train_map: decoded groups -> 0 (acc[8..9], ad_fake_dec :201-204)
        tgt, slot = [-1.0] * (2 * ndiv), [0] * (2 * ndiv)
        for real, fake in ((A_REAL, A_XBA), (B_REAL, B_XAB)):
            tgt[real], slot[real], tgt[fake], slot[fake] = 1.0, 0, 0.0, 2
        if train_map:
            for dec in (3, ndiv + 3):
                tgt[dec], slot[dec] = 0.0, 8
        dF = o.empty(F.shape[0], 4 * cf, dtype=torch.float32)
        dis.head_bce(F, r, tgt, slot, scale, dF, True, D.acc)
        if feat_mat:
            fscale = hp["feature_w"] / float(Bg * 4 * cf)
            # mean|F_b(x_ab) - F_a(x_aa)| + mean|F_a(x_ba) - F_b(x_bb)|      (lsps_trainer.py:171-177)
            for ga, gb in ((B_XAB, A_XAA), (A_XBA, B_XBB)):
                dis.l1_feat(F, ga * B, gb * B, B, dF, fscale, D.acc[4:])
        dFm = dis.mask_grad(dF, F)
        dis.features_bwd(sv, dFm, wgrad=True)
        self.ops.join_side()
        del sv
        self._allreduce(D)
        D.adam_step(active=lambda k: not k.startswith("Post."))
        D.refresh_dgrad_operands()
        acc = D.acc[:16].cpu().numpy().astype(np.float64)
        ad = (acc[0] + acc[2] + acc[8]) / (4 * Bg)
        feat = acc[4] / (Bg * 4 * cf) if feat_mat else 0.0
        self.dis_ad_loss, self.dis_feat_loss = np.float32(ad), np.float32(feat)
        self.dis_loss = np.float32(hp["gan_w"] * ad + hp["feature_w"] * feat)
        self.dis_true_acc = np.float32(acc[1] / (8 * Bg))
        self.dis_fake_acc = np.float32(acc[3] / (8 * Bg))

    # ------------------------------------------------------------------ gen_update (lsps_trainer.py:76-141)
    def gen_update(self, images_a, labels_a, images_b, labels_b, hyperparameters=None):
        hp = hyperparameters or self.hp
        G, D, ctx, gen, dis = self.gen_store, self.dis_store, self.ops.ctx, self.gen, self.dis
        world, _ = _world()
        ia, ib = _img(images_a, self.device), _img(images_b, self.device)
        B = ia.shape[0]
        Bg = B * world
        train_map = bool(hp.get("train_map", False))
        M = self.map_store
        G.zero_grad()
        if train_map and M is not None:
            M.zero_grad()
        n2 = self._latent_noise(2 * B, groups=2)  # host RNG draw order of the reference: gen(), forward_a2b, forward_b2a
        n3 = self._latent_noise(B)
        n4 = self._latent_noise(B)
        s1 = {}
        npx = float(Bg * _NPIX)
        o = self.ops
        doa = o.empty(2 * B, 128, 128, dtype=torch.float32)       # d/d(x_aa | x_ba)
        dob = o.empty(2 * B, 128, 128, dtype=torch.float32)       # d/d(x_ab | x_bb)
        d_bab, d_aba = o.empty(B, 128, 128, dtype=torch.float32), o.empty(B, 128, 128, dtype=torch.float32)
        # the four L1 reconstruction terms (lsps_trainer.py:118-121) are taken inside the decoder-head kernels
        oa, ob, shared = gen.forward(ia, ib, n2, G.acc[2:], s1, front=self._gen_front(ia, ib, keep=False),
                                     l1_a=(ia, 0, hp["ll_direct_link_w"] / npx, doa[:B], G.acc[5:]),
                                     l1_b=(ib, B, hp["ll_direct_link_w"] / npx, dob[B:], G.acc[6:]))
        x_aa, x_ba, x_ab, x_bb = oa[:B], oa[B:], ob[:B], ob[B:]
        s2 = {}
        x_bab, x_aba = gen.forward_cycle(x_ba, x_ab, n3 if isinstance(n3, tuple) else torch.cat((n3, n4), 0), G.acc[3:],
                                         G.acc[4:], s2,
                                         l1s=((ib, 0, hp["ll_cycle_link_w"] / npx, d_bab, G.acc[8:]),
                                              (ia, 0, hp["ll_cycle_link_w"] / npx, d_aba, G.acc[7:])))
        del n2, n3, n4
        dec_a, dec_b, nd = x_ba, x_ab, 1
        if train_map:                        # :84-99 (the vae.encode draw comes after the three latent draws, as there)
            sm, sdm = {}, {}
            z_map, dec_a, dec_b = self._map_decode(labels_a, labels_b, hp, sm, sdm)
            nd = 2
        # adversarial term through the discriminator (data gradient only)
        sd = {}
        F = dis.features(self.ops.cat((x_ba, dec_a)) if train_map else x_ba,
                         self.ops.cat((x_ab, dec_b)) if train_map else x_ab, sd)
        dF = self.ops.empty(F.shape[0], 4 * dis.cf, dtype=torch.float32)
        dis.head_bce(F, dis.rows(F), [1.0], [0], hp["gan_w"] / float(4 * nd * Bg), dF, False, G.acc)
        dFm = dis.mask_grad(dF, F)
        if train_map:
            dia, dib = torch.empty(2 * B, 128, 128, device=self.device), torch.empty(2 * B, 128, 128, device=self.device)
            dis.features_bwd(sd, dFm, wgrad=False, dimg_a=dia, dimg_b=dib)
            self.ops.copy_into(doa[B:], dia[:B])
            self.ops.copy_into(dob[:B], dib[:B])
            d_dec_a, d_dec_b = dia[B:], dib[B:]
        else:
            dis.features_bwd(sd, dFm, wgrad=False, dimg_a=doa[B:], dimg_b=dob[:B])
        del sd
        g_z = None
        if train_map:
            # matching losses (:97-99): ll_map_w * (L1(decode_A, images_a) + L1(decode_B, images_b)) on top of the
            # adversarial gradient; ll_map_z_w * mean((shared - z_pose2depth)^2) pulls both latents together
            ctx.l1_f32(dec_a.data_ptr(), ia.data_ptr(), d_dec_a.data_ptr(), hp["ll_map_w"] / npx, 1, G.acc[10:].data_ptr(), dec_a.numel())
            ctx.l1_f32(dec_b.data_ptr(), ib.data_ptr(), d_dec_b.data_ptr(), hp["ll_map_w"] / npx, 1, G.acc[11:].data_ptr(), dec_b.numel())
            g_z = torch.empty_like(shared)
            ctx.l2_bf16(shared.data_ptr(), z_map.data_ptr(), g_z.data_ptr(), 2.0 * hp["ll_map_z_w"] / float(2 * Bg * _LATENT),
                        G.acc[12:].data_ptr(), shared.numel())
            dzm = gen.decode_bwd(sdm, d_dec_a.contiguous(), d_dec_b.contiguous())
            ctx.axpy_bf16(dzm.data_ptr(), g_z.data_ptr(), -1.0, dzm.data_ptr(), dzm.numel())
            self.map.backward(sm, dzm)
            del sdm, sm
        gen.backward_cycle(s2, d_bab, d_aba, hp["kl_cycle_link_w"] / float(Bg * _LATENT), doa[B:], dob[:B])
        del s2
        # kl_direct * (enc + enc) with enc = mean over the 2B latents
        gen.backward(s1, doa, dob, 2.0 * hp["kl_direct_link_w"] / float(2 * Bg * _LATENT), dz_extra=g_z)
        self.ops.join_side()
        del s1
        self._allreduce(G)
        G.adam_step()
        G.refresh_dgrad_operands()
        if train_map:                        # gen_opt holds gen + map parameters (lsps_trainer.py:27)
            self._allreduce(M)
            M.adam_step()
            M.refresh_dgrad_operands()
        acc = G.acc[:16].cpu().numpy().astype(np.float64)
        ad = acc[0] / (4 * nd * Bg)
        enc, enc2 = acc[2] / (2 * Bg * _LATENT), (acc[3] + acc[4]) / (Bg * _LATENT)
        ll, ll2 = (acc[5] + acc[6]) / npx, (acc[7] + acc[8]) / npx
        self.gen_enc_loss, self.gen_enc_loss2 = np.float32(enc), np.float32(enc2)
        self.gen_ad_loss = np.float32(ad)
        self.gen_ll_loss, self.gen_ll_loss2 = np.float32(ll), np.float32(ll2)
        total = (hp["gan_w"] * ad + hp["ll_direct_link_w"] * ll + hp["ll_cycle_link_w"] * ll2 +
                 hp["kl_direct_link_w"] * 2.0 * enc + hp["kl_cycle_link_w"] * enc2)
        if train_map:
            mz, ml = acc[12] / (2 * Bg * _LATENT), (acc[10] + acc[11]) / npx
            self.gen_map_loss, self.gen_map_loss2 = np.float32(mz), np.float32(ml)
            total += hp["ll_map_z_w"] * mz + hp["ll_map_w"] * ml
        self.gen_total_loss = np.float32(total)
        u = lambda t: t.unsqueeze(1)
        return (u(x_aa), u(x_ba), u(x_ab), u(x_bb), u(x_aba), u(x_bab), u(dec_a), u(dec_b))

    # ------------------------------------------------------------------ post_update (lsps_trainer.py:220-262)
    def post_update(self, images_a, labels_a, images_b, labels_b, com_a=None, com_b=None, mode=3, hyperparameters=None):
        hp = hyperparameters or self.hp
        world, rank = _world()
        ia, ib = _img(images_a, self.device), _img(images_b, self.device)
        la = labels_a.detach().to(device=self.device, dtype=torch.float32).contiguous()
        lb = labels_b.detach().to(device=self.device, dtype=torch.float32).contiguous()
        # the reference uses the GLOBAL first 4 samples of each domain (lsps_trainer.py:238)
        src_a, src_b = ia[0:4], ib[0:4]
        if world > 1 and mode >= 2:
            src_a, src_b = src_a.clone(), src_b.clone()
            dist.broadcast(src_a, 0)
            dist.broadcast(src_b, 0)
        if self.graphs and self.noise_mode == "device":
            st = self._graphed(("post", mode, ia.shape[0], la.shape[1], self._hp_key(hp)), dict(ia=ia, ib=ib, la=la, lb=lb, sa=src_a, sb=src_b),
                               lambda t: self._post_body(t["ia"], t["la"], t["ib"], t["lb"], t["sa"], t["sb"], mode, hp),
                               lambda hyper: self._post_tail(mode, hyper), self.dis_store)
        else:
            st = self._post_body(ia, la, ib, lb, src_a, src_b, mode, hp)
            self._allreduce(self.dis_store)
            self._post_tail(mode, None)
        return self._post_finish(st, hp)

    def _post_body(self, ia, la, ib, lb, src_a, src_b, mode, hp):
        """zero_grad .. backward of post_update (everything before the gradient allreduce)."""
        D, ctx, dis = self.dis_store, self.ops.ctx, self.dis
        world, rank = _world()
        B = ia.shape[0]
        Bg = B * world
        D.zero_grad()
        reg_a, reg_b, feat = mode != 1, mode in (1, 4), mode >= 2
        outs = (ia, ia, ib, ib)
        fa_extra = fb_extra = None
        nf, n4 = 0, 4
        if feat:
            # per-sample ops make it exact to give source image a_i to rank i % world and b_i to rank (4+i) % world
            n4 = src_a.shape[0]
            noise = self._latent_noise(src_a.shape[0] + src_b.shape[0], shard=False)
            ka, kb, idx = feature_sources(src_a.shape[0], src_b.shape[0], world, rank)
            if ka or kb:
                # row picks as slices + device copies (a python-list index would be a host->device copy, illegal in a capture)
                pick = lambda t, idx: t if len(idx) == t.shape[0] else self.ops.cat([t[i:i + 1] for i in idx])
                xa = pick(src_a, ka) if ka else None
                xb = pick(src_b, kb) if kb else None
                nz = noise if isinstance(noise, tuple) else pick(noise, idx)
                oa, ob, _ = self.gen.forward(xa, xb, nz, self._scratch)     # (x_aa|x_ba), (x_ab|x_bb)
                na, nb = len(ka), len(kb)
                nf = na + nb
                fa_extra, fb_extra = oa, ob
                outs = (oa[:na], oa[na:], ob[:na], ob[na:])
        imgs_a = [t for t in (fa_extra, ia if reg_a else None) if t is not None]
        imgs_b = [t for t in (fb_extra, ib if reg_b else None) if t is not None]
        imgs_a = self.ops.cat(imgs_a) if imgs_a else None
        imgs_b = self.ops.cat(imgs_b) if imgs_b else None
        sv = {}
        F = dis.features(imgs_a, imgs_b, sv)
        cf = dis.cf
        per = 4 * cf
        n_a = imgs_a.shape[0] if imgs_a is not None else 0
        dF = self.ops.zeros(F.shape[0], per)
        pd = hp["dis"]["post_dim"]
        preds = []
        for dom, on, row0, labels, slot in (("a", reg_a, nf, la, 5), ("b", reg_b, n_a + nf, lb, 6)):
            if not on:
                continue
            p = dis.post(F, row0, B)
            e = self.vae.encode(labels)[0]
            dp = torch.empty_like(p)
            ctx.mse(p.data_ptr(), e.data_ptr(), dp.data_ptr(), 2.0 * hp["reg_w"] / float(Bg * pd), D.acc[slot:].data_ptr(),
                    p.numel())
            dis.post_bwd(F, row0, B, dp, dF)
            preds.append(p)
        if nf:
            na = outs[0].shape[0]
            nb = nf - na
            fscale = hp["feature_w_reg"] / float(n4 * per)
            # rows: FA part = (x_aa[na] | x_ba[nb]) ; FB part (offset n_a) = (x_ab[na] | x_bb[nb])
            if na:   # mean|f_ab - f_aa|
                dis.l1_feat(F, n_a, 0, na, dF, fscale, D.acc[7:])
            if nb:   # mean|f_ba - f_bb|
                dis.l1_feat(F, na, n_a + na, nb, dF, fscale, D.acc[7:])
        dFm = dis.mask_grad(dF, F)
        dis.features_bwd(sv, dFm, wgrad=True)
        self.ops.join_side()
        return dict(outs=outs, preds=preds, Bg=Bg, pd=pd, per=per, n4=n4, feat=feat)

    def _post_tail(self, mode, hyper):
        D = self.dis_store
        reg_a, reg_b, feat = mode != 1, mode in (1, 4), mode >= 2
        skip = ["D."] + ([] if (reg_a or feat) else ["model_A."]) + ([] if (reg_b or feat) else ["model_B."])
        segs = D.adam_step(active=lambda k: not any(k.startswith(s) for s in skip), hyper=hyper)
        D.refresh_dgrad_operands()
        return segs

    def _post_finish(self, st, hp):
        acc = self.dis_store.acc[:8].cpu().numpy().astype(np.float64)
        reg = (acc[5] + acc[6]) / (st["Bg"] * st["pd"])
        fm = acc[7] / (st["n4"] * st["per"]) if st["feat"] else 0.0
        self.dis_reg_loss = np.float32(reg)
        self.dis_total_loss = np.float32(hp["reg_w"] * reg + hp["feature_w_reg"] * fm)
        self.last_pred_post = st["preds"]
        u = lambda t: t.unsqueeze(1)
        x_aa, x_ba, x_ab, x_bb = st["outs"]
        return (u(x_aa), u(x_ba), u(x_ab), u(x_bb), u(x_aa), u(x_bb), u(x_aa), u(x_bb))

    # ------------------------------------------------------------------ CUDA graphs (launch-bound updates)
    @staticmethod
    def _hp_key(hp):
        """loss weights are baked into a captured graph: they are part of its key"""
        return tuple(sorted((k, float(v)) for k, v in hp.items() if isinstance(v, (int, float)) and not isinstance(v, bool)))

    def _graphed(self, key, inputs, body, tail, store):
        """Runs `body(inputs) -> state`, the gradient allreduce and `tail(hyper) -> adam segments` with the device work
        of body and tail replayed from two captured CUDA graphs (estimate-mode steps are ~120 small launches: host
        launch overhead, not the GPU, bounds them).  The first two calls per key run eagerly (they set kernel
        attributes / fill caches), the third captures.  The NCCL allreduce stays outside the graphs; Adam's
        step-dependent factors are read from a small device buffer refreshed before every replay."""
        ent = self._graphs.get(key)
        if ent is None:
            ent = self._graphs[key] = dict(calls=0)
        if ent["calls"] < 2:
            ent["calls"] += 1
            st = body(inputs)
            self._allreduce(store)
            tail(None)
            return st
        if "g1" not in ent:
            ent["static"] = {k: v.clone() for k, v in inputs.items()}
            ent["hyper"] = torch.zeros(32, dtype=torch.float32, device=self.device)
            ent["hyper_host"] = torch.zeros(32, dtype=torch.float32).pin_memory()
            torch.cuda.synchronize()
            ent["g1"] = torch.cuda.CUDAGraph()
            with torch.cuda.graph(ent["g1"]):
                ent["state"] = body(ent["static"])
            steps = {k: e.step for k, e in store.entries.items()}
            ent["g2"] = torch.cuda.CUDAGraph()
            with torch.cuda.graph(ent["g2"]):
                ent["segs"] = tail(ent["hyper"])
            for k, v in steps.items():           # capture executes nothing: undo its step bookkeeping
                store.entries[k].step = v
        if not store.segments_consistent(ent["segs"]):
            # another update type stepped only part of a captured Adam segment: fall back to eager and re-capture later
            self._graphs.pop(key)
            st = body(inputs)
            self._allreduce(store)
            tail(None)
            return st
        for k, v in inputs.items():
            ent["static"][k].copy_(v)
        ent["g1"].replay()
        self._allreduce(store)
        store.advance(ent["segs"], ent["hyper_host"])
        ent["hyper"].copy_(ent["hyper_host"], non_blocking=True)
        ent["g2"].replay()
        # the captured state lives in graph-owned memory that the next replay overwrites: hand out copies
        cl = lambda v: v.clone() if torch.is_tensor(v) else (type(v)(cl(x) for x in v) if isinstance(v, (tuple, list)) else v)
        return {k: cl(v) for k, v in ent["state"].items()}

    # ------------------------------------------------------------------ outputs / snapshots
    def assemble_outputs(self, images_a, images_b, network_outputs):
        """lsps_trainer.py:264-276: first sample of each tensor concatenated along width -> (1,1,128,1280)."""
        f = lambda t: t[0:1].detach().to(self.device).reshape(1, 1, t.shape[-2], t.shape[-1]).float()
        o = network_outputs
        return torch.cat((f(images_a), f(o[0]), f(o[2]), f(o[4]), f(o[6]), f(o[7]), f(images_b), f(o[3]), f(o[1]),
                          f(o[5])), 3)

    def save(self, snapshot_prefix, iterations):
        """lsps_trainer.py:307-319: <prefix>_gen_%08d.pkl / _dis_ ; state_dict keys and shapes of the reference."""
        torch.save({k: v.cpu() for k, v in self.gen_store.state_dict().items()}, "%s_gen_%08d.pkl" % (snapshot_prefix, iterations + 1))
        torch.save({k: v.cpu() for k, v in self.dis_store.state_dict().items()}, "%s_dis_%08d.pkl" % (snapshot_prefix, iterations + 1))
        opt = {"gen": self._cpu(self.gen_store.opt_state()), "dis": self._cpu(self.dis_store.opt_state())}
        if self.map_store is not None:   # the reference has this line commented out (:319) yet resume() looks for it (:299)
            torch.save({k: v.cpu() for k, v in self.map_store.state_dict().items()}, "%s_map_%08d.pkl" % (snapshot_prefix, iterations + 1))
            opt["map"] = self._cpu(self.map_store.opt_state())
        torch.save(opt, "%s_opt_%08d.pkl" % (snapshot_prefix, iterations + 1))

    @staticmethod
    def _cpu(st):
        return {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in st.items()}

    @staticmethod
    def _model_list(dirname, key, idx=-1):
        """helpers.py:9-18 `get_model_list`: files of `dirname` whose name contains `key` and "pkl", sorted."""
        if not os.path.isdir(dirname):
            return None
        files = sorted(os.path.join(dirname, f) for f in os.listdir(dirname)
                       if os.path.isfile(os.path.join(dirname, f)) and key in f and "pkl" in f)
        if not files:
            return None
        try:
            return files[idx]
        except IndexError:
            return None

    def resume(self, snapshot_prefix, idx=-1, load_opt=False, est=False):
        """lsps_trainer.py:278-305: gen and dis of the newest (or idx-th) snapshot in the prefix's directory -- with
        `est` both come from the `*_est_*` files; the iteration count is parsed from the GENERATOR file name; the
        Mapping snapshot is optional.  Optimiser state (which the reference's save() leaves commented out) is read
        from this trainer's own `*_opt_*` file when `load_opt` is set."""
        dirname = os.path.dirname(snapshot_prefix)
        f = self._model_list(dirname, "est_gen" if est else "gen", idx)
        if f is None:
            return 0
        self.gen_store.load_state_dict(torch.load(f, map_location="cpu"), strict=False)
        iterations = int(f[-12:-4])
        fd = self._model_list(dirname, "est_dis" if est else "dis", idx)
        if fd is not None:
            self.dis_store.load_state_dict(torch.load(fd, map_location="cpu"), strict=False)
        if load_opt:
            fo = f.replace("_gen_", "_opt_")
            if os.path.exists(fo):
                st = torch.load(fo, map_location="cpu")
                for net, store in (("gen", self.gen_store), ("dis", self.dis_store), ("map", self.map_store)):
                    if store is not None and st.get(net) is not None:
                        store.load_opt_state({k: (v.to(self.device) if torch.is_tensor(v) else v)
                                              for k, v in st[net].items()})
        if self.map_store is not None:
            fm = self._model_list(dirname, "map", idx)
            if fm is not None:
                self.map_store.load_state_dict(torch.load(fm, map_location="cpu"), strict=False)
        return iterations

    def save_vae(self, snapshot_prefix, iterations, frac):
        torch.save({k: v.cpu() for k, v in self.vae_store.state_dict().items()},
                   "%s_vae_%.2f_%08d.pkl" % (snapshot_prefix, frac, iterations + 1))

    def load_vae(self, snapshot_prefix, frac):
        f = self._model_list(os.path.dirname(snapshot_prefix), "vae_%.2f" % frac)
        if f is None:
            raise IOError("no pose-VAE snapshot vae_%.2f next to %s" % (frac, snapshot_prefix))
        self.vae_store.load_state_dict(torch.load(f, map_location="cpu"))
