"""Shared helpers of the parity tests: run a schedule of updates on a trainer-like object."""
import numpy as np
import torch

import lsps_oracle as O

LOSS_KEYS = ("dis_loss", "dis_ad_loss", "dis_feat_loss", "dis_true_acc", "dis_fake_acc", "gen_total_loss",
             "gen_ad_loss", "gen_ll_loss", "gen_ll_loss2", "gen_enc_loss", "gen_enc_loss2",
             "dis_reg_loss", "dis_total_loss", "vae_total_loss", "gen_map_loss", "gen_map_loss2")


def sample(t):
    """Same strided sample + moments as oracle/make_golden.py:_sample."""
    t = t.detach().float().cpu()
    flat = t.reshape(-1)
    step = max(1, flat.numel() // 256)
    return np.concatenate([flat[::step][:256].numpy(),
                           np.array([flat.mean().item(), flat.abs().mean().item(), flat.std().item()], np.float32)])


def run_schedule(tr, hp, schedule, batch, steps, kind="uniform", device=None, on_step=None):
    """Mirrors oracle/make_golden.py:run_case for one trainer.  Returns {key: value} like the golden files."""
    label_dim = hp["vae"]["input_dim"]
    g = torch.Generator().manual_seed(1234)
    torch.manual_seed(42)
    rec = {}
    for s in range(steps):
        ia, ib, la, lb = O.synthetic_batch(batch, label_dim, g, kind)
        com = torch.zeros(batch, 3)
        if on_step is not None:
            on_step(s, tr)
        if device is not None:
            ia, ib, la, lb = (t.to(device) for t in (ia, ib, la, lb))
        for upd in schedule:
            if upd == "vae":
                out = tr.vae_update(torch.cat((la, lb), 0), hp)
                rec["s%d_vae_dec" % s] = sample(out)
            elif upd == "dis":
                tr.dis_update(ia, la, ib, lb, com, com, hp)
            elif upd == "gen":
                outs = tr.gen_update(ia, la, ib, lb, hp)
                for i, nm in enumerate(("x_aa", "x_ba", "x_ab", "x_bb", "x_aba", "x_bab")):
                    rec["s%d_%s" % (s, nm)] = sample(outs[i])
                if hp["train_map"]:
                    rec["s%d_decode_A" % s], rec["s%d_decode_B" % s] = sample(outs[6]), sample(outs[7])
            elif upd.startswith("post"):
                outs = tr.post_update(ia, la, ib, lb, com, com, int(upd[4:]), hp)
                rec["s%d_post_x_ba" % s] = sample(outs[1])
        for k in LOSS_KEYS:
            if hasattr(tr, k):
                rec["s%d_%s" % (s, k)] = np.float32(np.asarray(getattr(tr, k)))
    return rec


def load_from_oracle(tr, oracle):
    tr.gen_store.load_state_dict(oracle.state_dict("gen"))
    tr.dis_store.load_state_dict(oracle.state_dict("dis"))
    tr.vae_store.load_state_dict(oracle.state_dict("vae"))
    if "map" in oracle.params:
        tr.map_store.load_state_dict(oracle.state_dict("map"))


GOLDEN_CASES = {
    # name: (config, schedule, batch, steps, kind)
    "vae_nnyu_b8": ("nnyu", ["vae"], 8, 10, "uniform"),
    "vae_nicvl_b8": ("nicvl", ["vae"], 8, 10, "uniform"),
    "pretrain_nnyu_b1": ("nnyu", ["dis", "gen"], 1, 2, "uniform"),
    "pretrain_nnyu_b2_hand": ("nnyu", ["dis", "gen"], 2, 1, "hand"),
    "estimate3_nnyu_b8": ("nnyu", ["post3"], 8, 3, "uniform"),
    "estimate0_nnyu_b4": ("nnyu", ["post0"], 4, 2, "uniform"),
    "estimate4_nnyu_b5": ("nnyu", ["post4"], 5, 1, "uniform"),
    "estimate1_nnyu_b4": ("nnyu", ["post1"], 4, 2, "uniform"),
    "estimate3_nicvl_b4": ("nicvl", ["post3"], 4, 1, "uniform"),
    # train_map=True branches (Mapping net): config name "<yaml>:map"
    "pretrain_map_nnyu_b1": ("nnyu:map", ["dis", "gen"], 1, 2, "uniform"),
    "pretrain_map_nnyu_b2_hand": ("nnyu:map", ["dis", "gen"], 2, 1, "hand"),
    # ResNeXt generator (gen.name = SharedResXGen; default k=1, cardinality 4, and k=2, cardinality 8)
    "pretrain_resx_nnyu_b1": ("nnyu:resx", ["dis", "gen"], 1, 2, "uniform"),
    "estimate3_resx_nnyu_b4": ("nnyu:resx", ["post3"], 4, 1, "uniform"),
    "pretrain_resx_k2c8_nnyu_b1": ("nnyu:resx_k2c8", ["dis", "gen"], 1, 1, "uniform"),
}


def load_hp(name):
    """Hyper-parameters of exps/<yaml>.yaml; a ":map" suffix switches the train_map branches on."""
    import os
    import yaml
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    base, _, opt = name.partition(":")
    with open(os.path.join(root, "exps", base + ".yaml")) as fh:
        hp = yaml.safe_load(fh)["train"]["hyperparameters"]
    if opt == "map":
        hp["train_map"] = True
    if opt.startswith("resx"):
        hp["gen"]["name"] = "SharedResXGen"
        if opt == "resx_k2c8":
            hp["gen"].update(n_resnext_k=2, n_resnext_c=8)
    return hp
