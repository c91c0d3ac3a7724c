"""TEST INFRASTRUCTURE ONLY -- pins `oracle/lsps_oracle.py` against the unmodified reference
and writes the committed fixtures tests/golden/*.npz.

Run in the build container (needs /root/reference):   python oracle/make_golden.py

For every case: the oracle's deterministic weights are loaded into the *reference*
LSPSTrainer (`load_state_dict`), both consume the same host RNG stream
(`torch.manual_seed(42)`), both run the same updates on the same synthetic batches
(generator seed 1234).  The reference's numbers are what is stored; the script fails if
the oracle port deviates from them by more than float32 round-off.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import lsps_oracle as O          # noqa: E402
import ref_loader                # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")
LOSS_KEYS = ("dis_loss", "dis_ad_loss", "dis_feat_loss", "dis_true_acc", "dis_fake_acc", "gen_total_loss",
             "gen_ad_loss", "gen_ll_loss", "gen_ll_loss2", "gen_enc_loss", "gen_enc_loss2",
             "dis_reg_loss", "dis_total_loss", "vae_total_loss", "gen_map_loss", "gen_map_loss2")


def _ref_trainer(trainers, hp, oracle):
    torch.manual_seed(0)
    tr = trainers.LSPSTrainer(hp)
    tr.gpu = 0
    tr.gen.load_state_dict(oracle.state_dict("gen"))
    tr.dis.load_state_dict(oracle.state_dict("dis"))
    tr.vae.load_state_dict(oracle.state_dict("vae"))
    if "map" in oracle.params:
        tr.map.load_state_dict(oracle.state_dict("map"))
    return tr


def _losses(obj):
    out = {}
    for k in LOSS_KEYS:
        if hasattr(obj, k):
            out[k] = float(np.asarray(getattr(obj, k)))
    return out


def _sample(t):
    """Strided sample + moments of a tensor: small enough to commit, sharp enough to catch a wrong kernel."""
    t = t.detach().float()
    flat = t.reshape(-1)
    step = max(1, flat.numel() // 256)
    return np.concatenate([flat[::step][:256].numpy(),
                           np.array([flat.mean().item(), flat.abs().mean().item(), flat.std().item()], np.float32)])


def run_case(trainers, name, hp, schedule, batch, kind="uniform", steps=2, seed=0):
    """schedule: list of update names executed per step."""
    label_dim = hp["vae"]["input_dim"]
    results = {}
    for who in ("ref", "oracle"):
        oracle = O.OracleTrainer(hp, seed=seed)
        tr = _ref_trainer(trainers, hp, oracle) if who == "ref" else oracle
        g = torch.Generator().manual_seed(1234)
        torch.manual_seed(42)
        rec = {}
        for s in range(steps):
            ia, ib, la, lb = O.synthetic_batch(batch, label_dim, g, kind)
            com = torch.zeros(batch, 3)
            for upd in schedule:
                if upd == "vae":
                    out = tr.vae_update(torch.cat((la, lb), 0), hp)
                    rec["s%d_vae_dec" % s] = _sample(out)
                elif upd == "dis":
                    tr.dis_update(ia, la, ib, lb, com, com, hp)
                elif upd == "gen":
                    outs = tr.gen_update(ia, la, ib, lb, hp)
                    for i, nm in enumerate(("x_aa", "x_ba", "x_ab", "x_bb", "x_aba", "x_bab")):
                        rec["s%d_%s" % (s, nm)] = _sample(outs[i])
                    if hp["train_map"]:
                        rec["s%d_decode_A" % s], rec["s%d_decode_B" % s] = _sample(outs[6]), _sample(outs[7])
                elif upd.startswith("post"):
                    outs = tr.post_update(ia, la, ib, lb, com, com, int(upd[4:]), hp)
                    rec["s%d_post_x_ba" % s] = _sample(outs[1])
            for k, v in _losses(tr).items():
                rec["s%d_%s" % (s, k)] = np.float32(v)
        sd = (lambda n: getattr(tr, n).state_dict()) if who == "ref" else tr.state_dict
        resx = hp["gen"].get("name") == "SharedResXGen"
        for net, keys in (("dis", ("model_S.3.model.0.weight", "model_A.0.model.0.weight", "D.weight", "Post.weight")),
                          ("gen", ("encode_A.0.model.0.weight", "enc_shared.0.model.0.weight", "decode_B.5.weight") +
                           (("enc_shared.0.model.3.weight", "decode_A.1.model.6.weight") if resx else ())),
                          ("vae", ("en_fc1.weight", "de_fc2.bias")),
                          ("map", ("model.0.model.0.weight", "model.1.model.0.weight", "model.2.model.0.bias",
                                   "model.3.weight") if hp["train_map"] else ())):
            if not keys:
                continue
            d = sd(net)
            for k in keys:
                rec["w_%s_%s" % (net, k)] = _sample(d[k])
        results[who] = rec
    worst = 0.0
    for k, v in results["ref"].items():
        o = results["oracle"][k]
        err = float(np.max(np.abs(np.asarray(v, np.float64) - np.asarray(o, np.float64)) /
                           (1e-6 + np.abs(np.asarray(v, np.float64)))) if np.ndim(v) == 0 else
                    np.max(np.abs(v - o)) / (1e-6 + np.max(np.abs(v))))
        worst = max(worst, err)
        if err > 2e-4:
            raise SystemExit("oracle deviates from the reference on %s/%s: %g" % (name, k, err))
    meta = dict(batch=batch, steps=steps, kind=kind, seed=seed, schedule=",".join(schedule))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **results["ref"],
                        **{"meta_" + k: np.array(v) for k, v in meta.items()})
    print("%-28s ok  (oracle vs reference worst rel %.2e, %d arrays)" % (name, worst, len(results["ref"])))


def main(only=None):
    global run_case
    if only:
        _run = run_case

        def run_case(trainers, name, *a, **k):   # noqa: regenerate just the named fixtures
            if name in only:
                _run(trainers, name, *a, **k)
    os.makedirs(OUT, exist_ok=True)
    trainers = ref_loader.load_reference()
    torch.set_num_threads(os.cpu_count())
    nnyu, nicvl = ref_loader.load_hyperparameters("nnyu"), ref_loader.load_hyperparameters("nicvl")
    # config 1 (BASELINE.json configs[0]): pose_train VAE, batch 8 per domain, 10 its
    run_case(trainers, "vae_nnyu_b8", nnyu, ["vae"], batch=8, steps=10)
    run_case(trainers, "vae_nicvl_b8", nicvl, ["vae"], batch=8, steps=10)
    # config 2: pretrain step = dis_update + gen_update
    run_case(trainers, "pretrain_nnyu_b1", nnyu, ["dis", "gen"], batch=1, steps=2)
    run_case(trainers, "pretrain_nnyu_b2_hand", nnyu, ["dis", "gen"], batch=2, steps=1, kind="hand")
    # config 3: estimate modes
    run_case(trainers, "estimate3_nnyu_b8", nnyu, ["post3"], batch=8, steps=3)
    run_case(trainers, "estimate0_nnyu_b4", nnyu, ["post0"], batch=4, steps=2)
    run_case(trainers, "estimate4_nnyu_b5", nnyu, ["post4"], batch=5, steps=1)
    run_case(trainers, "estimate1_nnyu_b4", nnyu, ["post1"], batch=4, steps=2)
    # config 4: ICVL shapes (48-d pose vector; conv nets identical)
    run_case(trainers, "estimate3_nicvl_b4", nicvl, ["post3"], batch=4, steps=1)
    # the train_map=True branches (Mapping net, ndiv=4 discriminator batch, map losses)
    nnyu_map = dict(nnyu, train_map=True)
    run_case(trainers, "pretrain_map_nnyu_b1", nnyu_map, ["dis", "gen"], batch=1, steps=2)
    run_case(trainers, "pretrain_map_nnyu_b2_hand", nnyu_map, ["dis", "gen"], batch=2, steps=1, kind="hand")
    # the ResNeXt generator (lsps_nets.py:277-387), selected from the YAML by gen.name like every net;
    # default k = 1, cardinality 4 (64-channel groups) and the k = 2, cardinality 8 variant
    import copy
    resx = copy.deepcopy(nnyu)
    resx["gen"]["name"] = "SharedResXGen"
    run_case(trainers, "pretrain_resx_nnyu_b1", resx, ["dis", "gen"], batch=1, steps=2)
    run_case(trainers, "estimate3_resx_nnyu_b4", resx, ["post3"], batch=4, steps=1)
    resx2 = copy.deepcopy(resx)
    resx2["gen"].update(n_resnext_k=2, n_resnext_c=8)
    run_case(trainers, "pretrain_resx_k2c8_nnyu_b1", resx2, ["dis", "gen"], batch=1, steps=1)


if __name__ == "__main__":
    main(only=set(sys.argv[1:]))
