"""Host-side logic that needs no GPU: parameter tables, layout conversions, LR schedule, config surface, sharding."""
import os

import torch
import yaml

import lsps_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _hp(name):
    from common import load_hp
    return load_hp(name)


def test_parameter_tables_match_reference_state_dict_keys():
    from lsps_b200.params import gen_entries, dis_entries, vae_entries, map_entries
    for cfg in ("nnyu", "nicvl"):
        hp = _hp(cfg)
        for mine, spec in ((gen_entries(hp["gen"]), O.gen_spec(hp["gen"])), (dis_entries(hp["dis"]), O.dis_spec(hp["dis"])),
                           (vae_entries(hp["vae"]), O.vae_spec(hp["vae"])), (map_entries(hp["map"]), O.map_spec(hp["map"]))):
            assert [e[0] for e in mine] == list(spec.keys())
            for e in mine:
                assert tuple(e[1]) == tuple(spec[e[0]][0]), e[0]
                assert e[4] == spec[e[0]][2], ("fan_in", e[0])


def test_kernel_layout_round_trip_and_semantics():
    from lsps_b200.params import to_kernel_layout, from_kernel_layout
    g = torch.Generator().manual_seed(0)
    for kind, shape in (("conv3", (8, 4, 3, 3)), ("deconv3", (4, 8, 3, 3)), ("deconv4", (4, 8, 4, 4)), ("map0", (5, 8, 4, 4)), ("post", (20, 16, 2, 2)), ("stem", (64, 1, 7, 7)),
                        ("head", (64, 1, 1, 1)), ("dhead", (1, 32, 1, 1)), ("linear", (5, 7)), ("bias", (9,))):
        t = torch.randn(shape, generator=g)
        k = to_kernel_layout(kind, t)
        assert torch.equal(from_kernel_layout(kind, k.reshape(-1), shape), t)
    w = torch.randn(8, 4, 3, 3, generator=g)
    k = to_kernel_layout("conv3", w)
    assert k.shape == (9, 8, 4) and torch.equal(k[1 * 3 + 2], w[:, :, 1, 2])
    wt = torch.randn(4, 8, 3, 3, generator=g)     # ConvTranspose2d IOHW
    k = to_kernel_layout("deconv3", wt)
    assert k.shape == (9, 8, 4) and torch.equal(k[2 * 3 + 0], wt[:, :, 2, 0].t())
    w4 = torch.randn(4, 8, 4, 4, generator=g)     # Mapping: ConvTranspose2d k4, IOHW -> [tap = r*4+s][co][ci]
    k = to_kernel_layout("deconv4", w4)
    assert k.shape == (16, 8, 4) and torch.equal(k[3 * 4 + 1], w4[:, :, 3, 1].t())
    # ... and layer 0 (k4 s1 p0 on a 1x1 input) read as a dense layer [16*co][ci] producing NHWC (4, 4, co)
    e = torch.randn(3, 4, generator=g)
    ref = torch.nn.functional.conv_transpose2d(e[:, :, None, None], w4).permute(0, 2, 3, 1).reshape(3, -1)
    assert torch.allclose(ref, e @ to_kernel_layout("map0", w4).reshape(16 * 8, 4).t(), atol=1e-5)
    p = torch.randn(20, 16, 2, 2, generator=g)    # Post conv == FC over (pos, channel)
    k = to_kernel_layout("post", p)
    f = torch.randn(3, 16, 2, 2, generator=g)
    ref = torch.nn.functional.conv2d(f, p).reshape(3, 20)
    got = f.permute(0, 2, 3, 1).reshape(3, -1) @ k.t()
    assert torch.allclose(ref, got, atol=1e-5)


def test_multistep_lr_matches_torch():
    from lsps_b200.params import MultiStepLR

    class S(object):
        lr = base_lr = 1e-4
    s = S()
    mine = MultiStepLR(s, [200, 300, 400, 450], 0.5)
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([p], lr=1e-4)
    ref = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[200, 300, 400, 450], gamma=0.5)
    for _ in range(460):
        opt.step()
        ref.step()
        mine.step()
        assert abs(mine.get_lr()[0] - ref.get_last_lr()[0]) < 1e-12


def test_config_surface():
    from lsps_b200.config import NetConfig
    cfg = NetConfig(os.path.join(ROOT, "exps", "nnyu.yaml"))
    for k in ("snapshot_save_iterations", "image_save_iterations", "image_display_iterations", "display",
              "snapshot_prefix", "hyperparameters", "datasets"):
        assert hasattr(cfg, k)
    hp = cfg.hyperparameters
    assert hp["trainer"] == "LSPSTrainerB200" and hp["gen"]["name"] == "SharedResGen"
    for k in ("lr", "gan_w", "feature_w", "feature_w_reg", "reg_w", "ll_direct_link_w", "ll_cycle_link_w",
              "kl_direct_link_w", "kl_cycle_link_w", "train_map", "ll_loss_vae", "kl_loss_vae", "batch_size",
              "batch_size_pose", "max_iterations"):
        assert k in hp
    import lsps_b200
    ns = {}
    exec("from lsps_b200 import *", ns)
    assert ns[hp["trainer"]] is lsps_b200.LSPSTrainerB200


def test_shard_rows_and_source_assignment():
    from lsps_b200.sharding import shard_rows, source_assignment
    t = torch.arange(2 * 4 * 3).reshape(24, 1)          # 2 groups x world 4 x 3 rows
    got = torch.cat([shard_rows(t, 2, 4, r) for r in range(4)], 0).reshape(4, 2, 3)
    # rank r holds rows [r*3,(r+1)*3) of block 0 and of block 1
    for r in range(4):
        assert got[r, 0].tolist() == list(range(r * 3, r * 3 + 3))
        assert got[r, 1].tolist() == list(range(12 + r * 3, 12 + r * 3 + 3))
    for world in (1, 2, 4, 8):
        seen_a, seen_b = [], []
        for r in range(world):
            ka, kb = source_assignment(4, 4, world, r)
            seen_a += ka
            seen_b += kb
        assert sorted(seen_a) == [0, 1, 2, 3] and sorted(seen_b) == [0, 1, 2, 3]


def test_synthetic_dataset_item_contract():
    from lsps_b200.data import SyntheticHandDataset
    ds = SyntheticHandDataset(n=4, label_dim=48)
    img, label, com, M, cube, cube2 = ds[1]
    assert img.shape == (1, 128, 128) and img.dtype == torch.float32 and label.shape == (48,)
    assert float(img.max()) == 1.0 and float(img.min()) >= -1.0 and com.shape == (3,) and M.shape == (3, 3)
    assert torch.equal(ds[1][0], img)


def test_conv_weights_are_whole_gemm_rows_apart():
    """Grouped launches (engine.py) put ONE weight tensor map over two convs of the flat buffer: encoder-A/B and
    decoder-A/B res-block weights must be a multiple of 256 elements (one K row of the 256-channel GEMM) apart.
    Checked on the entry tables alone (no device needed): offsets are cumulative sizes rounded up to params.ALIGN."""
    from lsps_b200.params import gen_entries, ALIGN, _round_up
    import math
    assert ALIGN % 256 == 0
    hp = _hp("nnyu")
    off, offs = 0, {}
    for key, shape, kind, law, fan in gen_entries(hp["gen"]):
        offs[key] = off
        off += _round_up(int(math.prod(shape)))
    for i in range(3, 3 + hp["gen"]["n_enc_res_blk"]):
        for conv in ("model.0", "model.3"):
            d = offs["encode_B.%d.%s.weight" % (i, conv)] - offs["encode_A.%d.%s.weight" % (i, conv)]
            assert d > 0 and d % 256 == 0
    for i in range(hp["gen"]["n_gen_res_blk"]):
        d = offs["decode_B.%d.model.0.weight" % i] - offs["decode_A.%d.model.0.weight" % i]
        assert d % 256 == 0


def test_joint_schedule_steps_both_stores():
    from lsps_b200.params import MultiStepLR
    from lsps_b200.trainer import _JointSchedule

    class S(object):
        lr = base_lr = 1e-4
    a, b = S(), S()
    sch = _JointSchedule(MultiStepLR(a, [2], 0.5), MultiStepLR(b, [2], 0.5))
    sch.step(); sch.step()
    assert a.lr == b.lr == 5e-5 and sch.get_lr() == [5e-5]


def test_train_map_noise_sharding_matches_the_global_draw():
    """train_map: vae.encode runs on cat(labels_a, labels_b) (lsps_trainer.py:86-88); under data parallelism rank r holds
    cat(labels_a[r-th shard], labels_b[r-th shard]) and must use the rows of the reference's (2B, z) host draw that
    belong to exactly those samples: shard_rows(noise, groups=2).  Checked with the oracle's poseVAE arithmetic."""
    from lsps_b200.sharding import shard_rows
    hp = _hp("nnyu")
    vae = O.PoseVAE(hp["vae"], O.init_params(O.vae_spec(hp["vae"]), 3))
    g = torch.Generator().manual_seed(5)
    B, world = 6, 3
    la, lb = torch.randn(B, 108, generator=g), torch.randn(B, 108, generator=g)
    torch.manual_seed(11)
    z_ref = vae.encode(torch.cat((la, lb), 0))[0]                 # single process: one (2B, 20) draw
    torch.manual_seed(11)
    noise = torch.normal(torch.zeros(2 * B, 20), std=0.05)        # the same global draw, made on every rank
    for rank in range(world):
        la_r, lb_r = shard_rows(la, 1, world, rank), shard_rows(lb, 1, world, rank)
        y = torch.cat((la_r, lb_r), 0)
        P = vae.P
        h = torch.nn.functional.leaky_relu(torch.nn.functional.linear(y, P["en_fc1.weight"], P["en_fc1.bias"]), 0.01)
        mu = torch.nn.functional.linear(h, P["en_mu.weight"], P["en_mu.bias"])
        sd = torch.nn.functional.softplus(torch.nn.functional.linear(h, P["en_sigma.weight"], P["en_sigma.bias"]))
        z = mu + sd * shard_rows(noise, 2, world, rank)
        want = torch.cat((shard_rows(z_ref[:B], 1, world, rank), shard_rows(z_ref[B:], 1, world, rank)), 0)
        assert torch.allclose(z, want, atol=1e-6)
