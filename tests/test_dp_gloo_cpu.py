"""world_size-2 `gloo` test of the data-parallel rules (host logic; no GPU): sharding of the batch and of the
reference's host noise stream, global-count loss normalisation, ONE sum-allreduce of the flat gradient buffer."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import yaml

import lsps_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from lsps_b200.sharding import shard_rows
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    with open(os.path.join(ROOT, "exps", "nnyu.yaml")) as fh:
        hp = yaml.safe_load(fh)["train"]["hyperparameters"]
    tr = O.OracleTrainer(hp, seed=0)
    # pose-VAE step on the global batch of 16 rows, rank r owns rows [8r, 8r+8)
    g = torch.Generator().manual_seed(1234)
    y = torch.randn(16, 108, generator=g) * 0.3
    torch.manual_seed(42)
    noise = torch.normal(torch.zeros(16, 20), std=0.05)         # the reference's global host draw
    y_loc, n_loc = shard_rows(y, 1, world, rank), shard_rows(noise, 1, world, rank)
    P = tr.params["vae"]
    import torch.nn.functional as F
    h = F.leaky_relu(F.linear(y_loc, P["en_fc1.weight"], P["en_fc1.bias"]), 0.01)
    mu = F.linear(h, P["en_mu.weight"], P["en_mu.bias"])
    sd = F.softplus(F.linear(h, P["en_sigma.weight"], P["en_sigma.bias"]))
    dec = tr.vae.decode(mu + sd * n_loc)
    rows = y_loc.shape[0] * world                                # GLOBAL count normalisation
    loss = hp["kl_loss_vae"] * (mu * mu + sd * sd - torch.log(sd * sd)).sum() / rows + \
        hp["ll_loss_vae"] * (dec - y_loc).abs().sum() / (rows * 108)
    loss.backward()
    flat = torch.cat([p.grad.reshape(-1) for p in P.values()] + [loss.detach().reshape(1)])
    dist.all_reduce(flat)                                        # the ONE collective of the update
    if rank == 0:
        torch.save(flat, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_vae_step_equals_single_process(tmp_path):
    out = str(tmp_path / "flat.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    flat = torch.load(out)
    with open(os.path.join(ROOT, "exps", "nnyu.yaml")) as fh:
        hp = yaml.safe_load(fh)["train"]["hyperparameters"]
    tr = O.OracleTrainer(hp, seed=0)
    g = torch.Generator().manual_seed(1234)
    y = torch.randn(16, 108, generator=g) * 0.3
    torch.manual_seed(42)
    tr._zero("vae")
    dec, z, mu, sd = tr.vae.forward(y)
    total = hp["kl_loss_vae"] * tr._kl(mu, sd) + hp["ll_loss_vae"] * torch.nn.functional.l1_loss(dec, y)
    total.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in tr.params["vae"].values()] + [total.detach().reshape(1)])
    assert torch.allclose(flat, ref, rtol=1e-4, atol=1e-6), (flat - ref).abs().max()


# ------------------------------------------------------------------ estimate3: the product's sharding rules end to end
def _worker_estimate3(rank, world, port, out):
    """Each rank runs ITS part of post_update(mode 3) with the PRODUCT's host logic (lsps_b200.sharding: batch shard,
    noise-row selection, the a_i -> rank i / b_i -> rank 4+i source assignment of the global images[0:4], global-count
    normalisation, the one sum-allreduce) around the oracle's arithmetic."""
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from lsps_b200.sharding import shard_rows, feature_sources, allreduce_sum_, world_rank
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert world_rank() == (world, rank)
    torch.set_num_threads(4)
    with open(os.path.join(ROOT, "exps", "nnyu.yaml")) as fh:
        hp = yaml.safe_load(fh)["train"]["hyperparameters"]
    tr = O.OracleTrainer(hp, seed=0)
    Bg = 4
    g = torch.Generator().manual_seed(1234)
    ia, ib, la, lb = O.synthetic_batch(Bg, 108, g, "uniform")
    torch.manual_seed(42)
    lat_noise = torch.randn(8, 256, 32, 32)                     # the reference's global draws, in its order
    vae_noise = torch.normal(torch.zeros(Bg, 20), std=0.05)
    # feature-matching sub-graph on the sources assigned to this rank
    ka, kb, idx = feature_sources(4, 4, world, rank)
    real_randn = torch.randn
    torch.randn = lambda *a, **k: lat_noise[idx].clone()        # Gen._enc_shared draws torch.randn(x.size())
    try:
        x_aa, x_ba, x_ab, x_bb, _ = tr.gen.forward(ia[0:4][ka], ib[0:4][kb])
    finally:
        torch.randn = real_randn
    x_aa, x_ba, x_ab, x_bb = (t.detach() for t in (x_aa, x_ba, x_ab, x_bb))
    fa = tr.dis.trunk(tr.dis.front("A", torch.cat((x_aa, x_ba), 0)))
    fb = tr.dis.trunk(tr.dis.front("B", torch.cat((x_ab, x_bb), 0)))
    na = len(ka)
    per = fa[0].numel()
    feat = ((fb[:na] - fa[:na]).abs().sum() + (fa[na:] - fb[na:]).abs().sum()) / (4 * per)      # GLOBAL count: 4 images
    # regression on this rank's shard of domain a, with its rows of the vae.encode draw
    ia_l, la_l, nz = shard_rows(ia, 1, world, rank), shard_rows(la, 1, world, rank), shard_rows(vae_noise, 1, world, rank)
    P = tr.params["vae"]
    import torch.nn.functional as F
    h = F.leaky_relu(F.linear(la_l, P["en_fc1.weight"], P["en_fc1.bias"]), 0.01)
    enc = F.linear(h, P["en_mu.weight"], P["en_mu.bias"]) + F.softplus(F.linear(h, P["en_sigma.weight"], P["en_sigma.bias"])) * nz
    pred = tr.dis.regress("A", ia_l)
    reg = ((pred - enc.detach()) ** 2).sum() / (Bg * 20)                                        # GLOBAL count
    loss = hp["reg_w"] * reg + hp["feature_w_reg"] * feat
    loss.backward()
    keys = [k for k, p in tr.params["dis"].items() if p.grad is not None]
    flat = torch.cat([tr.params["dis"][k].grad.reshape(-1) for k in keys] + [loss.detach().reshape(1)])
    allreduce_sum_(flat)
    if rank == 0:
        torch.save((keys, flat), out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_estimate3_step_equals_single_process(tmp_path):
    out = str(tmp_path / "flat3.pt")
    mp.spawn(_worker_estimate3, args=(2, _free_port(), out), nprocs=2, join=True)
    keys, flat = torch.load(out)
    with open(os.path.join(ROOT, "exps", "nnyu.yaml")) as fh:
        hp = yaml.safe_load(fh)["train"]["hyperparameters"]
    tr = O.OracleTrainer(hp, seed=0)
    g = torch.Generator().manual_seed(1234)
    ia, ib, la, lb = O.synthetic_batch(4, 108, g, "uniform")
    torch.manual_seed(42)
    tr.post_update(ia, la, ib, lb, None, None, 3, hp)
    ref_keys = [k for k, p in tr.params["dis"].items() if p.grad is not None]
    assert keys == ref_keys
    ref = torch.cat([tr.params["dis"][k].grad.reshape(-1) for k in ref_keys] + [torch.tensor([tr.dis_total_loss])])
    assert abs(flat[-1].item() - ref[-1].item()) < 1e-5 * abs(ref[-1].item())
    rel = ((flat[:-1] - ref[:-1]).norm() / ref[:-1].norm()).item()
    assert rel < 1e-4, rel
