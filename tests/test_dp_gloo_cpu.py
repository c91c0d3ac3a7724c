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
