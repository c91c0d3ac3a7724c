"""N-GPU == 1-GPU equivalence of the data-parallel step (run under torch.distributed.run, NCCL).

Every rank builds the same trainer (same weights), takes ITS shard of one global batch and the rows of the
reference's host noise that belong to its samples, and runs dis_update + gen_update + post_update(mode 3).  Rank 0
then repeats the step alone on the GLOBAL batch (fresh trainer, same seeds) and compares losses and post-step weights."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lsps_b200  # noqa
from lsps_b200.sharding import shard_rows  # noqa

KEYS = ("dis_loss", "dis_ad_loss", "dis_feat_loss", "gen_total_loss", "gen_ad_loss", "gen_ll_loss", "gen_ll_loss2",
        "gen_enc_loss", "gen_enc_loss2", "dis_total_loss", "dis_reg_loss", "vae_total_loss")


def run(tr, hp, ia, ib, la, lb):
    torch.manual_seed(42)
    tr.dis_update(ia, la, ib, lb, None, None, hp)
    tr.gen_update(ia, la, ib, lb, hp)
    tr.post_update(ia, la, ib, lb, None, None, 3, hp)
    tr.vae_update(torch.cat((la, lb), 0), hp)
    return {k: float(getattr(tr, k)) for k in KEYS}


MAP_KEYS = ("dis_loss", "dis_ad_loss", "gen_total_loss", "gen_ad_loss", "gen_ll_loss", "gen_map_loss", "gen_map_loss2")


def check_train_map(world, rank, local):
    """Same equivalence for the train_map=True branches: dis_update + gen_update with the Mapping net (its own flat
    gradient buffer and allreduce; the vae.encode noise is drawn for the (labels_a | labels_b) concatenation)."""
    hp = dict(lsps_b200.load_hyperparameters("nnyu"), train_map=True)
    per = 2
    g = torch.Generator().manual_seed(4321)
    ia, ib, la, lb = lsps_b200.synthetic_batch(per * world, 108, g, "hand")
    sh = lambda t: shard_rows(t, 1, world, rank).cuda()

    def run_map(tr, a, b, x, y):
        torch.manual_seed(43)
        tr.dis_update(a, x, b, y, None, None, hp)
        tr.gen_update(a, x, b, y, hp)
        return {k: float(getattr(tr, k)) for k in MAP_KEYS}
    tr = lsps_b200.LSPSTrainerB200(hp, device=local, seed=0, noise="host")
    got = run_map(tr, sh(ia), sh(ib), sh(la), sh(lb))
    w_dp = {k: v.clone() for k, v in tr.map_store.state_dict().items()}
    dist.barrier()
    ok = True
    if rank == 0:
        import lsps_b200.trainer as T
        saved = T._world
        T._world = lambda: (1, 0)
        tr1 = lsps_b200.LSPSTrainerB200(hp, device=local, seed=0, noise="host")
        ref = run_map(tr1, ia.cuda(), ib.cuda(), la.cuda(), lb.cuda())
        for k in MAP_KEYS:
            rel = abs(got[k] - ref[k]) / (abs(ref[k]) + 1e-12)
            print("train_map %-16s dp %.6f  single %.6f  rel %.2e" % (k, got[k], ref[k], rel))
            ok &= rel < 2e-3
        w1 = tr1.map_store.state_dict()
        # yardstick: the SAME single-process step run a second time.  The InstanceNorm statistics are accumulated with
        # fp32 atomics in the conv epilogues (order varies run to run), which moves isolated bf16 roundings; the first Adam
        # step (~ lr * sign(g)) turns that into sign flips where g ~ 0.  DP must be as close to single as single is to itself.
        tr2 = lsps_b200.LSPSTrainerB200(hp, device=local, seed=0, noise="host")
        run_map(tr2, ia.cuda(), ib.cuda(), la.cuda(), lb.cuda())
        w2 = tr2.map_store.state_dict()
        w0 = lsps_b200.LSPSTrainerB200(hp, device=local, seed=0, noise="host").map_store.state_dict()
        cosf = lambda a, b, k: torch.nn.functional.cosine_similarity((a[k] - w0[k]).reshape(1, -1), (b[k] - w0[k]).reshape(1, -1)).item()
        for k in ("model.0.model.0.weight", "model.1.model.0.weight", "model.3.weight", "model.3.bias"):
            cos, self_cos = cosf(w_dp, w1, k), cosf(w2, w1, k)
            print("train_map weight update %-24s cosine(dp, single) %.6f   cosine(single rerun, single) %.6f" % (k, cos, self_cos))
            ok &= cos > min(0.98, self_cos - 0.03)
        T._world = saved
    return ok


def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    hp = lsps_b200.load_hyperparameters("nnyu")
    per = 4
    g = torch.Generator().manual_seed(1234)
    ia, ib, la, lb = lsps_b200.synthetic_batch(per * world, 108, g, "uniform")
    tr = lsps_b200.LSPSTrainerB200(hp, device=local, seed=0, noise="host")
    sh = lambda t: shard_rows(t, 1, world, rank).cuda()
    # vae_update concatenates (la, lb): two blocks -> the trainer shards its noise block-wise itself (groups=1 on the
    # concatenated local rows), so feed the local rows of each block
    got = run(tr, hp, sh(ia), sh(ib), sh(la), sh(lb))
    w_dp = {k: v.clone() for k, v in tr.dis_store.state_dict().items()}
    dist.barrier()
    ok = True
    if rank == 0:
        # single-process reference on the global batch: temporarily hide the process group from the trainer
        import lsps_b200.trainer as T
        saved = T._world
        T._world = lambda: (1, 0)
        tr1 = lsps_b200.LSPSTrainerB200(hp, device=local, seed=0, noise="host")
        ref = run(tr1, hp, ia.cuda(), ib.cuda(), la.cuda(), lb.cuda())
        T._world = saved
        for k in KEYS:
            if k == "vae_total_loss":
                continue   # vae noise rows are drawn for the (la|lb) concatenation: different row order under DP
            rel = abs(got[k] - ref[k]) / (abs(ref[k]) + 1e-12)
            print("%-16s dp %.6f  single %.6f  rel %.2e" % (k, got[k], ref[k], rel))
            ok &= rel < 2e-3
        w1 = tr1.dis_store.state_dict()
        w0 = lsps_b200.LSPSTrainerB200(hp, device=local, seed=0, noise="host").dis_store.state_dict()
        for k in ("model_S.3.model.0.weight", "Post.weight", "model_A.0.model.0.weight"):
            cos = torch.nn.functional.cosine_similarity((w_dp[k] - w0[k]).reshape(1, -1),
                                                        (w1[k] - w0[k]).reshape(1, -1)).item()
            print("weight update %-28s cosine(dp, single) %.6f" % (k, cos))
            ok &= cos > 0.98
    del tr
    ok = check_train_map(world, rank, local) and ok
    if rank == 0:
        print("DP_CHECK", "OK" if ok else "FAIL")
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
