"""Batch sharding rules of the data-parallel step (pure host logic, importable without CUDA).

The path shards over batch: rank r owns samples [r*B, (r+1)*B) of BOTH domains.  Tensors the
reference builds by concatenating `groups` per-domain blocks along batch (e.g. the (2B,256,32,32) latent noise of
gen.forward = [domain-a block | domain-b block]) are sharded block-wise so that every rank sees exactly the rows
of its own samples."""
import torch
import torch.distributed as dist


def shard_rows(t, groups, world, rank):
    """t: [groups * world * per, ...] global tensor -> [groups * per, ...] rows of `rank`."""
    if world == 1:
        return t
    assert t.shape[0] % (groups * world) == 0, (t.shape, groups, world)
    per = t.shape[0] // (groups * world)
    return torch.cat([t[g * per * world + rank * per: g * per * world + (rank + 1) * per] for g in range(groups)], 0)


def source_assignment(n_a, n_b, world, rank):
    """post_update modes >= 2 run the generator on the GLOBAL first 4 samples of each domain
    (/root/reference/src/trainers/lsps_trainer.py:238).  Source image a_i goes to rank i % world, b_i to rank
    (n_a + i) % world; InstanceNorm is per-sample, so the split is exact."""
    ka = [i for i in range(n_a) if i % world == rank]
    kb = [i for i in range(n_b) if (n_a + i) % world == rank]
    return ka, kb


def global_count(local, world):
    """Every loss mean is normalised by the global element count so that sum-allreduced gradients equal the
    single-process gradients of the global batch."""
    return local * world


def world_rank():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def allreduce_sum_(flat):
    """The step's one collective: in-place sum over ranks of a flat buffer (gradients + loss sums).  NCCL on the GPUs,
    gloo in the CPU tests; a no-op for a single process."""
    if world_rank()[0] > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return flat


def feature_sources(n_a, n_b, world, rank):
    """(ka, kb, idx): the source images of post_update's feature-matching sub-graph this rank runs the generator on, and
    their rows in the reference's (n_a + n_b)-row latent-noise draw (domain-a rows first)."""
    ka, kb = source_assignment(n_a, n_b, world, rank)
    return ka, kb, ka + [n_a + i for i in kb]
