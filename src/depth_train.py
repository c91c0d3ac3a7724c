#!/usr/bin/env python
"""py3 restatement of the reference driver loop (/root/reference/src/depth_train.py:63-166) on synthetic datasets.

Same flags (--gpu --resume --frac --idx --config --mode --log), same name-based trainer selection
(`exec("trainer=%s(config.hyperparameters)")`, :99-102), same per-iteration call sequence and LR-milestone stepping.
Everything the reference does around the hot path with real datasets (importers, evaluation, HTML/JPEG dumps) is out
of scope.  Multi-GPU: launch with torch.distributed.run; each rank feeds its shard.
"""
import os
import sys
import time
from optparse import OptionParser

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lsps_b200 import *  # noqa: F401,F403  (name-based selection, like `from trainers import *`)
from lsps_b200.config import NetConfig
from lsps_b200.data import SyntheticHandDataset

parser = OptionParser()
parser.add_option('--gpu', type=int, help="gpu id", default=0)
parser.add_option('--resume', type=int, help="resume training?", default=0)
parser.add_option('--frac', type=float, help="fraction of real labels to use", default=1.)
parser.add_option('--idx', type=int, help="idx predtrain", default=-1)
parser.add_option('--config', type=str, help="net configuration")
parser.add_option('--mode', type=str, help="pretrain/estimate", default="pretrain")
parser.add_option('--log', type=str, help="log path", default="../logs")
parser.add_option('--batch', type=int, help="(new) per-GPU batch override for pretrain; reference hard-codes 1", default=0)
parser.add_option('--snapshot_prefix', type=str, help="(new) overrides the YAML snapshot_prefix", default="")
parser.add_option('--iters', type=int, help="(new) stop after this many iterations", default=0)
parser.add_option('--noise', type=str, help="(new) host = reference RNG stream, device = Philox", default="device")


def main(argv):
    (opts, args) = parser.parse_args(argv)
    config = NetConfig(opts.config)
    hp = config.hyperparameters
    if opts.snapshot_prefix:
        config.snapshot_prefix = opts.snapshot_prefix
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(opts.gpu)))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank() if world > 1 else 0
    estimate = 'estimate' in opts.mode
    mode_idx = int(opts.mode[len('estimate'):] or 0) if estimate else 0
    batch_size = hp['batch_size'] if estimate else (opts.batch or 1)       # depth_train.py:85
    max_iterations = opts.iters or hp['max_iterations']
    label_dim = hp['vae']['input_dim']
    mk = lambda spec: SyntheticHandDataset(config.datasets[spec], label_dim=label_dim)
    dataset_a, dataset_b = mk('train_a'), mk('train_b')
    kw = dict(batch_size=batch_size, shuffle=True, num_workers=0, drop_last=True)
    sampler = lambda ds: torch.utils.data.distributed.DistributedSampler(ds) if world > 1 else None
    la_, lb_ = sampler(dataset_a), sampler(dataset_b)
    loader_a = torch.utils.data.DataLoader(dataset_a, sampler=la_, **dict(kw, shuffle=la_ is None))
    loader_b = torch.utils.data.DataLoader(dataset_b, sampler=lb_, **dict(kw, shuffle=lb_ is None))

    ns = dict(globals(), config=config)
    exec("trainer=%s(config.hyperparameters, device=%d, noise=%r)" % (hp['trainer'], local, opts.noise), ns)
    trainer = ns['trainer']
    iterations = 0
    if opts.resume == 1:
        iterations = trainer.resume(config.snapshot_prefix, idx=-1, load_opt=True)
        for _ in range(iterations // 1000):
            trainer.dis_sch.step()
            trainer.gen_sch.step()
    trainer.cuda(local)
    try:
        trainer.load_vae(config.snapshot_prefix, 2 + opts.frac if (estimate and mode_idx in (3, 4)) else opts.frac)
    except Exception:  # noqa  (reference: bare except + print, depth_train.py:118-124)
        if rank == 0:
            print('Failed to load the parameters of vae')
    if estimate:
        # depth_train.py:126-130 -- the estimate phases start from the pretrained generator / discriminator
        # (`--idx` selects the snapshot, default -1 = newest; 0 = start from scratch); mode 5 resumes a *_est run
        if opts.idx != 0:
            trainer.resume(config.snapshot_prefix, idx=opts.idx, est=(mode_idx == 5))
        if 0. < opts.frac < 1. and hasattr(dataset_b, 'set_nmax'):
            dataset_b.set_nmax(opts.frac)
    start_time = time.time()
    while iterations < max_iterations:
        for (images_a, labels_a, com_a, _, _, _), (images_b, labels_b, com_b, _, _, _) in zip(loader_a, loader_b):
            images_a, images_b = images_a.cuda(local, non_blocking=True), images_b.cuda(local, non_blocking=True)
            labels_a, labels_b = labels_a.cuda(local, non_blocking=True), labels_b.cuda(local, non_blocking=True)
            trainer.dis.train()
            if opts.mode == 'pretrain':
                if (iterations + 1) % 1000 == 0:
                    trainer.dis_sch.step()
                    trainer.gen_sch.step()
                trainer.dis_update(images_a, labels_a, images_b, labels_b, com_a, com_b, hp)
                image_outputs = trainer.gen_update(images_a, labels_a, images_b, labels_b, hp)
            else:
                if (iterations + 1) % 100 == 0:
                    trainer.dis_sch.step()
                image_outputs = trainer.post_update(images_a, labels_a, images_b, labels_b, com_a, com_b, mode_idx, hp)
            trainer.assemble_outputs(images_a, images_b, image_outputs)
            if (iterations + 1) % config.display == 0 and rank == 0:
                members = [a for a in dir(trainer) if not callable(getattr(trainer, a)) and ('loss' in a or 'acc' in a)]
                print("Iteration: %08d/%08d  %.2fs  %s" % (iterations + 1, max_iterations, time.time() - start_time,
                                                          " ".join("%s=%.4f" % (m, float(getattr(trainer, m))) for m in members)))
                start_time = time.time()
            if (iterations + 1) % config.snapshot_save_iterations == 0 and rank == 0:
                os.makedirs(os.path.dirname(config.snapshot_prefix) or ".", exist_ok=True)
                trainer.save(config.snapshot_prefix + ('_est' if estimate else ''), iterations)
            iterations += 1
            if iterations >= max_iterations:
                break
    if opts.iters and rank == 0 and opts.iters % config.snapshot_save_iterations != 0:
        # (new) short runs (--iters) leave a snapshot so that the next phase can chain onto them
        os.makedirs(os.path.dirname(config.snapshot_prefix) or ".", exist_ok=True)
        trainer.save(config.snapshot_prefix + ('_est' if estimate else ''), iterations - 1)
    if world > 1:
        dist.destroy_process_group()
    return trainer


if __name__ == '__main__':
    main(sys.argv)
