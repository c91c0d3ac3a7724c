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
parser.add_option('--augment', type=int, help="(new) 1 = online crop augmentation on the GPU (lsps_augment_crops) in place "
                  "of the DataLoader workers' augmentCrop (data/dataset_hand2.py:34-119)", default=0)
parser.add_option('--eval_every', type=int, help="(new) estimate modes: device evaluation sweep over test_b every N "
                  "iterations (depth_train.py:186-253 runs it every image_save_iterations)", default=0)


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
    mk = lambda spec: SyntheticHandDataset(config.datasets[spec], label_dim=label_dim, camera_items=bool(opts.augment))
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
    augmenters, evaluator, test_loader = None, None, None
    if opts.augment:
        # the reference augments per sample inside its DataLoader workers (cfg key datasets.*.augment); here the host only
        # draws the random numbers in the reference's order and ONE kernel launch warps the whole batch on the GPU
        from lsps_b200.augment import CropAugmenter
        augmenters = [CropAugmenter(seed=int(config.datasets[k].get('seed', 23455)) + rank, device=local)
                      for k in ('train_a', 'train_b')]
    if estimate and opts.eval_every:
        from lsps_b200.evaluation import PoseEvaluator, NYU_RESTRICTED_JOINTS
        evaluator = PoseEvaluator(trainer, domain="b", restricted_joints=NYU_RESTRICTED_JOINTS if label_dim == 108 else None)
        test_loader = torch.utils.data.DataLoader(mk('test_b'), batch_size=32 * batch_size, shuffle=False, num_workers=0)

    def augment(aug, images, labels, com, M, cube):
        import numpy as np
        n = images.shape[0]
        c = cube.numpy()
        gt3d = [(labels[i].numpy().reshape(-1, 3) * (c[i][2] / 2.)).astype(np.float32) for i in range(n)]
        out, labs, cubes, coms, _, _ = aug(images.cuda(local, non_blocking=True), gt3d, [com[i].numpy().astype(np.float64) for i in range(n)],
                                           [tuple(c[i]) for i in range(n)], [M[i].numpy() for i in range(n)])
        lab = np.stack([np.asarray(labs[i], np.float32).reshape(-1) for i in range(n)])   # already normalised by the new cube
        return out, torch.from_numpy(lab), torch.from_numpy(np.stack([np.asarray(x, np.float32) for x in coms]))

    start_time = time.time()
    while iterations < max_iterations:
        for (images_a, labels_a, com_a, M_a, cube_a, _), (images_b, labels_b, com_b, M_b, cube_b, _) in zip(loader_a, loader_b):
            if augmenters is not None:
                images_a, labels_a, com_a = augment(augmenters[0], images_a, labels_a, com_a, M_a, cube_a)
                images_b, labels_b, com_b = augment(augmenters[1], images_b, labels_b, com_b, M_b, cube_b)
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
            if evaluator is not None and (iterations + 1) % opts.eval_every == 0:
                evaluator.reset()                                    # depth_train.py:186-253: regress_b -> vae.decode -> mm
                for (ti, tl, _, _, tcube, _) in test_loader:
                    evaluator.add_batch(ti.cuda(local, non_blocking=True), tl, tcube[0])
                    break                                            # synthetic test set: one batch of 32 x batch_size
                mean_err, within = evaluator.summary(40.0)
                if rank == 0:
                    print("Mean err %.2f mm, %.1f %% of frames within 40 mm" % (mean_err, within))
                trainer.last_eval = (mean_err, within)
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
