#!/usr/bin/env python
"""py3 restatement of /root/reference/src/pose_train.py:63-190 (phase 1: pose-VAE) on synthetic labels.
Loop of :121-135: labels = cat(labels_a, labels_b); vae_update; vae_sch.step() every 1000 iterations."""
import os
import sys
import time
from optparse import OptionParser

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lsps_b200 import *  # noqa: F401,F403
from lsps_b200.config import NetConfig
from lsps_b200.data import SyntheticHandDataset

parser = OptionParser()
parser.add_option('--gpu', type=int, help="gpu id", default=0)
parser.add_option('--resume', type=int, help="resume training?", default=0)
parser.add_option('--frac', type=float, help="fraction of real labels to use", default=1.)
parser.add_option('--config', type=str, help="net configuration")
parser.add_option('--log', type=str, help="log path", default="../logs")
parser.add_option('--snapshot_prefix', type=str, help="(new) overrides the YAML snapshot_prefix", default="")
parser.add_option('--iters', type=int, help="(new) stop after this many iterations", default=0)
parser.add_option('--batch', type=int, help="(new) per-domain batch override (default hyperparameters.batch_size_pose)", default=0)
parser.add_option('--noise', type=str, default="host")


def main(argv):
    (opts, args) = parser.parse_args(argv)
    config = NetConfig(opts.config)
    hp = config.hyperparameters
    if opts.snapshot_prefix:
        config.snapshot_prefix = opts.snapshot_prefix
    torch.cuda.set_device(opts.gpu)
    batch_size = opts.batch or hp['batch_size_pose']
    max_iterations = opts.iters or hp['max_iterations']
    label_dim = hp['vae']['input_dim']
    mk = lambda spec: SyntheticHandDataset(config.datasets[spec], label_dim=label_dim)
    kw = dict(batch_size=batch_size, shuffle=True, num_workers=0, drop_last=True)
    loader_a = torch.utils.data.DataLoader(mk('train_a'), **kw)
    loader_b = torch.utils.data.DataLoader(mk('train_b'), **kw)
    ns = dict(globals(), config=config)
    exec("trainer=%s(config.hyperparameters, device=%d, noise=%r)" % (hp['trainer'], opts.gpu, opts.noise), ns)
    trainer = ns['trainer']
    trainer.cuda(opts.gpu)
    iterations, start_time = 0, time.time()
    while iterations < max_iterations:
        for (_, labels_a, _, _, _, _), (_, labels_b, _, _, _, _) in zip(loader_a, loader_b):
            # pose_train.py:121-133: domain-a labels alone when no real labels are used (frac == 0)
            labels = (torch.cat((labels_a, labels_b), 0) if opts.frac > 0. else labels_a).cuda(opts.gpu)
            if (iterations + 1) % 1000 == 0:
                trainer.vae_sch.step()
            trainer.vae_update(labels, hp)
            if (iterations + 1) % config.display == 0:
                print("Iteration: %08d/%08d  %.2fs  vae_total_loss=%.5f" % (iterations + 1, max_iterations,
                                                                           time.time() - start_time, float(trainer.vae_total_loss)))
                start_time = time.time()
            if (iterations + 1) % (4 * config.snapshot_save_iterations) == 0:      # pose_train.py:183-184
                os.makedirs(os.path.dirname(config.snapshot_prefix) or ".", exist_ok=True)
                trainer.save_vae(config.snapshot_prefix, iterations, 2 + opts.frac)
            iterations += 1
            if iterations >= max_iterations:
                break
    if opts.iters and opts.iters % (4 * config.snapshot_save_iterations) != 0:
        # (new) short runs (--iters) still hand their pose-VAE to depth_train.py's load_vae
        os.makedirs(os.path.dirname(config.snapshot_prefix) or ".", exist_ok=True)
        trainer.save_vae(config.snapshot_prefix, iterations - 1, 2 + opts.frac)
    return trainer


if __name__ == '__main__':
    main(sys.argv)
