"""Evaluation sweep of the estimate phases on the device.

Restates the driver's test loop (/root/reference/src/depth_train.py:186-253): images -> dis.regress_a/b -> vae.decode ->
joints * cube/2 + com -> mean joint error (mm) and the share of frames whose WORST joint is within 40 mm
(/root/reference/src/utils/handpose_evaluation.py:92-97, 197-203).  The per-frame errors are computed by one kernel
(lsps_joint_errors); only two scalars per sweep come back to the host.  `com` cancels in gt - pred and is not needed.
"""
import torch

from . import _lib

# depth_train.py:229 -- the 14 joints the NYU benchmark evaluates, out of the 36 annotated ones
NYU_RESTRICTED_JOINTS = (0, 3, 6, 9, 12, 15, 18, 21, 24, 25, 27, 30, 31, 32)


class PoseEvaluator(object):
    def __init__(self, trainer, domain="b", restricted_joints=None):
        self.tr, self.domain = trainer, domain
        self.ctx = trainer.ops.ctx
        self.jidx = None
        if restricted_joints is not None:
            self.jidx = torch.tensor(list(restricted_joints), dtype=torch.int32, device=trainer.device)
        self._mean, self._max = [], []

    def reset(self):
        self._mean, self._max = [], []

    def add_batch(self, images, labels, cube):
        """images (n,1,128,128); labels (n, J*3) normalised joints; cube (n,3) or (3,) in mm (the driver scales every
        frame of a batch by the FIRST frame's cube, depth_train.py:235)."""
        tr = self.tr
        was_training = tr.dis.training
        tr.dis.eval()
        post = (tr.dis.regress_a if self.domain == "a" else tr.dis.regress_b)(images)[1]
        pose = tr.vae.decode(post.reshape(images.shape[0], -1))
        tr.dis.train(was_training)
        gt = labels.detach().to(device=tr.device, dtype=torch.float32).contiguous()
        n, j3 = gt.shape
        c = torch.as_tensor(cube, dtype=torch.float32).reshape(-1, 3)[0] / 2.0
        nj = self.jidx.numel() if self.jidx is not None else j3 // 3
        emean = torch.empty(n, dtype=torch.float32, device=tr.device)
        emax = torch.empty(n, dtype=torch.float32, device=tr.device)
        self.ctx.joint_errors(pose.data_ptr(), gt.data_ptr(), _lib.ptr(self.jidx), nj, j3, float(c[0]), float(c[1]), float(c[2]),
                              emean.data_ptr(), emax.data_ptr(), n)
        self._mean.append(emean)
        self._max.append(emax)
        return pose

    def summary(self, dist=40.0):
        """-> (mean error in mm, percentage of frames whose maximum joint error is <= dist)."""
        m, x = torch.cat(self._mean), torch.cat(self._max)
        out = torch.stack((m.mean(), (x <= dist).float().mean() * 100.0)).cpu()
        return float(out[0]), float(out[1])
