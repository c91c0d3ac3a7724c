"""Synthetic stand-in for the reference's hand datasets.

Item contract of /root/reference/src/data/dataset_hand2.py:352,366,406:
    (img float32 (1,128,128) in [-1,1] with background +1, label float32 (J*3,), com (3,), M (3,3), cube (3,), cube (3,))
"""
import torch


def synthetic_batch(batch, label_dim, generator, kind="uniform", device=None):
    """(ia, ib, la, lb) drawn from a CPU generator -- the seeded stream the parity tests feed to both sides."""
    def img():
        if kind == "uniform":
            return torch.rand(batch, 1, 128, 128, generator=generator) * 2 - 1
        x = torch.ones(batch, 1, 128, 128)
        yy, xx = torch.meshgrid(torch.arange(128.0), torch.arange(128.0), indexing="ij")
        for i in range(batch):
            ax = 30 + 20 * torch.rand(2, generator=generator)
            m = ((yy - 63.5) / ax[0]) ** 2 + ((xx - 63.5) / ax[1]) ** 2 <= 1.0
            v = (torch.randn(128, 128, generator=generator) * 0.35).clamp(-1, 1)
            x[i, 0][m] = v[m]
        return x
    ia, ib = img(), img()
    la = torch.randn(batch, label_dim, generator=generator) * 0.3
    lb = torch.randn(batch, label_dim, generator=generator) * 0.3
    out = (ia, ib, la, lb)
    if device is not None:
        out = tuple(t.to(device) for t in out)
    return out


class SyntheticHandDataset(torch.utils.data.Dataset):
    """6-tuples like dataset_hand_NYU; deterministic per index."""

    def __init__(self, specs=None, n=4096, label_dim=108, seed=23455, cube=300.0, camera_items=False):
        """camera_items: emit a plausible centre of mass (image coordinates + depth in mm) and its crop transform M instead
        of zeros / identity, so that the augmentation row (lsps_b200.augment.CropAugmenter) has something to move."""
        specs = specs or {}
        self.n, self.label_dim, self.seed, self.cube = n, label_dim, int(specs.get("seed", seed)), cube
        self.camera_items = camera_items

    def __len__(self):
        return self.n

    def set_nmax(self, frac):
        """dataset_hand2.py:202,368: keep the first `frac` of the labelled real samples."""
        self.n = max(1, int(self.n * frac))

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 1000003 + i)
        ia, _, la, _ = synthetic_batch(1, self.label_dim, g, kind="hand")
        cube = torch.full((3,), self.cube)
        if self.camera_items:
            import numpy as np
            from .augment import Camera, NYU_CAMERA, com_to_transform
            com = np.array([320.0 + 40.0 * float(torch.rand(1, generator=g)) - 20.0,
                            240.0 + 40.0 * float(torch.rand(1, generator=g)) - 20.0, 600.0 + 100.0 * float(torch.rand(1, generator=g))])
            M = com_to_transform(com, (self.cube,) * 3, Camera(*NYU_CAMERA))
            return ia[0], la[0], torch.from_numpy(com).float(), torch.from_numpy(np.asarray(M, np.float32)), cube, cube
        return ia[0], la[0], torch.zeros(3), torch.eye(3), cube, cube
