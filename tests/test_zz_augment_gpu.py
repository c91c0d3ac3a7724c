"""Row n3 on the device: lsps_b200.augment.CropAugmenter (host parameter draw + ONE lsps_augment_crops launch for the
batch) against the pinned numpy oracle of the reference pipeline -- every pixel bit-exact (same arithmetic definition,
tests/test_augment_core_cpu.py checks the host build of the same function), labels / com / cube / M identical.
(File name sorts last on purpose: this launch could not be exercised on hardware in round 1.)"""
import numpy as np
import pytest
import torch

import augment_oracle as A

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("modes", [["com", "rot", "sc", "none"], ["rot"], ["sc"]])
def test_crop_augmenter_matches_oracle(modes):
    from lsps_b200.augment import CropAugmenter
    cam = A.Camera(*A.NYU_CAMERA)
    n = 48
    imgs = np.empty((n, 1, 128, 128), np.float32)
    items = []
    for c in range(n):
        dpt, com, cube, M, gt = A.synthetic_crop(np.random.RandomState(7000 + c), cam)
        imgs[c, 0] = A.normalize(dpt.copy(), com, cube)
        items.append((gt, com, cube, M))
    aug = CropAugmenter(aug_modes=modes, seed=321)
    out, labels, cubes, coms, Ms, rots = aug(torch.from_numpy(imgs).cuda(), [i[0] for i in items], [i[1] for i in items],
                                            [i[2] for i in items], [i[3] for i in items])
    assert out.shape == (n, 1, 128, 128) and out.is_cuda
    out = out.cpu().numpy()
    rng = np.random.RandomState(321)
    for c in range(n):
        gt, com, cube, M = items[c]
        o_img, o_lab, o_cube, o_com, o_M, o_rot = A.augment_crop(imgs[c, 0].copy(), gt.copy(), com.copy(), cube.copy(), M.copy(),
                                                                 list(modes), cam, rng)
        assert np.array_equal(out[c, 0], o_img), (c, int((out[c, 0] != o_img).sum()))
        assert np.array_equal(np.asarray(labels[c], np.float32), np.asarray(o_lab, np.float32))
        assert np.array_equal(np.asarray(cubes[c], np.float32), np.asarray(o_cube, np.float32))
        assert np.array_equal(np.asarray(coms[c], np.float32), np.asarray(o_com, np.float32))
        assert np.array_equal(Ms[c], o_M) and float(rots[c]) == float(o_rot)
