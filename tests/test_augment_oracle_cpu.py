"""Groundwork for SURVEY 8f row n3 (input pipeline): the numpy oracle of normalize + augmentCrop must reproduce the
committed outputs of the UNMODIFIED reference pipeline (tests/golden/augment_*.npz, written by
oracle/make_augment_golden.py where /root/reference and its cv2 are available).  No product kernel exists for this row
yet; these tests only keep the definition the kernel will be held to from drifting."""
import os

import numpy as np
import pytest

import augment_oracle as A

MODES = (["com"], ["rot"], ["sc"], ["none"], ["com", "rot", "sc", "none"])


@pytest.mark.parametrize("mi", range(len(MODES)))
def test_augment_oracle_reproduces_reference(mi, golden_dir):
    modes = MODES[mi]
    gold = np.load(os.path.join(golden_dir, "augment_" + "_".join(modes) + ".npz"))
    assert str(gold["meta_modes"]) == ",".join(modes)
    cam = A.Camera(*A.NYU_CAMERA)
    c = 0
    while "c%d_img" % c in gold.files:
        rs = np.random.RandomState(int(gold["meta_seed0"]) + c)
        dpt, com, cube, M, gt = A.synthetic_crop(rs, cam)
        img = A.normalize(dpt.copy(), com, cube)
        ties = []
        o_img, o_lab, o_cube, o_com, o_M, o_rot = A.augment_crop(img, gt, com.copy(), cube.copy(), M.copy(), list(modes), cam,
                                                                 np.random.RandomState(int(gold["meta_rng0"]) + c), ties=ties)
        tie = np.unpackbits(gold["c%d_tie" % c]).astype(bool).reshape(128, 128)
        if ties:
            assert np.array_equal(ties[0], tie)
        ref = gold["c%d_img" % c]
        assert np.array_equal(o_img[~tie], ref[~tie]), (modes, c)       # bit-exact off the (measure-zero) tie pixels
        assert tie.mean() < 0.05      # worst case: the whole 1-pixel frame maps exactly onto the source edge
        assert -1.0 - 1e-6 <= o_img.min() and o_img.max() <= 1.0 + 1e-6      # float32 rounding of the scaled cube (same in the reference)
        np.testing.assert_allclose(np.asarray(o_lab, np.float32), gold["c%d_label" % c], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(np.asarray(o_cube, np.float32), gold["c%d_cube" % c], rtol=1e-6)
        np.testing.assert_allclose(np.asarray(o_com, np.float32), gold["c%d_com" % c], rtol=1e-6)
        np.testing.assert_allclose(o_M, gold["c%d_M" % c], rtol=1e-6, atol=1e-6)
        assert float(o_rot) == float(gold["c%d_rot" % c])
        c += 1
    assert c >= 6


def test_restated_warps_match_opencv_when_it_is_installed():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(3)
    for it in range(40):
        src = (rng.rand(128, 128) * 1000).astype(np.float32)
        s = rng.uniform(0.6, 1.6)
        M = np.array([[s, rng.uniform(-0.05, 0.05), rng.uniform(-40, 40)], [rng.uniform(-0.05, 0.05), s, rng.uniform(-40, 40)],
                      [rng.uniform(-1e-4, 1e-4), rng.uniform(-1e-4, 1e-4), 1.0]])
        ref = cv2.warpPerspective(src, M, (128, 128), flags=cv2.INTER_NEAREST, borderMode=cv2.BORDER_CONSTANT, borderValue=0.0)
        got, ties = A.warp_perspective_nn(src, M, (128, 128), 0.0, return_ties=True)
        assert np.array_equal(ref[~ties], got[~ties])
        ang = rng.uniform(-180, 180)
        R = cv2.getRotationMatrix2D((64, 64), ang, 1)
        assert np.array_equal(R, A.rotation_matrix_2d((64, 64), ang, 1))
        ref = cv2.warpAffine(src, R, (128, 128), flags=cv2.INTER_NEAREST, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
        assert np.array_equal(ref, A.warp_affine_nn(src, R, (128, 128), 0))


def test_normalize_maps_background_to_plus_one():
    cam = A.Camera(*A.NYU_CAMERA)
    dpt, com, cube, M, gt = A.synthetic_crop(np.random.RandomState(0), cam)
    img = A.normalize(dpt.copy(), com, cube)
    assert img.dtype == np.float32 and np.all(img[dpt == 0] == 1.0) and img[dpt != 0].max() < 1.0
