"""Row n3 without a GPU: the per-pixel function of the augmentation kernel (lsps_b200/csrc/augment_core.h) is compiled
for the host (tests/augment_host_harness.cpp, -ffp-contract=off) and driven with the product's own host-side parameter
code (lsps_b200/augment.py `sample_params`).  It must reproduce the pinned numpy oracle BIT FOR BIT on every pixel, and the
labels / com / cube / M the oracle (and therefore the reference) returns.  The device launch itself is not covered here."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import augment_oracle as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODES = (["com"], ["rot"], ["sc"], ["none"], ["com", "rot", "sc", "none"])


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("aug") / "aug_host.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "augment_host_harness.cpp")])
    lib = C.CDLL(so)
    lib.aug_host_sample_bytes.restype = C.c_int
    return lib


def test_param_record_layout_matches_the_c_struct(harness):
    from lsps_b200 import _lib
    from lsps_b200.augment import AugSample
    assert C.sizeof(AugSample) == harness.aug_host_sample_bytes() == _lib._lib.lsps_aug_sample_bytes()


@pytest.mark.parametrize("mi", range(len(MODES)))
def test_kernel_arithmetic_equals_oracle(harness, mi):
    from lsps_b200.augment import AugSample, Camera, sample_params
    modes = MODES[mi]
    cam_o, cam_p = A.Camera(*A.NYU_CAMERA), Camera(*A.NYU_CAMERA)
    n = 40
    imgs = np.empty((n, 128, 128), np.float32)
    recs = (AugSample * n)()
    want, meta = [], []
    for c in range(n):
        rs = np.random.RandomState(5000 + 100 * mi + c)
        dpt, com, cube, M, gt = A.synthetic_crop(rs, cam_o)
        imgs[c] = A.normalize(dpt.copy(), com, cube)
        o = A.augment_crop(imgs[c].copy(), gt.copy(), com.copy(), cube.copy(), M.copy(), list(modes), cam_o,
                           np.random.RandomState(90 + c))
        p = sample_params(gt.copy(), com.copy(), cube.copy(), M.copy(), list(modes), cam_p, np.random.RandomState(90 + c))
        recs[c] = p[0]
        want.append(o)
        meta.append(p)
    out = np.empty_like(imgs)
    harness.aug_host(imgs.ctypes.data_as(C.c_void_p), C.byref(recs), out.ctypes.data_as(C.c_void_p), n)
    seen = set()
    for c in range(n):
        o_img, o_lab, o_cube, o_com, o_M, o_rot = want[c]
        _, lab, cube, com, M, rot = meta[c]
        assert np.array_equal(out[c], o_img), (modes, c, int((out[c] != o_img).sum()))
        assert np.array_equal(np.asarray(lab, np.float32), np.asarray(o_lab, np.float32))
        assert np.array_equal(np.asarray(cube, np.float32), np.asarray(o_cube, np.float32))
        assert np.array_equal(np.asarray(com, np.float32), np.asarray(o_com, np.float32))
        assert np.array_equal(M, o_M) and float(rot) == float(o_rot)
        seen.add(int(recs[c].mode))
    if len(modes) > 1:
        assert seen == {0, 1, 2}      # none / perspective / affine all exercised
