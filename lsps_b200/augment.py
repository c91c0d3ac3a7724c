"""Batch augmentation of depth crops on the GPU.

The reference augments every sample on the host inside its DataLoader workers
(/root/reference/src/data/dataset_hand2.py:34-119 `augmentCrop`, driving utils/handdetector.py:682-808 and two cv2
nearest-neighbour warps; ~750 us per sample per core).  Here the host only draws the random numbers -- in the reference's
order, from the same `numpy.random.RandomState` -- and turns them into one small parameter record per crop
(`sample_params`); one kernel launch (`lsps_augment_crops`) then does everything that touches pixels for the whole batch.
The label / com / cube / M updates are 36x3 numbers per sample and stay on the host, written as the reference writes them
(float32 intermediates included) so that they come out identical.

Status: the per-pixel arithmetic (csrc/augment_core.h) is verified bit-for-bit against the pinned numpy oracle through a
host build (tests/test_augment_core_cpu.py); the device launch has not run on hardware yet (round-1 GPU budget was spent).
"""
import ctypes as C

import numpy as np

AUG_NONE, AUG_PERSPECTIVE, AUG_AFFINE = 0, 1, 2
NYU_CAMERA = (588.03, 587.07, 320.0, 240.0)     # fx, fy, ux, uy (data/importers.py, NYUImporter)


class AugSample(C.Structure):
    """Mirror of `lsps_aug_sample` (csrc/augment_core.h)."""
    _fields_ = [("mode", C.c_int), ("pad_", C.c_int), ("m", C.c_double * 9), ("dn_scale", C.c_float), ("dn_off", C.c_float),
                ("zstart", C.c_float), ("zend", C.c_float), ("far_", C.c_float), ("near_", C.c_float),
                ("out_off", C.c_float), ("out_scale", C.c_float)]


class Camera(object):
    """Pinhole projection of data/importers.py:84-123 (float32 result buffers as there)."""

    def __init__(self, fx, fy, ux, uy):
        self.fx, self.fy, self.ux, self.uy = fx, fy, ux, uy

    def img_to_3d(self, s):
        ret = np.zeros((3,), np.float32)
        ret[0] = (s[0] - self.ux) * s[2] / self.fx
        ret[1] = (s[1] - self.uy) * s[2] / self.fy
        ret[2] = s[2]
        return ret

    def to_img(self, s):
        ret = np.zeros((3,), np.float32)
        if s[2] == 0.:
            ret[0], ret[1] = self.ux, self.uy
            return ret
        ret[0] = s[0] / s[2] * self.fx + self.ux
        ret[1] = s[1] / s[2] * self.fy + self.uy
        ret[2] = s[2]
        return ret


def _bounds(com, size, cam):
    """handdetector.py:206-228 (com[2] != 0)."""
    xs = int(np.floor((com[0] * com[2] / cam.fx - size[0] / 2.) / com[2] * cam.fx + 0.5))
    xe = int(np.floor((com[0] * com[2] / cam.fx + size[0] / 2.) / com[2] * cam.fx + 0.5))
    ys = int(np.floor((com[1] * com[2] / cam.fy - size[1] / 2.) / com[2] * cam.fy + 0.5))
    ye = int(np.floor((com[1] * com[2] / cam.fy + size[1] / 2.) / com[2] * cam.fy + 0.5))
    return xs, xe, ys, ye, com[2] - size[2] / 2., com[2] + size[2] / 2.


def com_to_transform(com, size, cam, dsize=(128, 128)):
    """handdetector.py:230-260: crop rectangle of the cube around com -> dsize, aspect preserved, centred."""
    xs, xe, ys, ye, _, _ = _bounds(com, size, cam)
    trans = np.eye(3)
    trans[0, 2], trans[1, 2] = -xs, -ys
    wb, hb = xe - xs, ye - ys
    if wb > hb:
        scale = np.eye(3) * dsize[0] / float(wb)
        sz = (dsize[0], hb * dsize[0] / wb)
    else:
        scale = np.eye(3) * dsize[1] / float(hb)
        sz = (wb * dsize[1] / hb, dsize[1])
    scale[2, 2] = 1
    off = np.eye(3)
    off[0, 2] = int(np.floor(dsize[0] / 2. - sz[1] / 2.))
    off[1, 2] = int(np.floor(dsize[1] / 2. - sz[0] / 2.))
    return np.dot(off, np.dot(scale, trans))


def _invert3(S):
    """cv::invert of a 3x3 double matrix: closed form through the cofactors (what cv2.warpPerspective applies to M)."""
    d = (S[0, 0] * (S[1, 1] * S[2, 2] - S[1, 2] * S[2, 1]) - S[0, 1] * (S[1, 0] * S[2, 2] - S[1, 2] * S[2, 0]) +
         S[0, 2] * (S[1, 0] * S[2, 1] - S[1, 1] * S[2, 0]))
    if d == 0.:
        return np.zeros((3, 3))
    d = 1. / d
    t = np.empty((3, 3))
    t[0, 0] = (S[1, 1] * S[2, 2] - S[1, 2] * S[2, 1]) * d
    t[0, 1] = (S[0, 2] * S[2, 1] - S[0, 1] * S[2, 2]) * d
    t[0, 2] = (S[0, 1] * S[1, 2] - S[0, 2] * S[1, 1]) * d
    t[1, 0] = (S[1, 2] * S[2, 0] - S[1, 0] * S[2, 2]) * d
    t[1, 1] = (S[0, 0] * S[2, 2] - S[0, 2] * S[2, 0]) * d
    t[1, 2] = (S[0, 2] * S[1, 0] - S[0, 0] * S[1, 2]) * d
    t[2, 0] = (S[1, 0] * S[2, 1] - S[1, 1] * S[2, 0]) * d
    t[2, 1] = (S[0, 1] * S[2, 0] - S[0, 0] * S[2, 1]) * d
    t[2, 2] = (S[0, 0] * S[1, 1] - S[0, 1] * S[1, 0]) * d
    return t


def _rotation_inverse(center, angle_deg):
    """cv2.getRotationMatrix2D(center, angle, 1) followed by cv::invertAffineTransform (what cv2.warpAffine applies)."""
    a = angle_deg * (np.pi / 180.0)
    alpha, beta = np.cos(a), np.sin(a)
    M = np.array([[alpha, beta, (1 - alpha) * center[0] - beta * center[1]],
                  [-beta, alpha, beta * center[0] + (1 - alpha) * center[1]]], np.float64)
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[1, 1] * D, M[0, 0] * D
    M[0, 0] = A11
    M[0, 1] *= -D
    M[1, 0] *= -D
    M[1, 1] = A22
    b1 = -M[0, 0] * M[0, 2] - M[0, 1] * M[1, 2]
    b2 = -M[1, 0] * M[0, 2] - M[1, 1] * M[1, 2]
    M[0, 2], M[1, 2] = b1, b2
    return M


def _rotate_point_2d(p1, center, angle):
    """data/transformations.py:71-88."""
    alpha = angle * np.pi / 180.
    pp = p1.copy()
    pp[0:2] -= center[0:2]
    pr = np.zeros_like(pp)
    pr[0] = pp[0] * np.cos(alpha) - pp[1] * np.sin(alpha)
    pr[1] = pp[0] * np.sin(alpha) + pp[1] * np.cos(alpha)
    pr[2] = pp[2]
    pr[0:2] += center[0:2]
    return pr


def sample_params(gt3dcrop, com, cube, M, aug_modes, cam, rng, sigma_com=10., sigma_sc=0.05, rot_range=180.):
    """Everything of augmentCrop that does not touch pixels, for ONE sample.  Consumes rng exactly like the reference
    (randint, randn(3), uniform, randn).  Returns (AugSample, label, cube, com, M, rot) -- the last five as the
    reference returns them (dataset_hand2.py:119)."""
    com = np.asarray(com, np.float32)
    cube = np.asarray(cube, np.float32)
    p = AugSample()
    p.dn_scale, p.dn_off = cube[2] / 2., com[2]
    mode = rng.randint(0, len(aug_modes))
    off = rng.randn(3) * sigma_com
    rot = rng.uniform(-rot_range, rot_range)
    sc = abs(1. + rng.randn() * sigma_sc)
    name = aug_modes[mode]
    p.mode = AUG_NONE
    new_cube, new_com, new_M, joints = cube, com, M, gt3dcrop
    if name == 'com':
        rot, sc = 0., 1.
        if not np.allclose(off, 0.):
            new_com = cam.to_img(cam.img_to_3d(com) + off)
            if not (np.allclose(com[2], 0.) or np.allclose(new_com[2], 0.)):
                new_M = com_to_transform(new_com, cube, cam, (128, 128))
                T = np.dot(new_M, np.linalg.inv(M))
                p.mode = AUG_PERSPECTIVE
                p.m[:] = _invert3(np.asarray(T, np.float64)).reshape(-1).tolist()
                p.zstart, p.zend = new_com[2] - cube[2] / 2., new_com[2] + cube[2] / 2.
            joints = gt3dcrop + cam.img_to_3d(com) - cam.img_to_3d(new_com)
    elif name == 'rot':
        sc = 1.
        if not np.allclose(rot, 0.):
            rot = np.mod(rot, 360)
            p.mode = AUG_AFFINE
            p.m[:6] = _rotation_inverse((128 // 2, 128 // 2), -rot).reshape(-1).tolist()
            com3d = cam.img_to_3d(com)
            j2d = np.stack([cam.to_img(j) for j in (gt3dcrop + com3d)]).astype(np.float32)
            r2d = np.zeros_like(j2d)
            for k in range(j2d.shape[0]):
                r2d[k] = _rotate_point_2d(j2d[k], com[0:2], rot)
            joints = np.stack([cam.img_to_3d(j) for j in r2d]).astype(np.float32) - com3d
    elif name == 'sc':
        rot = 0.
        if not np.allclose(sc, 1.):
            new_cube = [s * sc for s in cube]
            if not np.allclose(com[2], 0.):
                new_M = com_to_transform(com, new_cube, cam, (128, 128))
                T = np.dot(new_M, np.linalg.inv(M))
                p.mode = AUG_PERSPECTIVE
                p.m[:] = _invert3(np.asarray(T, np.float64)).reshape(-1).tolist()
                p.zstart, p.zend = com[2] - cube[2] / 2., com[2] + cube[2] / 2.
    elif name == 'none':
        rot = 0.
    else:
        raise NotImplementedError(name)
    label = joints / (new_cube[2] / 2.)           # the cube as reassigned by the move (dataset_hand2.py:96-98: the NEW cube in sc mode)
    p.far_ = new_com[2] + (new_cube[2] / 2.)
    p.near_ = new_com[2] - (new_cube[2] / 2.)
    p.out_off = new_com[2]
    p.out_scale = new_cube[2] / 2.
    return p, label, np.asarray(new_cube), new_com, np.array(new_M, dtype='float32'), rot


class CropAugmenter(object):
    """Batch front-end: `imgs` is a float32 CUDA tensor (n,128,128) or (n,1,128,128) of NORMALISED crops (normalize(),
    dataset_hand2.py:27-31); labels / coms / cubes / Ms are per-sample host arrays.  Returns the augmented crops (new CUDA
    tensor) and the per-sample host results in the reference's item layout."""

    def __init__(self, camera=NYU_CAMERA, aug_modes=("com", "rot", "sc", "none"), seed=23455, device=None):
        self.cam = Camera(*camera)
        self.aug_modes = list(aug_modes)
        self.rng = np.random.RandomState(seed)
        self.device = device

    def __call__(self, imgs, gt3dcrops, coms, cubes, Ms):
        import torch
        from . import _lib
        n = imgs.shape[0]
        x = imgs.reshape(n, 128, 128).contiguous().float()
        recs = (AugSample * n)()
        labels, ncubes, ncoms, nMs, rots = [], [], [], [], []
        for i in range(n):
            p, lab, cube, com, M, rot = sample_params(np.asarray(gt3dcrops[i], np.float32), coms[i], cubes[i],
                                                      np.asarray(Ms[i], np.float32), self.aug_modes, self.cam, self.rng)
            recs[i] = p
            labels.append(lab); ncubes.append(cube); ncoms.append(com); nMs.append(M); rots.append(rot)
        raw = torch.frombuffer(bytearray(bytes(recs)), dtype=torch.uint8).to(x.device)
        premax = torch.empty(n, dtype=torch.float32, device=x.device)
        out = torch.empty_like(x)
        ctx = _lib.context(x.device.index)
        ctx.augment_crops(x.data_ptr(), raw.data_ptr(), premax.data_ptr(), out.data_ptr(), n)
        return out.reshape(imgs.shape), labels, ncubes, ncoms, nMs, rots
