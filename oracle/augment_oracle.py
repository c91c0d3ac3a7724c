"""TEST INFRASTRUCTURE ONLY -- CPU oracle ("port") of the reference's per-sample input pipeline:
`normalize` + `augmentCrop` of a 128x128 depth crop and its joint labels, i.e. what 3x4 DataLoader worker processes do
in the reference before every training step.  numpy only: the two OpenCV warps the reference calls
(cv2.warpPerspective / cv2.warpAffine, nearest neighbour, constant border) are restated here from OpenCV's published
algorithm so that a future CUDA kernel has an exact definition to match; `oracle/make_augment_golden.py` pins this file
against the UNMODIFIED reference functions (which call the real cv2) on seeded inputs and writes tests/golden/augment_*.npz.

Status: groundwork for round 2 -- there is NO product kernel for this row yet and nothing in lsps_b200/ imports this file.

Reference anchors (paths relative to /root/reference/src):
  data/dataset_hand2.py:27-31    normalize
  data/dataset_hand2.py:34-119   augmentCrop (modes 'com' | 'rot' | 'sc' | 'none'; normZeroOne=False as the datasets call it)
  utils/handdetector.py:206-260  comToBounds, comToTransform
  utils/handdetector.py:682-711  moveCoM      :713-751 rotateHand      :754-783 scaleHand      :785-808 recropHand
  data/importers.py:84-123       jointImgTo3D / joint3DToImg (pinhole camera)
  data/transformations.py:71-88  rotatePoint2D
Third-party arithmetic restated (OpenCV 4.x, modules/imgproc/src/imgwarp.cpp): warpPerspective inverts the 3x3 matrix in
double, evaluates (X, Y) = ((M0 x + M1 y + M2) / W, (M3 x + M4 y + M5) / W) in double per destination pixel (64-pixel
wide blocks: the x term is added to the block's base) and rounds half-to-even; warpAffine inverts the 2x3 matrix in
double and walks 10-bit fixed point (AB_BITS) with round-to-nearest tables for the x and y terms.
"""
import numpy as np

NYU_CAMERA = (588.03, 587.07, 320.0, 240.0)     # fx, fy, ux, uy (data/importers.py NYUImporter)
AB_BITS = 10


# ----------------------------------------------------------------------------------------------- camera
class Camera(object):
    """data/importers.py:55-123 -- the float32 `ret` buffers of the reference are kept (they round the results)."""

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

    def imgs_to_3d(self, pts):
        return np.stack([self.img_to_3d(p) for p in pts]).astype(np.float32)

    def to_imgs(self, pts):
        return np.stack([self.to_img(p) for p in pts]).astype(np.float32)


# ----------------------------------------------------------------------------------------------- OpenCV warps, NN
def _round_half_even(v):
    """cvRound / saturate_cast<int>(double): nearest integer, ties to even, saturated to int32."""
    return np.clip(np.rint(v), -2147483648.0, 2147483647.0).astype(np.int64)


def _gather_nn(src, X, Y, border):
    h, w = src.shape
    inside = (X >= 0) & (X < w) & (Y >= 0) & (Y < h)
    out = np.full(X.shape, border, dtype=src.dtype)
    out[inside] = src[Y[inside], X[inside]]
    return out


def _invert3(S):
    """cv::invert of a 3x3 CV_64F matrix (matrix_decomp / lapack.cpp: closed form through the cofactors, not LU)."""
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


TIE_EPS = 1e-7      # pixels: source coordinates this close to a decision boundary are "ties" (see warp_perspective_nn)


def warp_perspective_nn(src, M, dsize, border=0.0, return_ties=False):
    """cv2.warpPerspective(src, M, dsize, flags=INTER_NEAREST, borderMode=BORDER_CONSTANT, borderValue=border) as
    OpenCV 4.13 executes it (determined against the library, see make_augment_golden.py): the matrix is inverted with
    cv::invert's closed form, source coordinates are evaluated in double per destination pixel, a pixel whose
    CONTINUOUS source coordinate lies outside [0, w-1] x [0, h-1] takes the border value (older releases rounded
    first), and the nearest pixel is floor(c + 0.5).
    Exactness: bit-identical to the library on random matrices (9.8 M pixels).  Crop-to-crop transforms are ratios of
    small integers, so whole rows/columns of source coordinates can sit EXACTLY on k + 0.5 or on the image edge; there
    the last bit of the library's (unpublished, SIMD) evaluation order decides and no restatement of a*x + b*y + c in
    plain or fused double arithmetic reproduces every decision.  This function is the definition the CUDA kernel has to
    match; `return_ties` also returns the mask of pixels within TIE_EPS of such a boundary, which is where (and only
    where) it may differ from cv2."""
    wd, hd = int(dsize[0]), int(dsize[1])
    Mi = _invert3(np.asarray(M, np.float64))
    x = np.arange(wd, dtype=np.float64)[None, :]
    y = np.arange(hd, dtype=np.float64)[:, None]
    W = Mi[2, 0] * x + Mi[2, 1] * y + Mi[2, 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        fX = (Mi[0, 0] * x + Mi[0, 1] * y + Mi[0, 2]) / W
        fY = (Mi[1, 0] * x + Mi[1, 1] * y + Mi[1, 2]) / W
    h, w = src.shape
    inside = (fX >= 0) & (fX <= w - 1) & (fY >= 0) & (fY <= h - 1)
    X = np.floor(np.where(inside, fX, 0.0) + 0.5).astype(np.int64)
    Y = np.floor(np.where(inside, fY, 0.0) + 0.5).astype(np.int64)
    out = np.full((hd, wd), border, dtype=src.dtype)
    out[inside] = src[Y[inside], X[inside]]
    if not return_ties:
        return out
    with np.errstate(invalid="ignore"):
        def near(c, n):
            return (np.abs(c - np.floor(c) - 0.5) < TIE_EPS) | (np.abs(c) < TIE_EPS) | (np.abs(c - (n - 1)) < TIE_EPS)
        # a tie only matters if the pixel is (nearly) inside along the other axis as well
        ties = (near(fX, w) & (fY > -0.5) & (fY < h - 0.5)) | (near(fY, h) & (fX > -0.5) & (fX < w - 0.5))
    return out, ties


def rotation_matrix_2d(center, angle_deg, scale):
    """cv2.getRotationMatrix2D."""
    a = angle_deg * (np.pi / 180.0)                        # OpenCV: angle *= CV_PI/180
    alpha, beta = np.cos(a) * scale, np.sin(a) * scale
    return np.array([[alpha, beta, (1 - alpha) * center[0] - beta * center[1]],
                     [-beta, alpha, beta * center[0] + (1 - alpha) * center[1]]], np.float64)


def warp_affine_nn(src, M, dsize, border=0.0):
    """cv2.warpAffine(src, M, dsize, flags=INTER_NEAREST, borderMode=BORDER_CONSTANT, borderValue=border)."""
    wd, hd = int(dsize[0]), int(dsize[1])
    M = np.asarray(M, np.float64).copy()
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]              # invertAffineTransform, in place like OpenCV
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[1, 1] * D, M[0, 0] * D
    M[0, 0] = A11
    M[0, 1] *= -D
    M[1, 0] *= -D
    M[1, 1] = A22
    b1 = -M[0, 0] * M[0, 2] - M[0, 1] * M[1, 2]
    b2 = -M[1, 0] * M[0, 2] - M[1, 1] * M[1, 2]
    M[0, 2], M[1, 2] = b1, b2
    scale = float(1 << AB_BITS)
    x = np.arange(wd, dtype=np.float64)
    y = np.arange(hd, dtype=np.float64)
    adelta = _round_half_even(M[0, 0] * x * scale)
    bdelta = _round_half_even(M[1, 0] * x * scale)
    rd = (1 << AB_BITS) // 2
    X0 = _round_half_even((M[0, 1] * y + M[0, 2]) * scale) + rd
    Y0 = _round_half_even((M[1, 1] * y + M[1, 2]) * scale) + rd
    X = (X0[:, None] + adelta[None, :]) >> AB_BITS
    Y = (Y0[:, None] + bdelta[None, :]) >> AB_BITS
    return _gather_nn(src, X, Y, border)


# ----------------------------------------------------------------------------------------------- crop geometry
def com_to_bounds(com, size, cam):
    """utils/handdetector.py:206-228 (the com[2] == 0 branch needs the full depth map and is not part of this path)."""
    zstart = com[2] - size[2] / 2.
    zend = com[2] + size[2] / 2.
    xstart = int(np.floor((com[0] * com[2] / cam.fx - size[0] / 2.) / com[2] * cam.fx + 0.5))
    xend = int(np.floor((com[0] * com[2] / cam.fx + size[0] / 2.) / com[2] * cam.fx + 0.5))
    ystart = int(np.floor((com[1] * com[2] / cam.fy - size[1] / 2.) / com[2] * cam.fy + 0.5))
    yend = int(np.floor((com[1] * com[2] / cam.fy + size[1] / 2.) / com[2] * cam.fy + 0.5))
    return xstart, xend, ystart, yend, zstart, zend


def com_to_transform(com, size, cam, dsize=(128, 128)):
    """utils/handdetector.py:230-260."""
    xstart, xend, ystart, yend, _, _ = com_to_bounds(com, size, cam)
    trans = np.eye(3)
    trans[0, 2] = -xstart
    trans[1, 2] = -ystart
    wb, hb = (xend - xstart), (yend - ystart)
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


def recrop(crop, M, Mnew, target_size, cam, background=0., nv_val=0., com=None, size=None, ties=None):
    """utils/handdetector.py:785-808 with thresh_z=True (the only way augmentCrop calls it).  `ties`: optional list that
    receives the warp's tie mask."""
    warped, tm = warp_perspective_nn(crop, np.dot(M, Mnew), target_size, float(background), return_ties=True)
    if ties is not None:
        ties.append(tm)
    warped[np.isclose(warped, nv_val)] = background
    _, _, _, _, zstart, zend = com_to_bounds(com, size, cam)
    msk1 = np.logical_and(warped < zstart, warped != 0)
    msk2 = np.logical_and(warped > zend, warped != 0)
    warped[msk1] = zstart
    warped[msk2] = 0.
    return warped


def rotate_point_2d(p1, center, angle):
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


# ----------------------------------------------------------------------------------------------- the three moves
def move_com(dpt, cube, com, off, joints3d, M, cam, ties=None):
    """utils/handdetector.py:682-711."""
    if np.allclose(off, 0.):
        return dpt, joints3d, com, M
    new_com = cam.to_img(cam.img_to_3d(com) + off)
    if not (np.allclose(com[2], 0.) or np.allclose(new_com[2], 0.)):
        Mnew = com_to_transform(new_com, cube, cam, dpt.shape)
        new_dpt = recrop(dpt, Mnew, np.linalg.inv(M), dpt.shape, cam, background=0, nv_val=32000., com=new_com, size=cube,
                         ties=ties)
    else:
        Mnew, new_dpt = M, dpt
    new_joints3d = joints3d + cam.img_to_3d(com) - cam.img_to_3d(new_com)
    return new_dpt, new_joints3d, new_com, Mnew


def rotate_hand(dpt, cube, com, rot, joints3d, cam):
    """utils/handdetector.py:713-751."""
    if np.allclose(rot, 0.):
        return dpt, joints3d, rot
    rot = np.mod(rot, 360)
    M = rotation_matrix_2d((dpt.shape[1] // 2, dpt.shape[0] // 2), -rot, 1)
    new_dpt = warp_affine_nn(dpt, M, (dpt.shape[1], dpt.shape[0]), 0)
    com3d = cam.img_to_3d(com)
    joint_2d = cam.to_imgs(joints3d + com3d)
    data_2d = np.zeros_like(joint_2d)
    for k in range(data_2d.shape[0]):
        data_2d[k] = rotate_point_2d(joint_2d[k], com[0:2], rot)
    return new_dpt, cam.imgs_to_3d(data_2d) - com3d, rot


def scale_hand(dpt, cube, com, sc, joints3d, M, cam, ties=None):
    """utils/handdetector.py:754-783."""
    if np.allclose(sc, 1.):
        return dpt, joints3d, cube, M
    new_cube = [s * sc for s in cube]
    if not np.allclose(com[2], 0.):
        Mnew = com_to_transform(com, new_cube, cam, dpt.shape)
        new_dpt = recrop(dpt, Mnew, np.linalg.inv(M), dpt.shape, cam, background=0, nv_val=32000., com=com, size=cube,
                         ties=ties)
    else:
        Mnew, new_dpt = M, dpt
    return new_dpt, joints3d, new_cube, Mnew


# ----------------------------------------------------------------------------------------------- entry points
def normalize(img, com, cube):
    """data/dataset_hand2.py:27-31 (in place on a float32 crop in mm; background 0 -> far plane -> +1)."""
    img[img == 0] = com[2] + (cube[2] / 2.)
    img -= com[2]
    img /= (cube[2] / 2.)
    return img


def augment_crop(img, gt3dcrop, com, cube, M, aug_modes, cam, rng, sigma_com=10., sigma_sc=0.05, rot_range=180., ties=None):
    """data/dataset_hand2.py:34-119 with normZeroOne=False.  Consumes the SAME numpy RandomState draws in the same order:
    randint(len(modes)), randn(3), uniform(-rot, rot), randn().  `ties`: optional list receiving the tie mask of the
    perspective warp ('com' / 'sc' modes), see warp_perspective_nn."""
    img = img * (cube[2] / 2.) + com[2]
    premax = img.max()
    mode = rng.randint(0, len(aug_modes))
    off = rng.randn(3) * sigma_com
    rot = rng.uniform(-rot_range, rot_range)
    sc = abs(1. + rng.randn() * sigma_sc)
    name = aug_modes[mode]
    if name == 'com':
        rot, sc = 0., 1.
        imgD, new_j, com, M = move_com(img.astype('float32'), cube, com, off, gt3dcrop, M, cam, ties)
        label = new_j / (cube[2] / 2.)
    elif name == 'rot':
        imgD, new_j, rot = rotate_hand(img.astype('float32'), cube, com, rot, gt3dcrop, cam)
        label = new_j / (cube[2] / 2.)
    elif name == 'sc':
        rot = 0.
        imgD, new_j, cube, M = scale_hand(img.astype('float32'), cube, com, sc, gt3dcrop, M, cam, ties)
        label = new_j / (cube[2] / 2.)
    elif name == 'none':
        rot = 0.
        imgD = img
        label = gt3dcrop / (cube[2] / 2.)
    else:
        raise NotImplementedError(name)
    far = com[2] + (cube[2] / 2.)
    near = com[2] - (cube[2] / 2.)
    imgD[imgD == premax] = far
    imgD[imgD == 0] = far
    imgD[imgD >= far] = far
    imgD[imgD <= near] = near
    imgD -= com[2]
    imgD /= (cube[2] / 2.)
    return imgD, label, np.asarray(cube), com, np.array(M, dtype='float32'), rot


def synthetic_crop(rng, cam=None):
    """A seeded 128x128 hand-like depth crop in mm (background 0) with its com (image coordinates), cube, crop
    transform M and 36 joints relative to the com -- the item layout of dataset_hand2.py:329-366."""
    cam = cam or Camera(*NYU_CAMERA)
    yy, xx = np.mgrid[0:128, 0:128]
    cx, cy = 64 + rng.uniform(-8, 8), 62 + rng.uniform(-8, 8)
    ax, ay = rng.uniform(28, 48), rng.uniform(34, 54)
    blob = ((xx - cx) ** 2 / ax ** 2 + (yy - cy) ** 2 / ay ** 2) < 1
    z0 = rng.uniform(450, 900)
    dpt = np.zeros((128, 128), np.float32)
    dpt[blob] = (z0 + 40 * np.sin(xx[blob] / 9.) + 25 * np.cos(yy[blob] / 7.) + rng.randn(int(blob.sum())) * 4).astype(np.float32)
    com = np.array([rng.uniform(200, 440), rng.uniform(150, 330), z0], np.float32)
    cube = np.array([300., 300., 300.], np.float32)
    M = com_to_transform(com, cube, cam, (128, 128)).astype('float32')
    gt = (rng.randn(36, 3) * 40).astype('float32')
    return dpt, com, cube, M, gt
