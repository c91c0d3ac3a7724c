"""TEST INFRASTRUCTURE ONLY -- pins `oracle/augment_oracle.py` against the unmodified reference input pipeline
(`data/dataset_hand2.py` normalize + augmentCrop driving `utils/handdetector.py` and the real cv2) and writes the
committed fixtures tests/golden/augment_*.npz.   Run where /root/reference is mounted:  python oracle/make_augment_golden.py

Third-party dependency under the reference: OpenCV (cv2.warpPerspective / warpAffine / getRotationMatrix2D).  The
reference pins no version; this container has cv2 4.13.0 and the restated warps follow what THAT library computes
(nearest neighbour; its warpPerspective tests the continuous source coordinate against the image rectangle before
rounding) -- bit-exact on 600 random matrices each before any fixture is written."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import augment_oracle as A       # noqa: E402
import ref_loader                # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")
MODES = (["com"], ["rot"], ["sc"], ["none"], ["com", "rot", "sc", "none"])


def check_warps(n=600):
    import cv2
    rng = np.random.RandomState(0)
    for it in range(n):
        src = (rng.rand(128, 128) * 1000).astype(np.float32)
        s = rng.uniform(0.6, 1.6)
        tx, ty = rng.uniform(-40, 40, 2)
        M = np.array([[s, rng.uniform(-0.05, 0.05), tx], [rng.uniform(-0.05, 0.05), s * rng.uniform(0.9, 1.1), ty],
                      [rng.uniform(-1e-4, 1e-4), rng.uniform(-1e-4, 1e-4), 1.0]])
        if it % 2 == 0:
            M[2, :2] = 0
        ref = cv2.warpPerspective(src, M, (128, 128), flags=cv2.INTER_NEAREST, borderMode=cv2.BORDER_CONSTANT, borderValue=0.0)
        assert np.array_equal(ref, A.warp_perspective_nn(src, M, (128, 128), 0.0)), ("warpPerspective", it)
        ang = rng.uniform(-180, 180)
        R = cv2.getRotationMatrix2D((64, 64), ang, 1)
        assert np.array_equal(R, A.rotation_matrix_2d((64, 64), ang, 1)), ("getRotationMatrix2D", it)
        ref = cv2.warpAffine(src, R, (128, 128), flags=cv2.INTER_NEAREST, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
        assert np.array_equal(ref, A.warp_affine_nn(src, R, (128, 128), 0)), ("warpAffine", it)
    return cv2.__version__


def main():
    ver = check_warps()
    print("restated NN warps == cv2 %s on 600 random matrices each" % ver)
    normalize, augmentCrop, HandDetector, DepthImporter = ref_loader.load_reference_augment()
    di = DepthImporter(*A.NYU_CAMERA)
    cam = A.Camera(*A.NYU_CAMERA)
    os.makedirs(OUT, exist_ok=True)
    for mi, modes in enumerate(MODES):
        rec, worst_lab, px, ntie, ntie_diff = {}, 0.0, 0, 0, 0
        ncase = 24 if len(modes) == 1 else 64
        for c in range(ncase):
            rs = np.random.RandomState(1000 * mi + c)
            dpt, com, cube, M, gt = A.synthetic_crop(rs, cam)
            img_r = normalize(dpt.copy(), com, cube)
            img_o = A.normalize(dpt.copy(), com, cube)
            assert np.array_equal(img_r, img_o)
            hd = HandDetector(img_r.copy(), abs(di.fx), abs(di.fy), importer=di)
            assert np.allclose(hd.comToTransform(com, cube, (128, 128)), A.com_to_transform(com, cube, cam, (128, 128)), atol=0)
            r_img, _, r_lab, r_cube, r_com, r_M, r_rot = augmentCrop(img_r.copy(), gt.copy(), com.copy(), cube.copy(), M.copy(),
                                                                     list(modes), hd, rng=np.random.RandomState(7 + c))
            ties = []
            o_img, o_lab, o_cube, o_com, o_M, o_rot = A.augment_crop(img_o.copy(), gt.copy(), com.copy(), cube.copy(), M.copy(),
                                                                     list(modes), cam, np.random.RandomState(7 + c), ties=ties)
            tie = ties[0] if ties else np.zeros((128, 128), bool)
            if not np.array_equal(r_img[~tie], o_img[~tie]):
                raise SystemExit("image differs off the tie mask: modes %s case %d, %d pixels"
                                 % (modes, c, int((r_img != o_img)[~tie].sum())))
            ntie += int(tie.sum())
            ntie_diff += int((r_img != o_img).sum())
            for a, b, nm in ((r_lab, o_lab, "label"), (r_cube, o_cube, "cube"), (r_com, o_com, "com"), (r_M, o_M, "M")):
                if not np.allclose(np.asarray(a, np.float64), np.asarray(b, np.float64), rtol=1e-6, atol=1e-6):
                    raise SystemExit("%s differs: modes %s case %d" % (nm, modes, c))
            assert float(r_rot) == float(o_rot)
            worst_lab = max(worst_lab, float(np.max(np.abs(np.asarray(r_lab, np.float64) - np.asarray(o_lab, np.float64)))))
            px += r_img.size
            if c < 6:     # commit a few complete cases per mode list (inputs are regenerated from the seed)
                rec["c%d_img" % c] = r_img.astype(np.float32)
                rec["c%d_tie" % c] = np.packbits(tie)
                rec["c%d_label" % c] = np.asarray(r_lab, np.float32)
                rec["c%d_cube" % c] = np.asarray(r_cube, np.float32)
                rec["c%d_com" % c] = np.asarray(r_com, np.float32)
                rec["c%d_M" % c] = np.asarray(r_M, np.float32)
                rec["c%d_rot" % c] = np.float64(r_rot)
        name = "augment_" + "_".join(modes)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), meta_modes=np.array(",".join(modes)), meta_seed0=np.array(1000 * mi),
                            meta_rng0=np.array(7), meta_cv2=np.array(ver), **rec)
        print("%-28s ok  (%d cases, %d pixels; %d on the tie mask, %d of those differ from cv2; labels max abs diff %.2e)"
              % (name, ncase, px, ntie, ntie_diff, worst_lab))


if __name__ == "__main__":
    main()
