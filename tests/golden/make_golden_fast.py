"""Golden vectors for the FAST-9/16 detector: cv2.FastFeatureDetector on synthetic 8-bit images (run here, where cv2
4.13.0 is installed; the outputs travel as tests/golden/fast.npz).   python tests/golden/make_golden_fast.py"""
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def images():
    """Small images that exercise the detector: smoothed noise (natural-looking corners), raw noise (dense corners,
    many equal scores: the strict-> suppression rule), a checkerboard with saturated 0/255 cells, thin structures at
    the 3-pixel border, a constant image and a wide smoothed-noise frame (the KITTI-sized one is compared live, tests/test_gpu_fast.py)."""
    rng = np.random.default_rng(20260101)
    out = {}
    n = rng.integers(0, 256, (96, 128)).astype(np.uint8)
    out["noise"] = n
    out["smooth"] = cv2.GaussianBlur(rng.integers(0, 256, (120, 160)).astype(np.uint8), (0, 0), 2.0)
    cb = (((np.arange(90)[:, None] // 9) + (np.arange(117)[None, :] // 9)) % 2 * 255).astype(np.uint8)
    out["checker"] = cb
    b = np.full((40, 50), 100, np.uint8)
    b[3, :] = 200; b[:, 3] = 10; b[36, 5:45] = 255; b[10:30, 46] = 0; b[20, 20] = 255; b[21, 21] = 0
    out["border"] = b
    out["const"] = np.full((16, 16), 77, np.uint8)
    out["tiny"] = rng.integers(0, 256, (7, 9)).astype(np.uint8)
    lvl = rng.integers(0, 4, (64, 64)).astype(np.uint8) * 85            # few grey levels: many score ties
    out["levels"] = lvl
    k = cv2.GaussianBlur(rng.integers(0, 256, (141, 467)).astype(np.uint8), (0, 0), 1.5)   # odd sizes, wider than a CTA row
    out["wide"] = cv2.normalize(k, None, 0, 255, cv2.NORM_MINMAX)
    return out


def main():
    res = {}
    for name, im in images().items():
        res["img_" + name] = im
        for thr in (10, 40):
            for nms in (True, False):
                det = cv2.FastFeatureDetector_create(thr, nms)
                kps = det.detect(im, None)
                pts = np.array([k.pt for k in kps], dtype=np.float32).reshape(-1, 2)
                resp = np.array([k.response for k in kps], dtype=np.float32)
                res[f"pts_{name}_{thr}_{int(nms)}"] = pts
                res[f"resp_{name}_{thr}_{int(nms)}"] = resp
                print(name, thr, nms, len(kps))
    res["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "fast.npz"), **res)


if __name__ == "__main__":
    main()
