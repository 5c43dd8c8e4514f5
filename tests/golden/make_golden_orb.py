"""cv2 golden vectors for ORB as the reference configures it (kitti_ba.cpp:128: ORB::create(10000, 1.2f, 8, 15, 0, 2,
FAST_SCORE), detect then compute).  Run where cv2 is installed:  python tests/golden/make_golden_orb.py
Cases cover the plain configuration, feature budgets that make retainBest cut (with ties) at every level, other pyramid
shapes, and images too small for the upper levels to hold a keypoint."""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))                       # tests/ (orb_util)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))      # the repository root (epivo_b200.synth)
from orb_util import cv2_orb, scene  # noqa: E402

CASES = {   # name: (rows, cols, seed, nfeatures, scale, nlevels, edge, fast_thr)
    "kba_small": (160, 240, 1, 10000, 1.2, 8, 15, 20),
    "budget300": (200, 320, 2, 300, 1.2, 8, 15, 20),
    "budget40": (120, 160, 3, 40, 1.2, 8, 15, 20),
    "levels4_s15": (150, 210, 4, 1500, 1.5, 4, 19, 12),
    "tiny": (61, 64, 5, 10000, 1.2, 8, 15, 20),
    "wide_budget2000": (188, 620, 6, 2000, 1.2, 8, 15, 20),
    "one_level": (100, 140, 7, 500, 1.2, 1, 31, 30),
}


def main():
    out = {"cv2_version": np.array(cv2.__version__)}
    for name, (rows, cols, seed, nf, sc, nl, edge, thr) in CASES.items():
        img = scene(rows, cols, seed)
        kps, desc = cv2_orb(cv2, img, nf, sc, nl, edge, thr)
        out["img_" + name] = img
        out["cfg_" + name] = np.array([nf, sc, nl, edge, thr], dtype=np.float64)
        out["kps_" + name] = kps
        out["desc_" + name] = desc
        print(name, img.shape, len(kps), np.bincount(kps[:, 5].astype(int), minlength=nl).tolist())
    np.savez_compressed(os.path.join(HERE, "orb.npz"), **out)


if __name__ == "__main__":
    main()
