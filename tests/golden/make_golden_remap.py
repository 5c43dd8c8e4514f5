"""Golden vectors for the undistortion remap: cv2.initUndistortRectifyMap (CV_16SC2) + cv2.remap(INTER_LINEAR) with a
quarter-scale EuRoC cam0 calibration (euroc_E.cpp:88-113), and a random fixed-point map that leaves the image on all
sides.  Run here with cv2 4.13.0; outputs travel as tests/golden/remap.npz.   python tests/golden/make_golden_remap.py"""
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from epivo_b200 import datasets as D


def main():
    rng = np.random.default_rng(88)
    s = 0.25
    K = D.EUROC_CAM0_K.copy(); K[:2] *= s
    P = D.EUROC_CAM0_PROJ.copy(); P[:2] *= s
    w, h = 188, 120
    m1, m2 = cv2.initUndistortRectifyMap(K, D.EUROC_CAM0_DIST, D.EUROC_CAM0_RECT, P, (w, h), cv2.CV_16SC2)
    img = cv2.GaussianBlur(rng.integers(0, 256, (h, w)).astype(np.uint8), (0, 0), 1.0)
    res = {"K": K, "dist": D.EUROC_CAM0_DIST, "R": D.EUROC_CAM0_RECT, "P": P, "size": np.array([w, h]),
           "map1": m1, "map2": m2, "img": img, "out": cv2.remap(img, m1, m2, cv2.INTER_LINEAR)}
    r1 = rng.integers(-30, 220, (64, 80, 2)).astype(np.int16)
    r2 = rng.integers(0, 1024, (64, 80)).astype(np.uint16)
    res.update(rmap1=r1, rmap2=r2, rout0=cv2.remap(img, r1, r2, cv2.INTER_LINEAR),
               rout77=cv2.remap(img, r1, r2, cv2.INTER_LINEAR, borderValue=77), cv2_version=np.array(cv2.__version__))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "remap.npz"), **res)
    print("ok", m1.shape, m2.shape)


if __name__ == "__main__":
    main()
