"""Undistortion (euroc_E.cpp:105-113,169-174): the fixed-point remap restatement and the host-side map builder
against cv2 -- committed golden vectors and, where cv2 is importable, live at the EuRoC frame size."""
import os

import numpy as np
import pytest

from epivo_b200 import datasets as D
from oracle import frontend as OF

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "remap.npz"))


def test_oracle_remap_matches_cv2_golden():
    assert np.array_equal(OF.remap_bilinear_fixed(GOLD["img"], GOLD["map1"], GOLD["map2"]), GOLD["out"])
    assert np.array_equal(OF.remap_bilinear_fixed(GOLD["img"], GOLD["rmap1"], GOLD["rmap2"]), GOLD["rout0"])
    assert np.array_equal(OF.remap_bilinear_fixed(GOLD["img"], GOLD["rmap1"], GOLD["rmap2"], 77), GOLD["rout77"])


def test_map_builders_match_cv2_golden():
    size = tuple(int(v) for v in GOLD["size"])
    for f in (D.undistort_rectify_maps, OF.init_undistort_rectify_map):
        xy, fr = f(GOLD["K"], GOLD["dist"], GOLD["R"], GOLD["P"], size)
        assert np.array_equal(xy, GOLD["map1"]) and np.array_equal(fr, GOLD["map2"] & 1023)


def test_euroc_size_live_cv2():
    cv2 = pytest.importorskip("cv2")
    m1, m2 = cv2.initUndistortRectifyMap(D.EUROC_CAM0_K, D.EUROC_CAM0_DIST, D.EUROC_CAM0_RECT, D.EUROC_CAM0_PROJ, (752, 480),
                                         cv2.CV_16SC2)
    xy, fr = D.undistort_rectify_maps(D.EUROC_CAM0_K, D.EUROC_CAM0_DIST, D.EUROC_CAM0_RECT, D.EUROC_CAM0_PROJ, (752, 480))
    # direct float64 evaluation against OpenCV's incremental row walk: a position on a 1/32-pixel rounding boundary may differ
    assert ((xy != m1).any(axis=2) | (fr != (m2 & 1023))).sum() <= 8
    img = np.random.default_rng(4).integers(0, 256, (480, 752)).astype(np.uint8)
    assert np.array_equal(OF.remap_bilinear_fixed(img, m1, m2), cv2.remap(img, m1, m2, cv2.INTER_LINEAR))
