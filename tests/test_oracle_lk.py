"""LK tracker restatement (oracle/frontend.py) against cv2: pyramid and derivatives exactly, tracked points by the
rule of tests/lk_util.py -- on the committed golden vectors and, where cv2 is importable, live."""
import os

import numpy as np
import pytest

from oracle import frontend as OF
from lk_util import check_err, check_lk

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "lk.npz"))
NAMES = sorted(k[5:] for k in GOLD.files if k.startswith("prev_"))


@pytest.mark.parametrize("name", NAMES)
def test_oracle_lk_matches_cv2_golden(name):
    nxt, st, err = OF.calc_optical_flow_pyr_lk(GOLD["prev_" + name], GOLD["next_" + name], GOLD["pts_" + name], return_err=True)
    check_lk(nxt, st, GOLD["out_" + name], GOLD["status_" + name], name)
    check_err(err, st, GOLD["err_" + name], GOLD["status_" + name], name)


def test_oracle_pyramid_and_derivatives_exact():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(9)
    for rows, cols in [(47, 60), (64, 64), (95, 121), (376, 1241)]:
        im = rng.integers(0, 256, (rows, cols)).astype(np.uint8)
        assert np.array_equal(OF.pyr_down(im), cv2.pyrDown(im))
        d = OF.scharr_deriv(im).astype(np.int32)
        gx = cv2.Scharr(im, cv2.CV_16S, 1, 0, borderType=cv2.BORDER_REFLECT_101).astype(np.int32)
        gy = cv2.Scharr(im, cv2.CV_16S, 0, 1, borderType=cv2.BORDER_REFLECT_101).astype(np.int32)
        assert np.array_equal(d[:, :, 0], gx) and np.array_equal(d[:, :, 1], gy)
    assert len(OF.build_pyramid(np.zeros((376, 1241), np.uint8))) == 4          # 1241x376, 621x188, 311x94, 156x47
    assert len(OF.build_pyramid(np.zeros((47, 60), np.uint8))) == 2             # 24 x 30 is the last: 12 x 15 <= window


def test_oracle_lk_matches_live_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(31)
    base = cv2.GaussianBlur(rng.integers(0, 256, (300, 500)).astype(np.uint8), (0, 0), 2.0)
    base = cv2.normalize(base, None, 0, 255, cv2.NORM_MINMAX)
    src = base[10:250, 10:450].copy()
    M = np.array([[1.01, 0.002, 3.2], [-0.003, 1.008, -1.7]], np.float32)
    tgt = cv2.warpAffine(base, M, (500, 300))[10:250, 10:450].copy()
    pts = np.array([k.pt for k in cv2.FastFeatureDetector_create(15).detect(src, None)], np.float32)[:1500]
    ref, st, _ = cv2.calcOpticalFlowPyrLK(src, tgt, pts, None)
    nxt, st2 = OF.calc_optical_flow_pyr_lk(src, tgt, pts)
    check_lk(nxt, st2, ref, st, "live")


def test_oracle_lk_points_outside_the_image_live_cv2():
    """Points outside the frame (lost at level 0), on its edges (tracked through the reflected border) and far away."""
    cv2 = pytest.importorskip("cv2")
    a, b = GOLD["prev_affine"], GOLD["next_affine"]
    h, w = a.shape
    pts = np.array([[-5, -5], [-30, 10], [w + 5, 10], [w + 40, h + 40], [w - 1, h - 1], [0, 0], [1e6, 1e6], [-1e6, 5],
                    [w - 0.5, h - 0.5], [10.5, -12.25], [w + 9.75, h / 2]], np.float32)
    ref, st, _ = cv2.calcOpticalFlowPyrLK(a, b, pts, None)
    nxt, st2 = OF.calc_optical_flow_pyr_lk(a, b, pts)
    assert np.array_equal(st.ravel(), st2) and 0 < st2.sum() < len(pts)
    assert np.abs(ref.reshape(-1, 2) - nxt).max() <= 1e-3 * max(1.0, np.abs(nxt).max() * 1e-6)     # lost tracks included
