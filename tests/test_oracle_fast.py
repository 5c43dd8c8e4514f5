"""FAST-9/16 restatement (oracle/frontend.py) against cv2: the committed golden vectors and, where cv2 is importable,
live on fresh random images (threshold 40 = kitti_E.cpp:71, threshold 10 = kitti_ba.cpp:98's default)."""
import os

import numpy as np
import pytest

from oracle import frontend as OF

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "fast.npz"))
NAMES = sorted(k[4:] for k in GOLD.files if k.startswith("img_"))


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("thr", [10, 40])
@pytest.mark.parametrize("nms", [True, False])
def test_oracle_fast_matches_cv2_golden(name, thr, nms):
    pts, resp = OF.fast_detect(GOLD["img_" + name], thr, nms)
    assert np.array_equal(pts, GOLD[f"pts_{name}_{thr}_{int(nms)}"])
    assert np.array_equal(resp, GOLD[f"resp_{name}_{thr}_{int(nms)}"])


def test_oracle_fast_matches_live_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    for k in range(6):
        rows, cols = int(rng.integers(7, 90)), int(rng.integers(7, 130))
        im = rng.integers(0, 256, (rows, cols)).astype(np.uint8)
        if k % 2:
            im = cv2.GaussianBlur(im, (0, 0), 1.2)
        for thr in (0, 7, 40, 255):
            for nms in (True, False):
                kps = cv2.FastFeatureDetector_create(thr, nms).detect(im, None)
                ref = np.array([q.pt for q in kps], dtype=np.float32).reshape(-1, 2)
                rr = np.array([q.response for q in kps], dtype=np.float32)
                pts, resp = OF.fast_detect(im, thr, nms)
                assert np.array_equal(pts, ref), (rows, cols, thr, nms)
                assert np.array_equal(resp, rr), (rows, cols, thr, nms)
