"""The numpy restatement of cv::ORB::detectAndCompute (oracle/orb.py) against cv2: the committed golden vectors
(tests/golden/orb.npz, cv2 4.13.0) and, where cv2 is installed, live on fresh scenes.  Every keypoint field (position,
order, size, angle, response, octave) and every descriptor byte must be identical."""
import os

import numpy as np
import pytest

from epivo_b200.orb_pattern import BIT_PATTERN_31
from oracle import orb as OO
from orb_util import cv2_orb, kps_array, scene

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "orb.npz"))
NAMES = sorted(k[4:] for k in GOLD.files if k.startswith("img_"))


def _cfg(name):
    nf, sc, nl, edge, thr = GOLD["cfg_" + name]
    return int(nf), float(sc), int(nl), int(edge), int(thr)


@pytest.mark.parametrize("name", NAMES)
def test_oracle_orb_matches_cv2_golden(name):
    nf, sc, nl, edge, thr = _cfg(name)
    kps, desc = OO.detect_and_compute(GOLD["img_" + name], BIT_PATTERN_31, nf, sc, nl, edge, thr)
    assert np.array_equal(kps_array(kps), GOLD["kps_" + name])
    assert np.array_equal(desc, GOLD["desc_" + name])


def test_pattern_and_tables():
    assert BIT_PATTERN_31.shape == (256, 4) and BIT_PATTERN_31[0].tolist() == [8, -3, 9, 5]
    assert BIT_PATTERN_31[-1].tolist() == [-1, -6, 0, -11] and np.abs(BIT_PATTERN_31).max() == 13
    # the constants orb.cu hard-codes
    assert OO.umax_table() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert [float(v).hex() for v in OO.GAUSS7[:4]] == ["0x1.1f5f620000000p-4", "0x1.0c70fc0000000p-3", "0x1.8694720000000p-3",
                                                       "0x1.ba95c00000000p-3"]
    assert OO.features_per_level(10000, 8, 1.2) == [2172, 1810, 1508, 1257, 1047, 873, 727, 606]
    inc = open(os.path.join(os.path.dirname(__file__), "..", "epivo_b200", "csrc", "orb_pattern.inc")).read()
    vals = [int(v) for line in inc.splitlines() if not line.startswith("//") for v in line.replace(",", " ").split()]
    assert np.array_equal(np.array(vals).reshape(256, 4), BIT_PATTERN_31)


def test_oracle_pieces_vs_live_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (97, 131)).astype(np.uint8)
    for drows, dcols in [(81, 109), (37, 64), (96, 130), (20, 11)]:
        assert np.array_equal(OO.resize_linear_exact(img, drows, dcols),
                              cv2.resize(img, (dcols, drows), interpolation=cv2.INTER_LINEAR_EXACT))
    assert np.array_equal(OO.GAUSS7, cv2.getGaussianKernel(7, 2, cv2.CV_32F).ravel())
    for _ in range(2000):
        y, x = (np.float32(v) for v in rng.integers(-200000, 200000, 2))
        assert OO.fast_atan2(y, x) == np.float32(cv2.fastAtan2(float(y), float(x)))
    assert OO.fast_atan2(0, 0) == 0 and OO.fast_atan2(0, -3) == 180 and OO.fast_atan2(-2, 0) == 270


def test_oracle_orb_vs_live_cv2():
    cv2 = pytest.importorskip("cv2")
    for rows, cols, seed, nf in [(376, 1241, 21, 10000), (240, 376, 22, 700), (90, 111, 23, 50)]:
        img = scene(rows, cols, seed)
        ref_k, ref_d = cv2_orb(cv2, img, nf)
        kps, desc = OO.detect_and_compute(img, BIT_PATTERN_31, nf)
        assert np.array_equal(kps_array(kps), ref_k), (rows, cols, nf)
        assert np.array_equal(desc, ref_d), (rows, cols, nf)
