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


def test_level_geometry_of_the_library_matches_the_oracle():
    """The host-side set-up of epivo_orb_detect_and_compute (level sizes by the float reciprocal, feature budgets in
    float) is the oracle's for random configurations -- incl. 285 columns at scale 1.2f, which are 238 (not 237) wide on
    level 1 -- and refuses what OpenCV asserts on (a level that rounds to an empty image)."""
    from epivo_b200 import api
    r, c, f = api.orbLevelGeometry(174, 285, 10000, 1.2, 3)
    assert r.tolist() == [174, 145, 121] and c.tolist() == [285, 238, 198] and f.tolist() == [3956, 3297, 2747]
    rng = np.random.default_rng(8)
    refused = 0
    for _ in range(4000):
        rows, cols = int(rng.integers(1, 2200)), int(rng.integers(1, 2200))
        nf = int(rng.integers(0, 20001))
        sc = float(np.float32(rng.choice([1.05, 1.1, 1.2, 1.25, 1.3, 1.41, 1.5, 1.7, 2.0])))
        nl = int(rng.integers(1, 17))
        sizes = OO.layer_sizes(rows, cols, OO.layer_scales(nl, sc))
        if min(min(a, b) for a, b in sizes) < 1:
            with pytest.raises(Exception):
                api.orbLevelGeometry(rows, cols, nf, sc, nl)
            refused += 1
            continue
        r, c, f = api.orbLevelGeometry(rows, cols, nf, sc, nl)
        assert [tuple(x) for x in zip(r.tolist(), c.tolist())] == sizes, (rows, cols, sc, nl)
        assert f.tolist() == OO.features_per_level(nf, nl, sc), (nf, sc, nl)
    assert refused > 0
    for bad in [(100, 100, 500, 1.0, 8), (100, 100, 500, 2.5, 8), (100, 100, 500, 1.2, 0), (100, 100, 500, 1.2, 17)]:
        with pytest.raises(Exception):
            api.orbLevelGeometry(*bad)
