"""N4 front end, tracker: epivo_lk_track against cv2.calcOpticalFlowPyrLK with the reference's defaults
(kitti_E.cpp:79-84) -- committed cv2 golden vectors, the numpy restatement, live cv2 on a KITTI-sized sequence, and
the whole kitti_E front end (FAST 40 -> LK -> status filter) chained on the GPU.  Rule: tests/lk_util.py."""
import os

import numpy as np
import pytest

from epivo_b200 import api
from lk_util import check_err, check_lk
from oracle import frontend as OF

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "lk.npz"))
NAMES = sorted(k[5:] for k in GOLD.files if k.startswith("prev_"))


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_lk_matches_cv2_golden(ctx, name):
    nxt, st, err = api.calcOpticalFlowPyrLK(GOLD["prev_" + name], GOLD["next_" + name], GOLD["pts_" + name], ctx=ctx,
                                            returnErr=True)
    check_lk(nxt, st, GOLD["out_" + name], GOLD["status_" + name], name)
    check_err(err, st, GOLD["err_" + name], GOLD["status_" + name], name)
    # against the restatement (exact integer sums on both sides): the same points, to float rounding of the update
    o_nxt, o_st, o_err = OF.calc_optical_flow_pyr_lk(GOLD["prev_" + name], GOLD["next_" + name], GOLD["pts_" + name],
                                                     return_err=True)
    assert np.array_equal(st, o_st)
    both = st == 1
    assert np.abs(nxt[both] - o_nxt[both]).max() <= 1e-4
    assert np.abs(err[both] - o_err[both]).max() <= 1e-3 and (err[~both] == 0).all()


@pytest.mark.gpu
def test_gpu_lk_levels_and_criteria(ctx):
    a, b, pts = GOLD["prev_affine"], GOLD["next_affine"], GOLD["pts_affine"][:200]
    for kw in [dict(maxLevel=0), dict(maxLevel=1, maxCount=5), dict(maxLevel=5, epsilon=0.1), dict(minEigThreshold=1e-2)]:
        nxt, st = api.calcOpticalFlowPyrLK(a, b, pts, ctx=ctx, **kw)
        o_nxt, o_st = OF.calc_optical_flow_pyr_lk(a, b, pts, max_level=kw.get("maxLevel", 3), max_count=kw.get("maxCount", 30),
                                                  epsilon=kw.get("epsilon", 0.01),
                                                  min_eig_threshold=kw.get("minEigThreshold", 1e-4))
        assert np.array_equal(st, o_st), kw
        assert np.abs(nxt[st == 1] - o_nxt[st == 1]).max() <= 1e-4, kw
    empty, st = api.calcOpticalFlowPyrLK(a, b, np.zeros((0, 2), np.float32), ctx=ctx)
    assert empty.shape == (0, 2) and st.shape == (0,)
    with pytest.raises(Exception):
        api.calcOpticalFlowPyrLK(a[:20, :20], b[:20, :20], pts, ctx=ctx)        # not larger than the window


@pytest.mark.gpu
def test_gpu_lk_points_outside_the_image(ctx):
    """Out-of-frame, edge and far-away points: status and the positions OpenCV leaves behind, lost tracks included
    (the restatement is pinned to live cv2 on the same points in tests/test_oracle_lk.py)."""
    a, b = GOLD["prev_affine"], GOLD["next_affine"]
    h, w = a.shape
    pts = np.array([[-5, -5], [-30, 10], [w + 5, 10], [w + 40, h + 40], [w - 1, h - 1], [0, 0], [1e6, 1e6], [-1e6, 5],
                    [w - 0.5, h - 0.5], [10.5, -12.25], [w + 9.75, h / 2], [np.inf, 3], [np.nan, np.nan]], np.float32)
    nxt, st = api.calcOpticalFlowPyrLK(a, b, pts, ctx=ctx)
    o_nxt, o_st = OF.calc_optical_flow_pyr_lk(a, b, pts[:11])
    assert np.array_equal(st[:11], o_st)
    assert np.abs(nxt[:11] - o_nxt).max() <= 1e-3
    assert st[11] == 0 and st[12] == 0                   # inf and NaN: cvFloor gives INT_MIN, outside every range test


@pytest.mark.gpu
def test_gpu_kitti_front_end_vs_live_cv2(ctx):
    """kitti_E.cpp:66-95 for a five-frame KITTI-sized sequence: FAST(40) on frame i, LK into frame i + 1, keep
    status == 1 -- the GPU chain against the same cv2 calls."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(404)
    base = cv2.GaussianBlur(rng.integers(0, 256, (460, 1400)).astype(np.uint8), (0, 0), 1.6)
    base = cv2.normalize(base, None, 0, 255, cv2.NORM_MINMAX)
    frames = []
    for k in range(5):                                                        # forward motion: zoom about a point + drift
        s = 1.0 + 0.012 * k
        M = np.array([[s, 0.001 * k, -620 * (s - 1) + 1.5 * k], [-0.001 * k, s, -190 * (s - 1) - 0.8 * k]], np.float32)
        frames.append(cv2.warpAffine(base, M, (1400, 460))[40:416, 60:1301].copy())
    frames = np.stack(frames)
    assert frames.shape == (5, 376, 1241)
    det = api.fastDetect(frames[:-1], 40, True, ctx=ctx)
    pts = [d[0] for d in det]
    nxt, st = api.trackSequenceLK(frames, pts, ctx=ctx)
    cvdet = cv2.FastFeatureDetector_create(40)
    worst = 0.0
    for i in range(4):
        ref_pts = np.array([k.pt for k in cvdet.detect(frames[i], None)], np.float32).reshape(-1, 2)
        assert np.array_equal(pts[i], ref_pts) and len(ref_pts) > 300
        ref, rst, _ = cv2.calcOpticalFlowPyrLK(frames[i], frames[i + 1], ref_pts, None)
        worst = max(worst, check_lk(nxt[i], st[i], ref, rst, "pair %d" % i))
        assert rst.mean() > 0.9
    print("kitti front end: %d points per frame, worst position difference %.2e px" % (len(pts[0]), worst))
