"""N4 front end, detector: epivo_fast_detect against cv2.FastFeatureDetector (the call at kitti_E.cpp:71-74 and
kitti_ba.cpp:49,98) -- the committed cv2 golden vectors, the numpy restatement, and live cv2 on KITTI-sized frames.
Coordinates, order and response must be identical (integer / byte work: bit-exact)."""
import os

import numpy as np
import pytest

from epivo_b200 import api
from oracle import frontend as OF

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "fast.npz"))
NAMES = sorted(k[4:] for k in GOLD.files if k.startswith("img_"))


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_fast_matches_cv2_golden(ctx, name):
    im = GOLD["img_" + name]
    for thr in (10, 40):
        for nms in (True, False):
            pts, resp = api.fastDetect(im, thr, nms, ctx=ctx)
            assert np.array_equal(pts, GOLD[f"pts_{name}_{thr}_{int(nms)}"]), (name, thr, nms)
            assert np.array_equal(resp, GOLD[f"resp_{name}_{thr}_{int(nms)}"]), (name, thr, nms)


@pytest.mark.gpu
def test_gpu_fast_batch_vs_oracle_random_sizes(ctx):
    rng = np.random.default_rng(11)
    for rows, cols in [(7, 7), (8, 33), (33, 8), (64, 65), (50, 257)]:
        ims = rng.integers(0, 256, (5, rows, cols)).astype(np.uint8)
        ims[1] = (ims[1] // 64) * 64                                   # few levels: score ties under suppression
        ims[2] = 128
        for thr in (0, 5, 40, 255):
            for nms in (True, False):
                out = api.fastDetect(ims, thr, nms, ctx=ctx)
                assert len(out) == 5
                for i in range(5):
                    pts, resp = OF.fast_detect(ims[i], thr, nms)
                    assert np.array_equal(out[i][0], pts), (rows, cols, thr, nms, i)
                    assert np.array_equal(out[i][1], resp), (rows, cols, thr, nms, i)


@pytest.mark.gpu
def test_gpu_fast_kitti_size_vs_live_cv2(ctx):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(2026)
    frames = []
    for k in range(6):                                                 # 1241 x 376: the KITTI frame of kitti_E.cpp
        im = cv2.GaussianBlur(rng.integers(0, 256, (376, 1241)).astype(np.uint8), (0, 0), 1.0 + 0.3 * k)
        frames.append(cv2.normalize(im, None, 0, 255, cv2.NORM_MINMAX))
    frames = np.stack(frames)
    for thr in (40, 10):                                               # kitti_E.cpp:71 / kitti_ba.cpp:98
        out = api.fastDetect(frames, thr, True, ctx=ctx)
        det = cv2.FastFeatureDetector_create(thr, True)
        for i in range(len(frames)):
            kps = det.detect(frames[i], None)
            ref = np.array([q.pt for q in kps], dtype=np.float32).reshape(-1, 2)
            rr = np.array([q.response for q in kps], dtype=np.float32)
            assert len(ref) > 100
            assert np.array_equal(out[i][0], ref), (thr, i)
            assert np.array_equal(out[i][1], rr), (thr, i)


@pytest.mark.gpu
def test_gpu_fast_capacity_and_errors(ctx):
    rng = np.random.default_rng(3)
    im = rng.integers(0, 256, (40, 60)).astype(np.uint8)
    full, _ = api.fastDetect(im, 20, True, ctx=ctx)
    assert len(full) > 10
    cut, resp = api.fastDetect(im, 20, True, max_keypoints=7, ctx=ctx)     # the first 7 in OpenCV's order
    assert np.array_equal(cut, full[:7]) and len(resp) == 7
    kps = np.zeros((1, 4, 2), np.float32)
    cnt = np.zeros(1, np.int32)
    rc = ctx.lib.epivo_fast_detect(ctx.h, api._p(im), 1, 40, 60, 20, 1, 4, api._p(kps), None, api._p(cnt))
    assert rc == 0 and cnt[0] == len(full)                                 # counts report what was FOUND
    assert ctx.lib.epivo_fast_detect(ctx.h, api._p(im), 1, 40, 60, 300, 1, 4, api._p(kps), None, api._p(cnt)) != 0
    empty, _ = api.fastDetect(np.zeros((5, 20), np.uint8), 10, True, ctx=ctx)   # no pixel has a full circle
    assert empty.shape == (0, 2)
