"""N4 front end, undistortion: epivo_remap against cv2.remap with the fixed-point maps of initUndistortRectifyMap
(euroc_E.cpp:105-113,169-174) -- golden vectors, the restatement on random maps, and the EuRoC loop head
(remap -> FAST(10) -> LK, euroc_E.cpp:169-196) against live cv2.  Byte work: bit-exact."""
import os

import numpy as np
import pytest

from epivo_b200 import api
from epivo_b200 import datasets as D
from lk_util import check_lk
from oracle import frontend as OF

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "remap.npz"))
pytestmark = pytest.mark.gpu


def test_gpu_remap_matches_cv2_golden(ctx):
    assert np.array_equal(api.remap(GOLD["img"], GOLD["map1"], GOLD["map2"], ctx=ctx), GOLD["out"])
    assert np.array_equal(api.remap(GOLD["img"], GOLD["rmap1"], GOLD["rmap2"], ctx=ctx), GOLD["rout0"])
    assert np.array_equal(api.remap(GOLD["img"], GOLD["rmap1"], GOLD["rmap2"], borderValue=77, ctx=ctx), GOLD["rout77"])


def test_gpu_remap_batch_random_maps_vs_oracle(ctx):
    rng = np.random.default_rng(17)
    ims = rng.integers(0, 256, (3, 45, 61)).astype(np.uint8)
    m1 = rng.integers(-70, 130, (50, 33, 2)).astype(np.int16)          # far outside on every side, 1-pixel sources included
    m2 = rng.integers(0, 65536, (50, 33)).astype(np.uint16)            # only the low 10 bits count
    out = api.remap(ims, m1, m2, borderValue=9, ctx=ctx)
    for i in range(3):
        assert np.array_equal(out[i], OF.remap_bilinear_fixed(ims[i], m1, m2, 9))
    one = api.remap(ims[0, :1, :1], m1, m2, ctx=ctx)                   # a 1 x 1 source: every pixel takes the border path
    assert np.array_equal(one, OF.remap_bilinear_fixed(ims[0, :1, :1], m1, m2))
    assert ctx.lib.epivo_remap(ctx.h, api._p(ims), 1, 45, 61, api._p(m1), api._p(m2), 50, 33, 300, api._p(out)) != 0


def test_gpu_euroc_loop_head_vs_live_cv2(ctx):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(2200)
    base = cv2.GaussianBlur(rng.integers(0, 256, (560, 840)).astype(np.uint8), (0, 0), 1.8)
    base = cv2.normalize(base, None, 0, 255, cv2.NORM_MINMAX)
    raw = np.stack([base[40 + k:520 + k, 44 + 2 * k:796 + 2 * k] for k in range(3)])       # 752 x 480 "distorted" frames
    xy, fr = D.undistort_rectify_maps(D.EUROC_CAM0_K, D.EUROC_CAM0_DIST, D.EUROC_CAM0_RECT, D.EUROC_CAM0_PROJ, (752, 480))
    m1, m2 = cv2.initUndistortRectifyMap(D.EUROC_CAM0_K, D.EUROC_CAM0_DIST, D.EUROC_CAM0_RECT, D.EUROC_CAM0_PROJ, (752, 480),
                                         cv2.CV_16SC2)
    und = api.remap(raw, m1, m2, ctx=ctx)
    for i in range(3):
        assert np.array_equal(und[i], cv2.remap(raw[i], m1, m2, cv2.INTER_LINEAR))
    assert (api.remap(raw, xy, fr, ctx=ctx) != und).mean() < 1e-4          # own maps: at most a few boundary pixels apart
    det = api.fastDetect(und[:-1], 10, True, ctx=ctx)                      # euroc_E.cpp:177: FastFeatureDetector::create()
    nxt, st = api.trackSequenceLK(und, [d[0] for d in det], ctx=ctx)
    cvdet = cv2.FastFeatureDetector_create()
    for i in range(2):
        ref_pts = np.array([k.pt for k in cvdet.detect(und[i], None)], np.float32).reshape(-1, 2)
        assert np.array_equal(det[i][0], ref_pts) and len(ref_pts) > 300
        ref, rst, _ = cv2.calcOpticalFlowPyrLK(und[i], und[i + 1], ref_pts, None)
        check_lk(nxt[i], st[i], ref, rst, "euroc pair %d" % i)
