"""N4 front end, descriptor extractor: epivo_orb_detect_and_compute against cv2's ORB as kitti_ba.cpp:128-152 runs it
(ORB::create(10000, 1.2f, 8, 15, 0, 2, FAST_SCORE); detect; compute) -- the committed cv2 golden vectors, the numpy
restatement on other shapes, live cv2 on KITTI-sized frames, and the front end -> matcher chain.  Byte / integer work:
keypoints (position, order, size, angle, response, octave) and descriptors must be identical."""
import os

import numpy as np
import pytest

from epivo_b200 import api
from epivo_b200.orb_pattern import BIT_PATTERN_31
from oracle import orb as OO
from orb_util import FIELDS, cv2_orb, kps_array, scene

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "orb.npz"))
NAMES = sorted(k[4:] for k in GOLD.files if k.startswith("img_"))


def _report(kps, desc, ref_k, ref_d, what):
    """Stage by stage, so that a failure names the first stage that differs."""
    got = kps_array(kps)
    by_level = lambda a: np.bincount(a[:, 5].astype(int), minlength=16).tolist()
    assert len(got) == len(ref_k), f"{what}: {len(got)} keypoints, expected {len(ref_k)}; per level {by_level(got)} vs {by_level(ref_k)}"
    for c, f in enumerate(FIELDS):
        if f == "angle":
            continue
        bad = np.nonzero(got[:, c] != ref_k[:, c])[0]
        assert not len(bad), f"{what}: field {f} differs at {len(bad)} keypoints, first {bad[0]}: {got[bad[0]]} vs {ref_k[bad[0]]}"
    bad = np.nonzero(got[:, 3] != ref_k[:, 3])[0]
    assert not len(bad), f"{what}: angle differs at {len(bad)} keypoints, first {bad[0]}: {got[bad[0], 3]!r} vs {ref_k[bad[0], 3]!r}"
    bad = np.nonzero((desc != ref_d).any(axis=1))[0]
    assert not len(bad), (f"{what}: {len(bad)} of {len(ref_d)} descriptors differ (levels {np.unique(ref_k[bad, 5]).tolist()}), "
                          f"first {bad[0]}: {np.unpackbits(desc[bad[0]] ^ ref_d[bad[0]]).sum()} bits")
    assert (kps["class_id"] == -1).all()


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_gpu_orb_matches_cv2_golden(ctx, name):
    nf, sc, nl, edge, thr = GOLD["cfg_" + name]
    kps, desc = api.orbDetectAndCompute(GOLD["img_" + name], int(nf), float(sc), int(nl), int(edge), int(thr), ctx=ctx)
    _report(kps, desc, GOLD["kps_" + name], GOLD["desc_" + name], name)


@pytest.mark.gpu
def test_gpu_orb_batch_vs_oracle(ctx):
    """A batch of frames in one call, odd sizes, a flat frame (no corners) inside the batch."""
    # 174 x 285: level 1 is 238 columns wide (cols * (1.f / 1.2f) = 237.5), 237 by division -- OpenCV multiplies
    for rows, cols, nf in [(97, 131, 10000), (64, 200, 120), (174, 285, 10000), (33, 33, 500), (8, 9, 500), (3, 3, 500)]:
        ims = np.stack([scene(rows, cols, 40 + i) for i in range(4)])
        ims[2] = 90
        out = api.orbDetectAndCompute(ims, nf, ctx=ctx)
        assert len(out) == 4 and len(out[2][0]) == 0 and out[2][1].shape == (0, 32)          # upper levels may be 1 x 1
        for i in range(4):
            kps, desc = OO.detect_and_compute(ims[i], BIT_PATTERN_31, nf)
            _report(out[i][0], out[i][1], kps_array(kps), desc, f"{rows}x{cols} nf={nf} image {i}")


@pytest.mark.gpu
def test_gpu_orb_buffer_smaller_than_result(ctx):
    img = GOLD["img_kba_small"]
    ref_k, ref_d = GOLD["kps_kba_small"], GOLD["desc_kba_small"]
    kps, desc = api.orbDetectAndCompute(img, max_keypoints=100, ctx=ctx)
    _report(kps, desc, ref_k[:100], ref_d[:100], "first 100")
    kps, desc = api.orbDetectAndCompute(img, max_keypoints=0, ctx=ctx)
    assert len(kps) == 0
    with pytest.raises(Exception):
        api.orbDetectAndCompute(img, edgeThreshold=10, ctx=ctx)          # the orientation patch would leave the level
    with pytest.raises(Exception):
        api.orbDetectAndCompute(img, nlevels=0, ctx=ctx)


@pytest.mark.gpu
def test_gpu_orb_kitti_size_vs_live_cv2(ctx):
    """1241 x 376 frames with the reference's 10000-feature setting, a batch of three; and a budget that cuts every level."""
    cv2 = pytest.importorskip("cv2")
    ims = np.stack([scene(376, 1241, 60 + i) for i in range(3)])
    for nf in (10000, 1500):
        out = api.orbDetectAndCompute(ims, nf, ctx=ctx)
        for i in range(3):
            ref_k, ref_d = cv2_orb(cv2, ims[i], nf)
            assert len(ref_k) > nf * 0.9
            _report(out[i][0], out[i][1], ref_k, ref_d, f"kitti frame {i} nf={nf}")


@pytest.mark.gpu
def test_gpu_orb_feeds_the_matcher(ctx):
    """kitti_ba.cpp:128-152 -> :602,641: ORB on two frames, BFMatcher(NORM_HAMMING2, crossCheck) on the descriptors;
    the second frame is the first shifted by (5, 2) pixels, so mutual matches must sit on that shift."""
    big = scene(300, 500, 77)
    f0, f1 = big[10:250, 10:410].copy(), big[8:248, 5:405].copy()      # f1(x, y) = f0(x - 5, y - 2)
    (k0, d0), (k1, d1) = api.orbDetectAndCompute(np.stack([f0, f1]), ctx=ctx)
    assert len(k0) > 500 and len(k1) > 500
    qi, ti, dist = api.BFMatcher(api.NORM_HAMMING2, True, ctx=ctx).match(d0, d1)
    dx = k1["x"][ti] - k0["x"][qi]
    dy = k1["y"][ti] - k0["y"][qi]
    on = (np.abs(dx - 5) < 2.5) & (np.abs(dy - 2) < 2.5)
    assert len(qi) > 300 and on.mean() > 0.8, (len(qi), on.mean())


@pytest.mark.gpu
def test_gpu_orb_into_the_sequence_on_device(ctx):
    """epivo_seq_extract_orb (frames -> ORB -> frame slots without leaving the device) gives the same bytes as
    orbDetectAndCompute + KeyPoint::convert + upload + set_counts: matches, masks and poses of every pair identical; a
    slot smaller than the result keeps the first kp_per_frame keypoints."""
    from epivo_b200 import synth
    big = scene(260, 420, 91)
    frames = np.stack([big[2 * k:2 * k + 200, 3 * k:3 * k + 360] for k in range(4)])
    prm = api.default_params(synth.KITTI_K.astype(np.float32), method=api.RANSAC, prob=0.99, threshold=0.3)
    for cap, nf in [(4096, 10000), (600, 10000), (4096, 300)]:
        feats = api.orbDetectAndCompute(frames, nf, ctx=ctx)
        counts = np.array([min(len(k), cap) for k, _ in feats], dtype=np.int32)
        kps = np.zeros((4, cap, 2), np.float32)
        descs = np.zeros((4, cap, 32), np.uint8)
        for i, (k, d) in enumerate(feats):
            kps[i, :counts[i], 0], kps[i, :counts[i], 1] = k["x"][:counts[i]], k["y"][:counts[i]]
            descs[i, :counts[i]] = d[:counts[i]]
        a = api.SequencePipeline(4, cap, ctx=ctx)
        a.upload(kps, descs)
        a.set_counts(counts)
        a.run(prm, 0, 3)
        ra = a.download(0, 3)
        b = api.SequencePipeline(4, cap, ctx=ctx)
        found = b.extract_orb(frames[:3], 0, nf)                         # two calls: slots 0-2, then slot 3
        found = np.concatenate([found, b.extract_orb(frames[3:], 3, nf)])
        assert np.array_equal(found, [len(k) for k, _ in feats])
        b.run(prm, 0, 3)
        rb = b.download(0, 3)
        assert ra.tobytes() == rb.tobytes(), (cap, nf)
        assert ra["n_matches"].min() > 50
        for i in range(3):
            for x, y in zip(a.matches(i), b.matches(i)):
                assert np.array_equal(x, y)
            for x, y in zip(a.masks(i), b.masks(i)):
                assert np.array_equal(x, y)
        a.close()
        b.close()
