"""K1 parity: CUDA matcher (through the C ABI) vs the oracle and the cv2 golden vectors.
Bit-exact: indices and integer distances."""
import numpy as np
import pytest

from epivo_b200 import api, synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu
CASES = ["rand", "ties", "one", "kitti600"]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("norm", [api.NORM_HAMMING, api.NORM_HAMMING2])
@pytest.mark.parametrize("cc", [False, True])
def test_match_golden(ctx, golden_match, case, norm, cc):
    q, t = golden_match[f"{case}_q"], golden_match[f"{case}_t"]
    ref = golden_match[f"{case}_n{norm}_cc{int(cc)}"]
    qi, ti, d = api.BFMatcher(norm, cc, ctx=ctx).match(q, t)
    assert np.array_equal(np.stack([qi, ti, d], 1).reshape(-1, 3), ref)


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("norm", [api.NORM_HAMMING, api.NORM_HAMMING2])
def test_knn2_golden(ctx, golden_match, case, norm):
    q, t = golden_match[f"{case}_q"], golden_match[f"{case}_t"]
    idx, d = api.BFMatcher(norm, ctx=ctx).knnMatch2(q, t)
    assert np.array_equal(idx, golden_match[f"{case}_n{norm}_knn_idx"])
    assert np.array_equal(d, golden_match[f"{case}_n{norm}_knn_dist"])


@pytest.mark.parametrize("nq,nt", [(2000, 2000), (1500, 1500), (1, 1), (1023, 1025), (1025, 127), (3000, 129)])
@pytest.mark.parametrize("norm", [api.NORM_HAMMING, api.NORM_HAMMING2])
def test_match_vs_oracle_sizes(ctx, nq, nt, norm):
    rng = np.random.default_rng(nq * 7 + nt)
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    n = min(nq, nt) // 2                      # plant true correspondences so the cross-check keeps many
    t[:n] = synth.flip_bits(q[:n], rng)
    for cc in (False, True):
        got = api.BFMatcher(norm, cc, ctx=ctx).match(q, t)
        want = O.bf_match(q, t, norm, cc)
        for g, w in zip(got, want):
            assert np.array_equal(g, w)
    got = api.BFMatcher(norm, ctx=ctx).ratioMatch(q, t, 0.8)
    want = O.ratio_match(q, t, 0.8, norm)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)


@pytest.mark.parametrize("desc_bytes", [16, 64])
def test_other_descriptor_sizes(ctx, desc_bytes):
    rng = np.random.default_rng(desc_bytes)
    q = rng.integers(0, 256, (300, desc_bytes), dtype=np.uint8)
    t = rng.integers(0, 256, (280, desc_bytes), dtype=np.uint8)
    for norm in (api.NORM_HAMMING, api.NORM_HAMMING2):
        got = api.BFMatcher(norm, True, ctx=ctx).match(q, t)
        want = O.bf_match(q, t, norm, True)
        for g, w in zip(got, want):
            assert np.array_equal(g, w)


def test_empty_and_errors(ctx):
    z = np.zeros((0, 32), dtype=np.uint8)
    one = np.arange(96, dtype=np.uint8).reshape(3, 32)
    assert len(api.BFMatcher(api.NORM_HAMMING2, True, ctx=ctx).match(z, one)[0]) == 0
    assert len(api.BFMatcher(api.NORM_HAMMING2, True, ctx=ctx).match(one, z)[0]) == 0
    qi, ti, d = api.BFMatcher(api.NORM_HAMMING2, True, ctx=ctx).match(one, one)
    assert np.array_equal(qi, [0, 1, 2]) and np.array_equal(ti, [0, 1, 2]) and not d.any()
    with pytest.raises(api.EpivoError):
        api.BFMatcher(api.NORM_HAMMING, ctx=ctx).match(np.zeros((2, 20), np.uint8), np.zeros((2, 20), np.uint8))
    with pytest.raises(api.EpivoError):
        api.BFMatcher(4, ctx=ctx).match(one, one)          # NORM_L2 is not a binary norm


def test_large_orb_setting_properties(ctx):
    """10000 x 10000 (ORB::create(10000), kitti_ba.cpp:128): checked through properties."""
    rng = np.random.default_rng(5)
    n = 10000
    q = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    perm = rng.permutation(n)
    t = synth.flip_bits(q, rng)[perm]
    qi, ti, d = api.BFMatcher(api.NORM_HAMMING2, True, ctx=ctx).match(q, t)
    inv = np.empty(n, dtype=np.int64)
    inv[perm] = np.arange(n)
    assert len(qi) > 0.99 * n
    assert (inv[qi] == ti).mean() > 0.999
    assert np.all(np.diff(qi) > 0)                       # sorted by queryIdx, unique
    assert len(np.unique(ti)) == len(ti)                 # mutual NN is one-to-one
    # swapping the roles gives the transposed match set (symmetry of the cross-check)
    qi2, ti2, d2 = api.BFMatcher(api.NORM_HAMMING2, True, ctx=ctx).match(t, q)
    a = set(zip(qi.tolist(), ti.tolist()))
    b = set(zip(ti2.tolist(), qi2.tolist()))
    assert a == b


@pytest.mark.parametrize("norm", [api.NORM_HAMMING2, api.NORM_HAMMING])
def test_large_orb_setting_exact_vs_live_cv2(ctx, norm):
    """10000 x 10000 (ORB::create(10000), kitti_ba.cpp:128) against cv2.BFMatcher itself, run live: every
    (queryIdx, trainIdx, distance) triple of the cross-check match, and the plain nearest neighbour with its ties
    (a quarter of the train set are duplicates with a few bits cleared, so equal distances are common)."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11)
    n = 10000
    q = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    t = synth.flip_bits(q, rng)[rng.permutation(n)]
    t[: n // 4] = t[n // 4: n // 2] & np.uint8(0xFC)            # near-duplicates: ties in the row and column minima
    for cross in (True, False):
        qi, ti, d = api.BFMatcher(norm, cross, ctx=ctx).match(q, t)
        ms = cv2.BFMatcher(norm, cross).match(q, t)
        want = np.array([(m.queryIdx, m.trainIdx, int(m.distance)) for m in ms], dtype=np.int64).reshape(-1, 3)
        got = np.stack([qi, ti, d], axis=1).astype(np.int64)
        assert got.shape == want.shape and np.array_equal(got, want), (norm, cross, got.shape, want.shape)
