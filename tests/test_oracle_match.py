"""Oracle matcher vs the cv2 golden vectors (reference call: kitti_ba.cpp:602,641)."""
import numpy as np
import pytest

from oracle import oracle as O

CASES = ["rand", "ties", "one", "kitti600"]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("norm", [O.NORM_HAMMING, O.NORM_HAMMING2])
@pytest.mark.parametrize("cc", [False, True])
def test_bf_match_matches_cv2(golden_match, case, norm, cc):
    q, t = golden_match[f"{case}_q"], golden_match[f"{case}_t"]
    ref = golden_match[f"{case}_n{norm}_cc{int(cc)}"]
    qi, ti, d = O.bf_match(q, t, norm, cc)
    assert np.array_equal(np.stack([qi, ti, d], 1).reshape(-1, 3), ref)


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("norm", [O.NORM_HAMMING, O.NORM_HAMMING2])
def test_knn2_matches_cv2(golden_match, case, norm):
    q, t = golden_match[f"{case}_q"], golden_match[f"{case}_t"]
    idx, d = O.knn2(q, t, norm)
    assert np.array_equal(idx, golden_match[f"{case}_n{norm}_knn_idx"])
    assert np.array_equal(d, golden_match[f"{case}_n{norm}_knn_dist"])


def test_hamming2_definition():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (5, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (7, 32), dtype=np.uint8)
    D = O.hamming_matrix(a, b, O.NORM_HAMMING2)
    for i in range(5):
        for j in range(7):
            x = np.unpackbits(a[i] ^ b[j]).reshape(-1, 2)
            assert D[i, j] == int((x.sum(1) > 0).sum())


def test_empty_sets():
    z = np.zeros((0, 32), dtype=np.uint8)
    one = np.zeros((3, 32), dtype=np.uint8)
    assert len(O.bf_match(z, one)[0]) == 0
    assert len(O.bf_match(one, z)[0]) == 0
