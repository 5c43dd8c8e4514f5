"""8-point hypotheses (north_star: "5-point (and 8-point)"; the reference has no call site -- every findEssentialMat
call runs OpenCV's 5-point estimator -- so the oracle is the textbook algorithm in numpy): GPU against the oracle, the
true essential matrix on noiseless scenes, and the hypothesize-and-score loop with the RANSAC scorer (K3)."""
import numpy as np
import pytest

from epivo_b200 import synth
from oracle import oracle as O


def _scene(seed, n=400, noise=0.0):
    pr = synth.make_kitti_pair(seed, n)
    K = pr.K
    rng = np.random.default_rng(seed)
    R = synth.rodrigues(rng.normal(0, 0.03, 3))
    t = np.array([0.1, -0.05, 1.0]) + rng.normal(0, 0.05, 3)
    t /= np.linalg.norm(t)
    X = np.stack([rng.uniform(-20, 20, n), rng.uniform(-6, 6, n), rng.uniform(8, 60, n)], axis=1)
    Y = X @ R.T + t
    x1 = X[:, :2] / X[:, 2:3]
    x2 = Y[:, :2] / Y[:, 2:3] + noise * rng.normal(size=(n, 2))
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    E = tx @ R
    return x1, x2, E / np.linalg.norm(E), K


def _same(a, b):
    return min(np.abs(a - b).max(), np.abs(a + b).max())


def test_oracle_eight_point_recovers_the_true_model():
    x1, x2, E, _ = _scene(3)
    for k in range(5):
        idx = np.arange(8) + 8 * k
        assert _same(O.eight_point(x1[idx], x2[idx]), E) < 1e-9
    assert O.eight_point(np.zeros((8, 2)), np.zeros((8, 2))) is None          # degenerate: rank-deficient system


@pytest.mark.gpu
def test_gpu_eight_point_vs_oracle_and_truth(ctx):
    from epivo_b200 import api
    x1, x2, E, _ = _scene(5, noise=2e-4)
    rng = np.random.default_rng(1)
    S = np.stack([rng.choice(len(x1), 8, replace=False) for _ in range(500)])
    Eg, ok = api.eightPoint(x1[S], x2[S], ctx=ctx)
    assert ok.all()
    for i in range(0, 500, 7):
        Eo = O.eight_point(x1[S[i]], x2[S[i]])
        assert _same(Eg[i], Eo) < 1e-9 * max(1.0, 1.0)
        assert abs(np.linalg.norm(Eg[i]) - 1.0) < 1e-12
        s = np.linalg.svd(Eg[i], compute_uv=False)
        assert abs(s[0] - s[1]) < 1e-12 and s[2] < 1e-12                    # on the essential manifold
    x1c, x2c, Ec, _ = _scene(6)                                               # noiseless: every sample gives the truth
    Eg, ok = api.eightPoint(x1c[:64].reshape(8, 8, 2), x2c[:64].reshape(8, 8, 2), ctx=ctx)
    assert ok.all() and max(_same(e, Ec) for e in Eg) < 1e-8
    Eg, ok = api.eightPoint(np.zeros((3, 8, 2)), np.zeros((3, 8, 2)), ctx=ctx)
    assert not ok.any() and np.all(Eg == 0)


@pytest.mark.gpu
def test_gpu_eight_point_hypotheses_scored_by_k3(ctx):
    """Hypothesize with the 8-point kernel, score with the Sampson scorer: the best of 256 samples on a pair with 30 %
    outliers explains the inliers, and the scorer's counts equal the oracle's for the same models."""
    from epivo_b200 import api
    pr = synth.make_kitti_pair(2, 1200)
    Kf = pr.K.astype(np.float32)
    qi, ti, _ = O.bf_match(pr.desc0, pr.desc1)
    p0, p1 = pr.kp0[qi], pr.kp1[ti]
    x1, x2 = O.normalize_points(p0, Kf), O.normalize_points(p1, Kf)
    rng = np.random.default_rng(4)
    S = np.stack([rng.choice(len(x1), 8, replace=False) for _ in range(256)])
    Eg, ok = api.eightPoint(x1[S], x2[S], ctx=ctx)
    counts, _, best, mask = api.scoreSampson(Eg[ok == 1], p0, p1, Kf, 1.0, ctx=ctx, medians=False)
    t32 = O.ransac_threshold(1.0, Kf)
    for i in range(0, int(ok.sum()), 17):
        assert counts[i] == int(O.find_inliers(O.sampson_err_f32(Eg[ok == 1][i], x1, x2), t32).sum())
    assert counts[best] == counts.max() and counts.max() > 0.5 * len(p0) and int(mask.sum()) == counts[best]
