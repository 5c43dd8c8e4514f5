"""BASELINE.json config 3 at its full size (4541 frames x 2000 keypoints = 4540 pairs in one batch) and
config 2 (EuRoC shape): sampled pairs against the CPU oracle, plus properties that do not depend on size."""
import numpy as np
import pytest

from conftest import esame
from epivo_b200 import api, synth
from oracle import pipeline as OP

pytestmark = pytest.mark.gpu


def rot_angle(Ra, Rb):
    return float(np.arccos(np.clip((np.trace(Ra.T @ Rb) - 1) / 2, -1, 1)))


def _check_pair(pipe, res, seq, i, method, thr):
    o = OP.pair_pipeline(seq.kps[i], seq.descs[i], seq.kps[i + 1], seq.descs[i + 1], seq.K, method, 0.99, thr)
    qi, ti, d = pipe.matches(i)
    for g, w in zip((qi, ti, d), o["matches"]):
        assert np.array_equal(g, w)
    em, pm = pipe.masks(i)
    assert np.array_equal(em, o["e_mask"]) and np.array_equal(pm, o["pose_mask"])
    r = res[i]
    assert r["ransac_iters"] == o["e_info"]["iters"] and r["n_good"] == o["n_good"]
    assert esame(r["E"], o["E"]) < 1e-9
    assert rot_angle(r["T"][:3, :3], o["T"][:3, :3]) < 1e-4


def test_full_kitti_sequence_one_batch(ctx):
    seq = synth.make_sequence(4541, 2000, seed=synth.seed_for(3, 0))
    P = seq.n_pairs
    assert P == 4540
    pipe = api.SequencePipeline(seq.n_frames, 2000, ctx=ctx)
    prm = api.default_params(seq.K.astype(np.float32))
    res = pipe.process(prm, seq.kps, seq.descs).copy()          # host buffers in, results out
    # sampled pairs (first, last, a piece boundary of the staged upload, a few in between) vs the oracle
    for i in (0, 221, 222, 1700, 3333, P - 1):
        _check_pair(pipe, res, seq, i, api.RANSAC, 1.0)
    # size-independent properties over all 4540 pairs
    assert (res["n_inliers"] <= res["n_matches"]).all() and (res["n_good"] <= res["n_inliers"]).all()
    assert (res["n_matches"] <= 2000).all() and (res["n_matches"] > 1000).all()
    assert (res["ransac_iters"] >= 1).all() and (res["ransac_iters"] <= 1000).all()
    E = res["E"]
    assert np.allclose(np.linalg.norm(E.reshape(P, 9), axis=1), 1.0, atol=1e-12)                 # unit norm
    assert np.abs(np.linalg.det(E)).max() < 1e-10                                              # rank 2
    R = res["R"]
    assert np.abs(R @ np.transpose(R, (0, 2, 1)) - np.eye(3)).max() < 1e-9                     # rotations
    assert np.allclose(np.linalg.det(R), 1.0, atol=1e-9)
    assert np.allclose(np.linalg.norm(res["t"], axis=1), 1.0, atol=1e-9)
    rot_err = np.array([rot_angle(R[i], seq.R[i]) for i in range(P)])
    assert np.median(rot_err) < 3e-3 and (rot_err < 3e-2).mean() > 0.99                        # close to the ground truth
    # idempotence: the resident-data path gives the same results as the host-buffer path
    pipe.run(prm, 0, P)
    res2 = pipe.download(0, P)
    for f in ("E", "R", "t", "T", "n_matches", "n_inliers", "n_good", "ransac_iters", "lm_iters"):
        assert np.array_equal(res[f], res2[f]), f
    # the pose chain and cloud of the whole run
    poses, pts, limits = pipe.cloud(with_points=False)
    assert poses.shape == (P + 1, 4, 4) and np.isfinite(poses).all() and (np.diff(limits) >= 0).all()
    pipe.close()


def test_euroc_shaped_sequence(ctx):
    """BASELINE config 2: 752x480, 1500 keypoints, RANSAC(0.99, 0.3) as euroc_E.cpp:202-208."""
    seq = synth.make_sequence(9, 1500, seed=synth.seed_for(2, 3), K=synth.EUROC_K, size=synth.EUROC_SIZE,
                              depth=(1.0, 8.0), px_sigma=0.3, outlier_frac=0.25)
    pipe = api.SequencePipeline(seq.n_frames, 1500, ctx=ctx)
    prm = api.default_params(seq.K.astype(np.float32), threshold=0.3, fallback_t=(0.0, 0.0, 1.0))   # euroc_E.cpp:264-271
    res = pipe.process(prm, seq.kps, seq.descs).copy()
    for i in range(seq.n_pairs):
        o = OP.pair_pipeline(seq.kps[i], seq.descs[i], seq.kps[i + 1], seq.descs[i + 1], seq.K, api.RANSAC, 0.99, 0.3,
                             fallback_t=(0.0, 0.0, 1.0))
        em, pm = pipe.masks(i)
        assert np.array_equal(em, o["e_mask"]) and np.array_equal(pm, o["pose_mask"])
        assert res[i]["ransac_iters"] == o["e_info"]["iters"]
        assert rot_angle(res[i]["T"][:3, :3], o["T"][:3, :3]) < 1e-4
    pipe.close()
