"""epivo_seq_process_points: the geometry of the loop body (kitti_E.cpp:96-201) for correspondences the caller already
has -- the LK tracks of kitti_E.cpp:86-95 -- instead of descriptor matches.  Same kernels as the descriptor pipeline:
fed with that pipeline's own matches it must return the same bytes; ragged / tiny / empty pairs against the oracle."""
import numpy as np
import pytest

from epivo_b200 import api, synth
from oracle import pipeline as OP

pytestmark = pytest.mark.gpu


def test_points_pipeline_equals_descriptor_pipeline(ctx):
    s = synth.make_sequence(n_frames=41, n=2000, seed=synth.seed_for(3, 9))
    Kf = s.K.astype(np.float32)
    for kw in (dict(), dict(method=api.LMEDS, threshold=0.01), dict(threshold=0.3)):
        prm = api.default_params(Kf, **kw)
        pipe = api.SequencePipeline(s.n_frames, 2000, ctx=ctx)
        pipe.upload(s.kps, s.descs)
        pipe.run(prm, 0, s.n_pairs)
        ref = pipe.download(0, s.n_pairs).copy()
        p0, p1 = [], []
        for i in range(s.n_pairs):
            qi, ti, _ = pipe.matches(i)
            p0.append(s.kps[i][qi])
            p1.append(s.kps[i + 1][ti])
        masks = [pipe.masks(i) for i in range(3)]
        pipe2 = api.SequencePipeline(s.n_frames, 2000, ctx=ctx)
        out = pipe2.process_points(prm, p0, p1)
        assert out.tobytes() == ref.tobytes(), kw
        for i in range(3):
            em, pm = pipe2.masks(i)
            assert np.array_equal(em, masks[i][0]) and np.array_equal(pm, masks[i][1])
        pipe.close()
        pipe2.close()


def test_points_pipeline_ragged_vs_oracle(ctx):
    s = synth.make_sequence(n_frames=7, n=600, seed=synth.seed_for(3, 11))
    Kf = s.K.astype(np.float32)
    prm = api.default_params(Kf)
    pipe = api.SequencePipeline(s.n_frames, 600, ctx=ctx)
    pipe.upload(s.kps, s.descs)
    pipe.run(prm, 0, s.n_pairs)
    p0, p1 = [], []
    for i in range(s.n_pairs):
        qi, ti, _ = pipe.matches(i)
        keep = [len(qi), 300, 60, 5, 4, 0][i]                      # full, ragged, below the LM's 48, minimal, too few, empty
        p0.append(s.kps[i][qi][:keep])
        p1.append(s.kps[i + 1][ti][:keep])
    pipe.close()
    pipe2 = api.SequencePipeline(s.n_frames, 600, ctx=ctx)
    out = pipe2.process_points(prm, p0, p1)
    for i in range(s.n_pairs):
        o = OP.points_pipeline(p0[i], p1[i], Kf)
        assert out[i]["n_matches"] == len(p0[i])
        em, pm = pipe2.masks(i)
        if o["E"] is None or np.asarray(o["E"]).shape != (3, 3):   # < 5 points, or the stacked solutions of exactly 5
            if len(p0[i]) < 5:
                assert out[i]["n_inliers"] == 0 and out[i]["lm_ran"] == 0
            continue
        assert np.array_equal(em, o["e_mask"]), i
        assert np.array_equal(pm, o["pose_mask"]), i
        assert out[i]["n_good"] == o["n_good"] and bool(out[i]["lm_ran"]) == o["lm_ran"]
        assert np.abs(out[i]["T"] - o["T"]).max() < 1e-6
    with pytest.raises(Exception):
        pipe2.process_points(prm, [np.zeros((700, 2), np.float32)], [np.zeros((700, 2), np.float32)])   # more points than kp_per_frame
    pipe2.close()


def test_front_end_into_geometry_runs(ctx):
    """kitti_E.cpp:54-201 from frames: FAST(40) -> LK -> status filter -> findEssentialMat(LMEDS) -> recoverPose -> LM,
    every stage on the device; the stages are pinned individually elsewhere, this checks the chain hands over cleanly."""
    rng = np.random.default_rng(8)
    tex = rng.integers(0, 256, (300, 500)).astype(np.float32)
    for _ in range(2):
        tex = sum(np.roll(np.roll(tex, dy, 0), dx, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1)) / 9.0
    tex = ((tex - tex.min()) / (tex.max() - tex.min()) * 255).astype(np.uint8)
    frames = np.stack([tex[20 + k:260 + k, 30 + 2 * k:430 + 2 * k] for k in range(4)])
    det = api.fastDetect(frames[:-1], 40, True, ctx=ctx)
    nxt, st = api.trackSequenceLK(frames, [d[0] for d in det], ctx=ctx)
    p0 = [det[i][0][st[i] == 1] for i in range(3)]
    p1 = [nxt[i][st[i] == 1] for i in range(3)]
    cap = max(len(p) for p in p0)
    assert cap > 200
    K = np.array([[400.0, 0, 200.0], [0, 400.0, 120.0], [0, 0, 1]], np.float32)
    pipe = api.SequencePipeline(4, cap, ctx=ctx)
    out = pipe.process_points(api.default_params(K, method=api.LMEDS, threshold=0.01), p0, p1)
    assert (out["n_matches"] == [len(p) for p in p0]).all()
    assert (out["n_inliers"] > 0.3 * out["n_matches"]).all() and np.isfinite(out["T"]).all()
    pipe.close()
