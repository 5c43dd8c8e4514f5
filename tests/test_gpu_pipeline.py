"""The fused sequence pipeline (epivo_seq_*) against the oracle pipeline, pair by pair."""
import numpy as np
import pytest

from conftest import esame
from epivo_b200 import api, synth
from oracle import oracle as O
from oracle import pipeline as OP

pytestmark = pytest.mark.gpu


def rot_angle(Ra, Rb):
    return float(np.arccos(np.clip((np.trace(Ra.T @ Rb) - 1) / 2, -1, 1)))


@pytest.fixture(scope="module")
def seq():
    return synth.make_sequence(n_frames=7, n=2000, seed=synth.seed_for(3, 1))


@pytest.mark.parametrize("method,thr", [(api.RANSAC, 1.0), (api.LMEDS, 0.01)])
def test_sequence_vs_oracle(ctx, seq, method, thr):
    P = seq.n_pairs
    pipe = api.SequencePipeline(seq.n_frames, 2000, ctx=ctx)
    pipe.upload(seq.kps, seq.descs)
    prm = api.default_params(seq.K.astype(np.float32), method=method, threshold=thr)   # cam is a float Mat (kitti_E.cpp:38)
    pipe.run(prm, 0, P)
    res = pipe.download(0, P)
    for i in range(P):
        o = OP.pair_pipeline(seq.kps[i], seq.descs[i], seq.kps[i + 1], seq.descs[i + 1], seq.K, method, 0.99, thr)
        qi, ti, d = pipe.matches(i)
        for g, w in zip((qi, ti, d), o["matches"]):
            assert np.array_equal(g, w)                       # bit-exact match indices
        em, pm = pipe.masks(i)
        assert np.array_equal(em, o["e_mask"])                # bit-exact inlier mask
        assert np.array_equal(pm, o["pose_mask"])
        r = res[i]
        assert r["n_matches"] == len(qi) and r["n_inliers"] == int(o["e_mask"].sum())
        assert r["n_good"] == o["n_good"] and r["ransac_iters"] == o["e_info"]["iters"]
        assert esame(r["E"], o["E"]) < 1e-9
        assert rot_angle(r["R"], o["R"]) < 1e-4
        assert np.arccos(np.clip(r["t"] @ o["t"], -1, 1)) < 1e-3
        assert bool(r["lm_ran"]) == o["lm_ran"]
        assert np.abs(r["T0"] - o["T0"]).max() < 1e-7
        if o["lm_ran"]:
            assert bool(r["lm_reverted"]) == o["lm_reverted"]
            assert abs(r["r_norm"] - o["lm"]["r_norm"]) <= 1e-5 * o["lm"]["r_norm"]
        assert rot_angle(r["T"][:3, :3], o["T"][:3, :3]) < 1e-4
        # and the estimate is close to the synthetic ground truth
        assert rot_angle(r["R"], seq.R[i]) < 5e-3
    ms = pipe.stage_ms()
    assert ms[0] > 0 and ms[7] > 0
    pipe.close()


def test_sequence_chunking_and_subranges(ctx, seq):
    """Running a sub-range, or the same range twice, gives identical results (idempotence)."""
    pipe = api.SequencePipeline(seq.n_frames, 2000, ctx=ctx)
    pipe.upload(seq.kps, seq.descs)
    prm = api.default_params(seq.K.astype(np.float32))
    pipe.run(prm, 0, seq.n_pairs)
    a = pipe.download(0, seq.n_pairs).copy()
    pipe.run(prm, 2, 3)
    b = pipe.download(2, 3).copy()
    assert a[2:5].tobytes() == b.tobytes()
    pipe.run(prm, 0, seq.n_pairs)
    c = pipe.download(0, seq.n_pairs)
    assert a.tobytes() == c.tobytes()
    pipe.close()


def test_overlap_on_off_identical(ctx):
    """Two-stream group pipelining must not change any result (more than one group: 2100 pairs)."""
    s = synth.make_sequence(n_frames=2101, n=256, seed=synth.seed_for(3, 2))
    pipe = api.SequencePipeline(s.n_frames, 256, ctx=ctx)
    pipe.upload(s.kps, s.descs)
    prm = api.default_params(s.K.astype(np.float32))
    pipe.set_overlap(True)
    pipe.run(prm, 0, s.n_pairs)
    a = pipe.download(0, s.n_pairs).copy()
    pipe.set_overlap(False)
    pipe.run(prm, 0, s.n_pairs)
    b = pipe.download(0, s.n_pairs).copy()
    assert a.tobytes() == b.tobytes()
    assert (a["n_matches"] > 50).all()
    pipe.close()


def test_process_host_buffers_equals_upload_run_download(ctx):
    s = synth.make_sequence(n_frames=1201, n=256, seed=synth.seed_for(3, 3))
    pipe = api.SequencePipeline(s.n_frames, 256, ctx=ctx)
    prm = api.default_params(s.K.astype(np.float32))
    a = pipe.process(prm, s.kps, s.descs).copy()
    pipe2 = api.SequencePipeline(s.n_frames, 256, ctx=ctx)
    pipe2.upload(s.kps, s.descs)
    pipe2.run(prm, 0, s.n_pairs)
    b = pipe2.download(0, s.n_pairs)
    assert a.tobytes() == b.tobytes()
    pipe.close()
    pipe2.close()


def test_process_geometry_interleaved_or_not_identical(ctx, monkeypatch):
    """Host-buffer path: many small upload pieces, the geometry between the matcher pieces (mode 2) or after them
    (default) over repeated calls -- every run must give the bytes of upload + run + download."""
    monkeypatch.setenv("EPIVO_UPLOAD_DIV", "8")          # 74-pair pieces: 17 of them for 1200 pairs
    monkeypatch.setenv("EPIVO_UPLOAD_CAP", "1")
    s = synth.make_sequence(n_frames=1201, n=256, seed=synth.seed_for(3, 5))
    prm = api.default_params(s.K.astype(np.float32))
    ref_pipe = api.SequencePipeline(s.n_frames, 256, ctx=ctx)
    ref_pipe.upload(s.kps, s.descs)
    ref_pipe.run(prm, 0, s.n_pairs)
    ref = ref_pipe.download(0, s.n_pairs).tobytes()
    ref_pipe.close()
    pipe = api.SequencePipeline(s.n_frames, 256, ctx=ctx)
    for mode in (2, 0, 0, 2):
        pipe.set_overlap(mode)
        assert pipe.process(prm, s.kps, s.descs).tobytes() == ref, mode
    assert pipe.stage_ms()[0] > 0
    # LMedS, ratio matcher and plain Hamming (no plane pre-pass on the copy stream) through the interleaved form
    for kw in (dict(method=api.LMEDS, threshold=0.01), dict(norm=api.NORM_HAMMING, match_mode=api.MATCH_RATIO)):
        p2 = api.default_params(s.K.astype(np.float32), **kw)
        pipe.set_overlap(0)
        a = pipe.process(p2, s.kps, s.descs).tobytes()
        pipe.set_overlap(2)
        assert pipe.process(p2, s.kps, s.descs).tobytes() == a, kw
    pipe.close()


def test_ratio_mode_pipeline(ctx, seq):
    pipe = api.SequencePipeline(seq.n_frames, 2000, ctx=ctx)
    pipe.upload(seq.kps, seq.descs)
    prm = api.default_params(seq.K.astype(np.float32), norm=api.NORM_HAMMING, match_mode=api.MATCH_RATIO)
    pipe.run(prm, 0, 2)
    res = pipe.download(0, 2)
    for i in range(2):
        qi, ti, d = pipe.matches(i)
        w = O.ratio_match(seq.descs[i], seq.descs[i + 1], 0.8, O.NORM_HAMMING)
        assert np.array_equal(qi, w[0]) and np.array_equal(ti, w[1]) and np.array_equal(d, w[2])
        assert rot_angle(res[i]["R"], seq.R[i]) < 5e-3
    pipe.close()
