"""N2: pose chain + depth / point cloud (kitti_E.cpp:203-254) on the GPU vs the CPU restatement,
and the text formats the reference's viewers read."""
import numpy as np
import pytest

from epivo_b200 import api, io, synth
from oracle import pipeline as OP

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def run(ctx):
    seq = synth.make_sequence(n_frames=7, n=700, seed=synth.seed_for(3, 21))
    Kf = seq.K.astype(np.float32)
    pipe = api.SequencePipeline(seq.n_frames, 700, ctx=ctx)
    pipe.upload(seq.kps, seq.descs)
    prm = api.default_params(Kf)
    pipe.run(prm, 0, seq.n_pairs)
    res = pipe.download(0, seq.n_pairs)
    inl0, inl1 = [], []
    for i in range(seq.n_pairs):
        qi, ti, _ = pipe.matches(i)
        em, _ = pipe.masks(i)
        inl0.append(seq.kps[i][qi][em == 1])
        inl1.append(seq.kps[i + 1][ti][em == 1])
    yield seq, pipe, res, inl0, inl1, Kf
    pipe.close()


@pytest.mark.parametrize("with_scales", [False, True])
def test_chain_and_cloud_vs_oracle(run, with_scales):
    seq, pipe, res, inl0, inl1, Kf = run
    scales = np.linspace(0.6, 1.4, seq.n_pairs) if with_scales else None
    poses, pts, limits = pipe.cloud(scales)
    o_T, o_X, o_lim = OP.chain_and_cloud(res["T"], inl0, inl1, Kf, scales)
    assert poses.shape == o_T.shape
    assert np.abs(poses - o_T).max() < 1e-10                      # block scan vs sequential products
    assert np.array_equal(limits, o_lim)
    assert pts.shape == o_X.shape
    assert np.abs(pts - o_X).max() <= 1e-9 * max(1.0, np.abs(o_X).max())
    assert poses[0].tolist() == np.eye(4).tolist()


def test_cloud_subrange_and_counts_only(run):
    seq, pipe, res, inl0, inl1, Kf = run
    poses, pts, limits = pipe.cloud(None, first_pair=2, n_pairs=3)
    o_T, o_X, o_lim = OP.chain_and_cloud(res["T"][2:5], inl0[2:5], inl1[2:5], Kf)
    assert np.abs(poses - o_T).max() < 1e-10 and np.array_equal(limits, o_lim) and pts.shape == o_X.shape
    p2, none, l2 = pipe.cloud(None, with_points=False)
    assert none.shape[1] == 3 and l2.shape == (seq.n_pairs,)
    with pytest.raises(api.EpivoError):
        pipe.cloud(None, first_pair=0, n_pairs=seq.n_pairs + 5)


def test_viewer_file_formats_roundtrip(run, tmp_path):
    seq, pipe, res, inl0, inl1, Kf = run
    poses, pts, limits = pipe.cloud()
    io.write_cloud(str(tmp_path / "pts.cld"), pts)
    io.write_limits(str(tmp_path / "lims"), limits)
    io.write_poses(str(tmp_path / "kitti.T"), poses[:-1])          # all_T has one pose per pair
    assert np.array_equal(io.read_cloud(str(tmp_path / "pts.cld")), pts)
    assert np.array_equal(io.read_limits(str(tmp_path / "lims")), limits)
    assert np.array_equal(io.read_poses(str(tmp_path / "kitti.T")), poses[:-1])


@pytest.mark.parametrize("with_scales", [False, True])
def test_three_pose_chains_agree(run, ctx, with_scales):
    """D1 has three implementations -- the chain inside epivo_seq_cloud (single GPU), epivo_chain_poses (the gathered
    poses of a sharded sequence, same device block scan) and shard.chain_poses (host restatement): one rule in all
    three, including the reference's unguarded t / |t| (kitti_E.cpp:221), which turns a zero translation into NaN
    from that pair on."""
    from epivo_b200 import shard
    seq, pipe, res, inl0, inl1, Kf = run
    scales = np.linspace(0.6, 1.4, seq.n_pairs) if with_scales else None
    a, _, _ = pipe.cloud(scales, with_points=False)
    b = api.chain_poses(res["T"], scales, ctx=ctx)
    c = shard.chain_poses(res["T"], scales)
    assert np.array_equal(a, b)                                   # same kernel, same inputs
    assert np.abs(a - c).max() < 1e-10
    T = np.array(res["T"])
    T[2, :3, 3] = 0.0                                             # |t| = 0: not guarded, as in the reference
    b0, c0 = api.chain_poses(T, scales, ctx=ctx), shard.chain_poses(T, scales)
    for p in (b0, c0):
        assert np.isfinite(p[:3]).all() and np.isnan(p[3:, :3, 3]).all()       # positions are lost from pair 2 on
    assert api.chain_poses(np.zeros((0, 4, 4)), ctx=ctx).tolist() == [np.eye(4).tolist()]
