"""CPU: the post-LM loop restatement (kitti_E.cpp:203-254) and the viewer file formats."""
import numpy as np

from epivo_b200 import io, shard, synth
from oracle import pipeline as OP


def _rand_T(rng):
    T = np.eye(4)
    T[:3, :3] = synth.rodrigues(rng.normal(0, 0.05, 3))
    T[:3, 3] = rng.normal(0, 1, 3)
    return T


def test_chain_matches_host_chain_poses():
    rng = np.random.default_rng(5)
    Ts = np.array([_rand_T(rng) for _ in range(9)])
    scales = rng.uniform(0.5, 1.5, 9)
    empty = [np.zeros((0, 2), np.float32)] * 9
    all_T, X, lim = OP.chain_and_cloud(Ts, empty, empty, synth.KITTI_K, scales)
    assert X.shape == (0, 3) and np.array_equal(lim, np.zeros(9, np.int64))
    assert np.allclose(all_T, shard.chain_poses(Ts, scales), atol=1e-12)


def test_cloud_points_are_depth_from_parallax():
    """A point seen from two poses: X = d * K^-1 x0 reproduces the landmark when dT is the true motion."""
    rng = np.random.default_rng(6)
    K = synth.KITTI_K
    R = synth.rodrigues(np.array([0.01, -0.02, 0.005]))
    t = np.array([1.0, 0.1, 0.05])                          # sideways: parallax |B| ~ 1/depth > 1e-2 (kitti_E.cpp:248)
    Xw = np.column_stack([rng.uniform(-3, 3, 50), rng.uniform(-1, 1, 50), rng.uniform(4, 20, 50)])
    x0 = (K @ Xw.T).T
    x0 = (x0[:, :2] / x0[:, 2:]).astype(np.float32)
    X1 = (R @ Xw.T).T + t
    x1 = (K @ X1.T).T
    x1 = (x1[:, :2] / x1[:, 2:]).astype(np.float32)
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = R, t
    all_T, X, lim = OP.chain_and_cloud([T], [x0], [x1], K, [np.linalg.norm(t)])
    assert lim.tolist() == [0] and len(X) == 50
    assert np.abs(X - Xw).max() < 2e-2                      # float32 pixels, depth up to 20


def test_file_formats_roundtrip(tmp_path):
    rng = np.random.default_rng(7)
    pts = rng.normal(0, 10, (37, 3))
    poses = np.array([_rand_T(rng) for _ in range(5)])
    lims = np.array([0, 7, 7, 20, 31])
    io.write_cloud(str(tmp_path / "pts.cld"), pts)
    io.write_limits(str(tmp_path / "lims"), lims)
    io.write_poses(str(tmp_path / "kitti.T"), poses)
    assert np.array_equal(io.read_cloud(str(tmp_path / "pts.cld")), pts)
    assert np.array_equal(io.read_limits(str(tmp_path / "lims")), lims)
    assert np.array_equal(io.read_poses(str(tmp_path / "kitti.T")), poses)
    # layout the reference writes: a blank line after every point / after every 4x4 block
    assert open(tmp_path / "pts.cld").read().count("\n\n") == 37
    assert open(tmp_path / "kitti.T").read().count("\n\n") == 5
