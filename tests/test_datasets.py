"""N4 host-side feeds: KITTI / EuRoC ground-truth loaders and per-pair scales (kitti_E.cpp:18-34,203-222,
euroc_E.cpp:23-86,218-246,300-304)."""
import numpy as np
import pytest

from epivo_b200 import datasets as D
from epivo_b200 import shard, synth


def _chain(n, seed):
    rng = np.random.default_rng(seed)
    T = [np.eye(4)]
    for _ in range(n - 1):
        d = np.eye(4)
        d[:3, :3] = synth.rodrigues(rng.normal(0, 0.02, 3))
        d[:3, 3] = rng.normal(0, 0.5, 3) + np.array([0, 0, 1.0])
        T.append(T[-1] @ d)
    return np.array(T)


def test_kitti_pose_file_roundtrip_and_scales(tmp_path):
    poses = _chain(12, 1)
    path = tmp_path / "00.txt"
    with open(path, "w") as f:
        for T in poses:
            f.write(" ".join("%.12e" % v for v in T[:3, :4].ravel()) + "\n")
    got = D.load_kitti_poses(str(path))
    assert got.shape == (12, 4, 4) and np.allclose(got, poses, atol=1e-11)
    dT = D.gt_point_transforms(got)
    s = D.gt_scales(got)
    assert np.allclose(s, np.linalg.norm(dT[:, :3, 3], axis=1)) and (s > 0).all()
    # chaining the ground-truth point transforms with their own scales reproduces the ground-truth trajectory
    # (kitti_E.cpp:218-228: cT <- cT dT^-1 starting from the identity = first-frame coordinates)
    chained = shard.chain_poses(dT, s)
    rel = np.array([np.linalg.inv(got[0]) @ T for T in got])
    assert np.allclose(chained, rel, atol=1e-9)
    with pytest.raises(ValueError):
        bad = tmp_path / "bad.txt"
        bad.write_text("1 2 3\n4 5 6\n")
        D.load_kitti_poses(str(bad))


def test_quat_to_R_matches_rodrigues():
    rng = np.random.default_rng(2)
    for _ in range(20):
        w = rng.normal(0, 1, 3)
        th = np.linalg.norm(w)
        q = np.concatenate([[np.cos(th / 2)], np.sin(th / 2) * w / th]) * rng.uniform(0.5, 2.0)   # un-normalised
        R = D.quat_to_R(q)
        assert np.allclose(R, synth.rodrigues(w), atol=1e-12)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-12)


def test_euroc_groundtruth_association(tmp_path):
    rng = np.random.default_rng(3)
    n_gt, period = 400, 5_000_000                                  # 200 Hz ground truth
    ts = 1_403_715_000_000_000_000 + np.arange(n_gt) * period
    p = np.cumsum(rng.normal(0, 0.01, (n_gt, 3)), axis=0)
    q = rng.normal(0, 1, (n_gt, 4))
    gt_path = tmp_path / "gt.csv"
    with open(gt_path, "w") as f:
        f.write("#timestamp,p_x,p_y,p_z,q_w,q_x,q_y,q_z,v_x,v_y,v_z\n")
        for i in range(n_gt):
            f.write(",".join(["%d" % ts[i]] + ["%.9f" % v for v in p[i]] + ["%.9f" % v for v in q[i]] + ["0", "0", "0"]) + "\n")
    img_rows = np.arange(10, 390, 10)                              # 20 Hz camera, offset by < half a period
    img_ts = ts[img_rows] + rng.integers(-2_000_000, 2_000_000, len(img_rows))
    cam_path = tmp_path / "cam.csv"
    with open(cam_path, "w") as f:
        f.write("#timestamp [ns],filename\n")
        for t in img_ts:
            f.write("%d,%d.png\n" % (t, t))
    gt = D.load_euroc_groundtruth(str(gt_path))
    its = D.load_euroc_image_timestamps(str(cam_path))
    assert gt.shape == (n_gt, 11) and len(its) == len(img_rows)
    T = D.associate_euroc(gt, its)
    for k in range(len(img_rows)):
        # the reference takes the FIRST row within the tolerance (a forward scan, euroc_E.cpp:227-236); the
        # tolerance is almost a full ground-truth period, so that can be the row before the nearest one
        j = next(r for r in range(n_gt) if abs(ts[r] - its[k]) < D.EUROC_TS_TOLERANCE)
        assert abs(j - img_rows[k]) <= 1
        assert np.allclose(T[k, :3, 3], p[j], atol=1e-8)
        assert np.allclose(T[k, :3, :3], D.quat_to_R(q[j]), atol=1e-7)
    T_DC = np.eye(4)
    T_DC[:3, :3] = synth.rodrigues(np.array([0.01, 1.2, -0.3]))
    T_DC[:3, 3] = [0.02, -0.06, 0.01]
    s = D.euroc_gt_scales(T, T_DC)
    assert s.shape == (len(img_rows) - 1,) and (s >= 0).all()
    with pytest.raises(ValueError):
        D.associate_euroc(gt, np.array([ts[-1] + 50 * period], dtype=np.float64))
