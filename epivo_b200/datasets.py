"""N4 (host side): the ground-truth feeds of the reference drivers -- KITTI pose files, EuRoC ground-truth
CSVs with quaternion poses and timestamp association -- and the per-pair ground-truth step lengths the
drivers use to scale the unit-norm translation of every pair (kitti_E.cpp:216-222, euroc_E.cpp:300-304).
These produce the `scales` argument of `SequencePipeline.cloud` / `shard.chain_poses`.

Image loading, FAST/ORB detection and LK tracking stay out of scope (the benchmark inputs are keypoint /
descriptor sets); everything here is plain numpy on the host.
"""
from __future__ import annotations

import numpy as np


def load_kitti_poses(path: str) -> np.ndarray:
    """KITTI odometry `poses/NN.txt`: 12 space-separated numbers per line = a 3x4 row-major [R|t]
    (load_csv, kitti_E.cpp:18-34; reshaped to 4x3 and transposed at kitti_E.cpp:207-214) -> (n, 4, 4)."""
    rows = np.loadtxt(path, dtype=np.float64, ndmin=2)
    if rows.shape[1] != 12:
        raise ValueError(f"{path}: expected 12 numbers per line, got {rows.shape[1]}")
    T = np.tile(np.eye(4), (rows.shape[0], 1, 1))
    T[:, :3, :4] = rows.reshape(-1, 3, 4)
    return T


def gt_point_transforms(poses: np.ndarray) -> np.ndarray:
    """dT_i = (pT_i^-1 T_{i+1})^-1: the ground-truth point transform of pair i (kitti_E.cpp:216-217)."""
    poses = np.asarray(poses, dtype=np.float64)
    return np.stack([np.linalg.inv(np.linalg.inv(poses[i]) @ poses[i + 1]) for i in range(len(poses) - 1)])


def gt_scales(poses: np.ndarray) -> np.ndarray:
    """`scale = dT.block<3,1>(0,3).norm()` per pair (kitti_E.cpp:219)."""
    return np.linalg.norm(gt_point_transforms(poses)[:, :3, 3], axis=1)


def quat_to_R(q) -> np.ndarray:
    """(qw, qx, qy, qz) -> rotation matrix, normalising first (euroc_E.cpp:66-86)."""
    qw, qx, qy, qz = (float(v) for v in q)
    n = 1.0 / np.sqrt(qx * qx + qy * qy + qz * qz + qw * qw)
    qw, qx, qy, qz = qw * n, qx * n, qy * n, qz * n
    return np.array([[1.0 - 2.0 * qy * qy - 2.0 * qz * qz, 2.0 * qx * qy - 2.0 * qz * qw, 2.0 * qx * qz + 2.0 * qy * qw],
                     [2.0 * qx * qy + 2.0 * qz * qw, 1.0 - 2.0 * qx * qx - 2.0 * qz * qz, 2.0 * qy * qz - 2.0 * qx * qw],
                     [2.0 * qx * qz - 2.0 * qy * qw, 2.0 * qy * qz + 2.0 * qx * qw, 1.0 - 2.0 * qx * qx - 2.0 * qy * qy]])


def load_euroc_groundtruth(path: str) -> np.ndarray:
    """EuRoC `state_groundtruth_estimate0/data.csv`: header line, then comma-separated rows
    timestamp, px, py, pz, qw, qx, qy, qz, ... (load_csv with the first line skipped, euroc_E.cpp:23-44)."""
    return np.loadtxt(path, delimiter=",", skiprows=1, dtype=np.float64, ndmin=2)


def load_euroc_image_timestamps(path: str) -> np.ndarray:
    """EuRoC `cam0/data.csv`: header line, then `timestamp,filename`; the reference keeps the first column
    (load_fns, euroc_E.cpp:47-63) and later parses it with stod."""
    out = []
    with open(path) as f:
        next(f, None)
        for line in f:
            cell = line.split(",")[0].strip()
            if cell:
                out.append(float(cell))
    return np.array(out, dtype=np.float64)


EUROC_TS_TOLERANCE = 5760512 - 760576      # ns, euroc_E.cpp:228: just under the 5 ms period of the 200 Hz ground truth


def associate_euroc(gt: np.ndarray, image_ts: np.ndarray, tol: float = EUROC_TS_TOLERANCE) -> np.ndarray:
    """For every image timestamp the first ground-truth row within `tol` of it (euroc_E.cpp:226-246, which
    scans forward from a guessed row and takes the first hit) -> (n, 4, 4) body poses [R(q) | p]; raises if
    an image has no match (the reference asserts cnt_found == 2)."""
    ts = gt[:, 0]
    T = np.tile(np.eye(4), (len(image_ts), 1, 1))
    for i, t in enumerate(image_ts):
        j = int(np.searchsorted(ts, t - tol, side="right"))        # first row with ts > t - tol
        if j >= len(ts) or not abs(ts[j] - t) < tol:
            raise ValueError(f"no ground-truth row within {tol} ns of image timestamp {t:.0f}")
        T[i, :3, :3] = quat_to_R(gt[j, 4:8])
        T[i, :3, 3] = gt[j, 1:4]
    return T


def euroc_gt_scales(body_poses: np.ndarray, T_DC: np.ndarray) -> np.ndarray:
    """`dT = ((pT T_DC)^-1 (T T_DC))^-1`, scale = |dT.t| (euroc_E.cpp:300-301): camera-frame step lengths."""
    cam = np.asarray(body_poses) @ np.asarray(T_DC, dtype=np.float64)
    return gt_scales(cam)
